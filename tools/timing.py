"""Timing of the forward kernel and the training step (device-resident inputs, CUDA events); library of baseline_configs.py."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg

def flops_fwd(N, E, K, L, H, multi=True):
    mac_line = 3 * ((5 + L) * H + H * H + H * L) if multi else ((5 + L) * H + H * H + H)
    mac_bus = 2 * ((4 + 2 * L) * H + H * H + H) + ((4 + 2 * L) * H + H * H + H * L)
    return 2 * K * (E * mac_line + N * mac_bus)

def run(n_bus, S, K=4, L=20, reps=5, train=False):
    """train: False = inference forward, True = forward + backward, "fwdonly" = training forward alone"""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=True).cuda()
    model.validate_topology = False
    base = min(S, 4096)
    b, l, g, _ = pkg.data.make_batch(n_bus, base, seed=1)
    rep = (S + base - 1) // base
    b, l, g = (t.repeat(rep, 1, 1)[:S].contiguous().cuda() for t in (b, l, g))
    E = l.shape[1]
    BLG = pkg.get_BLG()
    def step():
        if train == "fwdonly":
            out = model(b, l, g, *BLG)
        elif train:
            model.zero_grad(set_to_none=True)
            out = model(b, l, g, *BLG)
            out[2].mean().backward()
        else:
            with torch.no_grad():
                model(b, l, g, *BLG)
    for _ in range(3): step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        step(); ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))[reps // 2]
    fl = flops_fwd(n_bus, E, K, L, 10) * (3 if train is True else 1)
    gps = S / (ms * 1e-3)
    info = model._last_plan.launch_info(S, K, L, 10, True, backward=(train is True))
    print(f"case{n_bus} S={S} K={K} L={L} {'fwd+bwd' if train is True else ('fwd(train)' if train else 'fwd')}: {ms:.3f} ms  {gps/1e6:.3f} M grids/s  "
          f"{gps*fl/1e12:.2f} TFLOP/s algorithmic  geom={info} env VG={os.environ.get('GNS_FWD_VG')} NGQ={os.environ.get('GNS_FWD_NGQ')}", flush=True)

if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    train = mode == "train"
    for n_bus, S in [(300, 16384), (118, 32768), (30, 65536), (14, 131072)]:
        run(n_bus, S, train=train)
    run(300, 4096, K=8, L=64, train=train)
