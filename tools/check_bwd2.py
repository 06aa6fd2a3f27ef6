"""A/B of the two backward kernels (GNS_BWD2=0: first kernel, =1: warp-specialised kernel) against the f64 oracle:
per-tensor gradient error on small batches, then timing on full batches.  Usage: check_bwd2.py [parity|time|all]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg
from oracle import gns_oracle as orc

BLG = pkg.get_BLG()


def grads_of(model, b, l, g, which):
    os.environ["GNS_BWD2"] = which
    model.zero_grad(set_to_none=True)
    out = model(b, l, g, *BLG)
    out[2].mean().backward()
    torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters()}, out[2].detach().clone()


def parity(n_bus, S, multi, L=20, K=4):
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=multi).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(n_bus, S, seed=7)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (_, _, otot, _), want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=K,
                                                   latent_dim=L, gamma=0.9, multiple_phi=multi)
    b, l, g = buses.cuda(), lines.cuda(), gens.cuda()
    gmax = max(float(w.abs().max()) for w in want.values())
    for which in ("0", "1"):
        got, tot = grads_of(model, b, l, g, which)
        rows = []
        for n, w in want.items():
            err = float((got[n].cpu().double() - w.double()).abs().max())
            rows.append((err / gmax, n, err, float(w.abs().max())))
        rows.sort(reverse=True)
        lerr = float(((tot.cpu().double() - otot).abs() / otot.abs()).max())
        print(f"case{n_bus} S={S} multi={multi} L={L} K={K} BWD2={which}: worst rel-to-gmax {rows[0][0]:.3e} (gmax {gmax:.3e}) loss rel {lerr:.2e}",
              "OK" if rows[0][0] <= 1e-3 else "FAIL", flush=True)
        if rows[0][0] > 1e-4:
            for r in rows[:14]:
                print(f"    {r[1]:32s} err {r[2]:.3e}  max|g| {r[3]:.3e}")
    got2, _ = grads_of(model, b, l, g, "1")
    same = all(torch.equal(got[n], got2[n]) for n in got)
    print(f"    BWD2=1 run-to-run bit-identical: {same}", flush=True)


def timing(n_bus, S, which, K=4, L=20, reps=5):
    os.environ["GNS_BWD2"] = which
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=True).cuda()
    model.validate_topology = False
    base = min(S, 4096)
    b, l, g, _ = pkg.data.make_batch(n_bus, base, seed=1)
    rep = (S + base - 1) // base
    b, l, g = (t.repeat(rep, 1, 1)[:S].contiguous().cuda() for t in (b, l, g))

    def step(mode):
        if mode == "fwd":
            out = model(b, l, g, *BLG)
            return out
        model.zero_grad(set_to_none=True)
        out = model(b, l, g, *BLG)
        out[2].mean().backward()

    res = {}
    for mode in ("fwd", "both"):
        for _ in range(3):
            step(mode)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            step(mode); ev[i + 1].record()
        torch.cuda.synchronize()
        res[mode] = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))[reps // 2]
    print(f"case{n_bus} S={S} K={K} L={L} BWD2={which}: train-fwd {res['fwd']:.3f} ms, fwd+bwd {res['both']:.3f} ms "
          f"(bwd ~{res['both'] - res['fwd']:.3f}) -> {S / res['both'] / 1e3:.3f} M grids/s", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    if mode in ("parity", "all"):
        parity(300, 3, True)
        parity(300, 2, False)
        parity(118, 9, True)
        parity(300, 301, True, L=10, K=3)
    if mode in ("time", "all"):
        for which in ("0", "1"):
            timing(300, 16384, which)
        for which in ("0", "1"):
            timing(118, 16384, which)
