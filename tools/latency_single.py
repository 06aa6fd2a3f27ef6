"""Per-call latency of the drop-in module in the reference's own usage pattern (one grid per call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg
BLG = pkg.get_BLG()
for n_bus in (14, 300):
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    b, l, g, _ = pkg.data.make_batch(n_bus, 64, seed=1)
    for mode in ("cpu tensors, validate", "cuda tensors, validate", "cuda tensors, no validate"):
        model.validate_topology = "no validate" not in mode
        bb, ll, gg = (b, l, g) if mode.startswith("cpu") else (b.cuda(), l.cuda(), g.cuda())
        for train in (False, True):
            def step(i):
                if train:
                    out = model(bb[i], ll[i], gg[i], *BLG); out[2].backward()
                else:
                    with torch.no_grad(): model(bb[i], ll[i], gg[i], *BLG)
            for i in range(8): step(i)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for i in range(64): step(i)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 64
            print(f"case{n_bus} {mode:28s} {'fwd+bwd' if train else 'fwd    '} {dt*1e6:8.1f} us/grid", flush=True)
