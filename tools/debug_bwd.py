import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import opf_graph_neural_solver_b200 as pkg
from oracle import gns_oracle as orc
from test_backward_gpu import per_tensor_report
BLG = pkg.get_BLG()
for (n_bus, S, K, L, multi) in [(14, 4, 1, 20, True), (14, 4, 2, 20, True), (14, 4, 4, 20, True), (14, 4, 2, 20, False), (30, 9, 4, 10, True)]:
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=multi).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(n_bus, S, seed=7)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (_, _, otot, _), want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=K,
                                                   latent_dim=L, gamma=0.9, multiple_phi=multi)
    out = model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)
    out[2].mean().backward()
    got = {n: (p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}
    print(f"=== case{n_bus} S={S} K={K} L={L} multi={multi}: loss {float(out[2].mean()):.6f} vs {float(otot.mean()):.6f}")
    print(per_tensor_report(got, want))
