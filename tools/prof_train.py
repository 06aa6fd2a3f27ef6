"""Small fwd+bwd run for profiling: 4 training steps.  Usage: prof_train.py [S=2048] [n_bus=300] [latent=20] [K=4]
(GNS_BWD2=1 in the environment selects the warp-specialised backward kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg
arg = lambda i, d: int(sys.argv[i]) if len(sys.argv) > i else d
S, n_bus, L, K = arg(1, 2048), arg(2, 300), arg(3, 20), arg(4, 4)
torch.manual_seed(0)
model = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=True).cuda()
model.validate_topology = False
b, l, g, _ = pkg.data.make_batch(n_bus, S, seed=1)
b, l, g = b.cuda(), l.cuda(), g.cuda()
for _ in range(4):
    model.zero_grad(set_to_none=True)
    out = model(b, l, g, *pkg.get_BLG())
    out[2].mean().backward()
torch.cuda.synchronize()
print("ok", float(out[2].mean().detach()))
