"""Small inference run for profiling: 3 forward calls of S case300 grids (default 16384)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.manual_seed(0)
model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
model.validate_topology = False
b, l, g, _ = pkg.data.make_batch(300, min(S, 4096), seed=1)
rep = (S + b.shape[0] - 1) // b.shape[0]
b, l, g = (t.repeat(rep, 1, 1)[:S].contiguous().cuda() for t in (b, l, g))
with torch.no_grad():
    for _ in range(3):
        out = model(b, l, g, *pkg.get_BLG())
torch.cuda.synchronize()
print("ok", float(out[2].mean()))
