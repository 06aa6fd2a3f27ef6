import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg
torch.manual_seed(0)
model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
model.validate_topology = False
b, l, g, _ = pkg.data.make_batch(30, 4096, seed=1)
b, l, g = b.cuda(), l.cuda(), g.cuda()
for _ in range(4):
    model.zero_grad(set_to_none=True)
    out = model(b, l, g, *pkg.get_BLG())
    out[2].mean().backward()
torch.cuda.synchronize()
