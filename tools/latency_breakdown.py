import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg
from opf_graph_neural_solver_b200 import model as M
BLG = pkg.get_BLG()
torch.manual_seed(0)
m = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
m.validate_topology = False
b, l, g, _ = pkg.data.make_batch(300, 4, seed=1)
b, l, g = b[:1].cuda(), l[:1].cuda(), g[:1].cuda()
out = m(b, l, g); out[2].sum().backward()
plan = m._last_plan; flat = m.flat_parameters()
lib = pkg.load_library()
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("raw forward (no grad)      %.1f us" % timeit(lambda: M._run_forward(m, plan, False, b, l, g, flat)))
print("raw forward (need_grad)    %.1f us" % timeit(lambda: M._run_forward(m, plan, True, b, l, g, flat)))
v, th, tot, last, ws = M._run_forward(m, plan, True, b, l, g, flat)
gt = torch.ones(1, device="cuda"); gp = torch.empty_like(flat)
st = torch.cuda.current_stream().cuda_stream
def bwd():
    rc = lib.gns_backward(plan.handle, flat.data_ptr(), b.data_ptr(), l.data_ptr(), g.data_ptr(), 1, 4, 20, 10, 1, 0.9,
                          gt.data_ptr(), None, None, None, gp.data_ptr(), ws.data_ptr(), ws.numel(), st)
    assert rc == 0
print("raw backward C call        %.1f us" % timeit(bwd))
def full():
    m.zero_grad(set_to_none=True)
    o = m(b, l, g); o[2].sum().backward()
print("module fwd+bwd             %.1f us" % timeit(full))
def fwd_only():
    o = m(b, l, g)
print("module fwd (grad enabled)  %.1f us" % timeit(fwd_only))
