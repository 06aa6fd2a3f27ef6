"""Where the end-to-end host pipeline spends its time: copies alone, chunked kernels alone, both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg
from opf_graph_neural_solver_b200 import model as M
S, chunk = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
model.validate_topology = False
b, l, g, _ = pkg.data.make_batch(300, 8192, seed=1)
host = [t.repeat(8, 1, 1).contiguous().pin_memory() for t in (b, l, g)]
var, const = pkg.data.pack_varying(*host)
var = tuple(t.pin_memory() for t in var)
dev = [t.cuda() for t in host]
out = [torch.empty(S, 300).pin_memory(), torch.empty(S, 300).pin_memory(), torch.empty(S).pin_memory(), torch.empty(S).pin_memory()]
BLG = pkg.get_BLG()
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
dvar = [torch.empty_like(t, device="cuda") for t in var]
def h2d_only():
    for (a, e) in M.chunk_bounds(S, chunk):
        for d, s in zip(dvar, var): d[a:e].copy_(s[a:e], non_blocking=True)
def whole():
    with torch.no_grad(): model(*dev, *BLG)
def chunks_dev():
    with torch.no_grad():
        for (a, e) in M.chunk_bounds(S, chunk): model(dev[0][a:e], dev[1][a:e], dev[2][a:e], *BLG)
res = None
def d2h_only():
    for (a, e) in M.chunk_bounds(S, chunk):
        for d, s in zip(out, res): d[a:e].copy_(s[a:e], non_blocking=True)
with torch.no_grad(): res = model(*dev, *BLG)
print(f"chunk {chunk}: {len(M.chunk_bounds(S, chunk))} chunks")
print("H2D compact only      %.2f ms" % timeit(h2d_only))
print("D2H only              %.2f ms" % timeit(d2h_only))
print("forward whole batch   %.2f ms" % timeit(whole))
print("forward in chunks     %.2f ms" % timeit(chunks_dev))
print("pipeline compact      %.2f ms" % timeit(lambda: model.infer_host_compact(var, const, out=out, chunk=chunk)))
print("pipeline full rows    %.2f ms" % timeit(lambda: model.infer_host(*host, out=out, chunk=chunk)))
