"""Which overlap of the host pipeline is missing: run it with copies in / out switched off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg
from opf_graph_neural_solver_b200 import model as M, _lib
S, chunk = 65536, 4096
torch.manual_seed(0)
model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
model.validate_topology = False
b, l, g, _ = pkg.data.make_batch(300, 8192, seed=1)
host = [t.repeat(8, 1, 1).contiguous().pin_memory() for t in (b, l, g)]
var, const = pkg.data.pack_varying(*host)
var = tuple(t.pin_memory() for t in var)
out = [torch.empty(S, 300).pin_memory(), torch.empty(S, 300).pin_memory(), torch.empty(S).pin_memory(), torch.empty(S).pin_memory()]
lib = _lib.load_library()
dev = torch.device("cuda")
N, E, Gn = 300, 411, 69
cdev = [t.cuda() for t in const]
flat = model.flat_parameters()
plan = model.plan_for(host[1][:4].cuda(), host[2][:4].cuda(), N)

def pipeline(do_h2d=True, do_d2h=True, reuse_ws=False):
    comp = torch.cuda.current_stream(dev)
    h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    dbuf = [[torch.empty((chunk,) + tuple(t.shape[1:]), device=dev) for t in var] for _ in range(2)]
    full = [[torch.empty(chunk, n, c, device=dev) for n, c in ((N, 6), (E, 7), (Gn, 7))] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]; free = [torch.cuda.Event() for _ in range(2)]
    start = torch.cuda.Event(); start.record(comp); h2d.wait_event(start)
    keep = []
    for i, (a, e) in enumerate(M.chunk_bounds(S, chunk)):
        slot = i % 2
        with torch.cuda.stream(h2d):
            if i >= 2: h2d.wait_event(free[slot])
            if do_h2d:
                for dst, src in zip(dbuf[slot], var): dst[:e - a].copy_(src[a:e], non_blocking=True)
            d = [t[:e - a] for t in dbuf[slot]]
            lib.gns_expand_inputs(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), cdev[0].data_ptr(), cdev[1].data_ptr(),
                                  cdev[2].data_ptr(), e - a, N, E, Gn, full[slot][0].data_ptr(), full[slot][1].data_ptr(),
                                  full[slot][2].data_ptr(), h2d.cuda_stream)
            d = [t[:e - a] for t in full[slot]]
            ready[slot].record(h2d)
        comp.wait_event(ready[slot])
        res = M._run_forward(model, plan, False, d[0], d[1], d[2], flat)[:4]
        free[slot].record(comp)
        if do_d2h:
            done = torch.cuda.Event(); done.record(comp); d2h.wait_event(done)
            with torch.cuda.stream(d2h):
                for dst, src in zip(out, res):
                    src.record_stream(d2h); dst[a:e].copy_(src, non_blocking=True)
        keep.append(res)
    d2h.synchronize(); torch.cuda.synchronize()

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
# first fill the device staging with valid data
pipeline()
print("full pipeline        %.2f ms" % timeit(pipeline))
print("no D2H               %.2f ms" % timeit(lambda: pipeline(do_d2h=False)))
print("no H2D               %.2f ms" % timeit(lambda: pipeline(do_h2d=False)))
print("no copies at all     %.2f ms" % timeit(lambda: pipeline(do_h2d=False, do_d2h=False)))
t0 = time.perf_counter()
for _ in range(5):
    for (a, e) in M.chunk_bounds(S, chunk): pass
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); pipeline(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
