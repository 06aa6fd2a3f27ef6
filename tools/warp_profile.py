"""Per-warp line load of a plan (slots are lanes: a warp walks max(in-lines of its lanes) iterations of the line loop):
python tools/warp_profile.py [n_bus ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import opf_graph_neural_solver_b200 as pkg
from opf_graph_neural_solver_b200.plan import TopologyPlan
for n in [int(a) for a in sys.argv[1:]] or [300, 118]:
    b, l, g, _ = pkg.data.make_batch(n, 1, seed=1)
    p = TopologyPlan.from_tensors(l, g, n, device=0)
    deg = p.export("slot_in_end") - p.export("slot_in_begin")
    print(f"case{n}: {len(deg)} slots, {deg.sum()} lines, lines per slot histogram {np.bincount(deg).tolist()}")
    for w in range((len(deg) + 31) // 32):
        d = deg[w * 32:(w + 1) * 32]
        print(f"   warp {w}: {len(d)} lanes, max {d.max()}, sum {d.sum()}")
