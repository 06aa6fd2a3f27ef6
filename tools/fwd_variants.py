"""Inference-forward timing of launch-geometry variants (environment knobs read by the library):
python tools/fwd_variants.py [n_bus] [S]   — prints ms per call and M grids/s for each variant, plus a result check."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg

n_bus = int(sys.argv[1]) if len(sys.argv) > 1 else 300
S = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
BLG = pkg.get_BLG()
base = min(S, 4096)
b, l, g, _ = pkg.data.make_batch(n_bus, base, seed=1)
rep = (S + base - 1) // base
b, l, g = (t.repeat(rep, 1, 1)[:S].contiguous().cuda() for t in (b, l, g))
variants = [{}, {"GNS_FWD_VG": "1", "GNS_FWD_NGQ": "2"}, {"GNS_FWD_VG": "1", "GNS_FWD_NGQ": "1"}, {"GNS_FWD_VG": "2", "GNS_FWD_NGQ": "2"}]
if os.environ.get("FWD_VARIANTS") == "cap":
    variants = [{}, {"GNS_NO_TMA": "1"}, {"GNS_DEG_CAP": "2"}, {"GNS_DEG_CAP": "2", "GNS_NO_TMA": "1"}, {"GNS_DEG_CAP": "4"}]
ref = None
for env in variants:
    for k in ("GNS_FWD_VG", "GNS_FWD_NGQ", "GNS_DEG_CAP", "GNS_NO_TMA"):
        os.environ.pop(k, None)
    os.environ.update(env)
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    model.validate_topology = False
    try:
        with torch.no_grad():
            for _ in range(3):
                out = model(b, l, g, *BLG)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            for i in range(5):
                out = model(b, l, g, *BLG); ev[i + 1].record()
            torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(5))[2]
        if ref is None:
            ref = [t.clone() for t in out]
        err = max(float((o - r).abs().max()) for o, r in zip(out, ref))
        print(f"case{n_bus} S={S} {env}: {ms:.3f} ms -> {S / ms / 1e3:.3f} M grids/s   max|diff vs default| {err:.2e}", flush=True)
    except Exception as ex:
        print(f"case{n_bus} S={S} {env}: failed: {str(ex)[:150]}", flush=True)
