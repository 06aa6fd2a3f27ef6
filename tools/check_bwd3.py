"""Parity (vs the f64 oracle) and timing of the fragment-space backward kernel (GNS_BWD3=1) against the first one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opf_graph_neural_solver_b200 as pkg
from oracle import gns_oracle as orc
import check_bwd2 as c2

BLG = pkg.get_BLG()

def setk(which):
    os.environ["GNS_BWD2"] = "0"
    os.environ["GNS_BWD3"] = "1" if which == "3" else "0"

def parity(n_bus, S, L=20, K=4):
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(n_bus, S, seed=7)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (_, _, otot, _), want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=K,
                                                   latent_dim=L, gamma=0.9, multiple_phi=True)
    b, l, g = buses.cuda(), lines.cuda(), gens.cuda()
    gmax = max(float(w.abs().max()) for w in want.values())
    res = {}
    for which in ("0", "3"):
        setk(which)
        model.zero_grad(set_to_none=True)
        out = model(b, l, g, *BLG)
        out[2].mean().backward()
        torch.cuda.synchronize()
        got = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
        res[which] = got
        rows = sorted(((float((got[n].cpu().double() - w.double()).abs().max()) / gmax, n, float(w.abs().max())) for n, w in want.items()), reverse=True)
        print(f"case{n_bus} S={S} L={L} K={K} kernel {which}: worst rel-to-gmax {rows[0][0]:.3e} (gmax {gmax:.3e})", "OK" if rows[0][0] <= 1e-3 else "FAIL", flush=True)
        if rows[0][0] > 1e-4:
            for r in rows[:16]:
                print(f"    {r[1]:32s} err/gmax {r[0]:.3e}  max|g| {r[2]:.3e}")
    setk("3")
    model.zero_grad(set_to_none=True)
    out = model(b, l, g, *BLG); out[2].mean().backward(); torch.cuda.synchronize()
    print("    run-to-run bit-identical:", all(torch.equal(res["3"][n], p.grad) for n, p in model.named_parameters()), flush=True)

def timing(n_bus, S, which):
    setk(which)
    c2_env = os.environ.get("GNS_BWD2")
    # reuse the timing loop of check_bwd2 without letting it touch GNS_BWD2
    import types
    src = c2.timing
    os.environ["GNS_BWD2"] = "0"
    src(n_bus, S, "0")

def geometry(n_bus, S=16384, L=20, K=4):
    """what the library would launch for the backward pass of this case (gns_launch_info)"""
    buses, lines, gens, _ = pkg.data.make_batch(n_bus, 1, seed=7)
    model = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=True).cuda()
    plan = model.plan_for(lines.cuda(), gens.cuda(), n_bus)
    print(f"    case{n_bus} backward geometry:", plan.launch_info(S, K, L, 10, True, backward=True), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    for env in ({"GNS_BWD3_TILES": "1"}, {"GNS_BWD3_TILES": "1", "GNS_DETERMINISTIC": "0"}, {"GNS_BWD3_TILES": "2"}) if mode == "geom" else ():
        os.environ.pop("GNS_DETERMINISTIC", None)
        os.environ.update(env); setk("3"); print(env); geometry(300); geometry(118)
    extra = [a for a in sys.argv[2:] if "=" in a]
    for a in extra:
        k, v = a.split("=", 1); os.environ[k] = v
    if mode in ("parity", "all"):
        parity(300, 3)
        parity(118, 9)
        parity(300, 301, L=10, K=3)
    if mode in ("time", "all"):
        for which in ("0", "3"):
            print("kernel", which, flush=True)
            timing(300, 16384, which)
        for which in ("0", "3"):
            print("kernel", which, flush=True)
            timing(118, 16384, which)
