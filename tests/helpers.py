"""Shared helpers of the parity tests (the oracle is the checker, never the product)."""
import glob
import os

import numpy as np
import torch

from oracle import gns_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# FP32 parity tolerances (BASELINE.json north_star: rel 1e-4 per bus, 1e-3 on loss / grads)
TOL_BUS = 1e-5      # measured ~1e-6; north_star allows 1e-4 (a regression of 10x is caught)
TOL_LOSS = 1e-3
TOL_GRAD = 1e-3


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "ref_case14_*.npz")))


def load_golden(path):
    z = np.load(path)
    lat, hid, K, multi = (int(x) for x in z["hyper"])
    names = orc.param_names(K, bool(multi))
    return dict(
        name=os.path.basename(path)[:-4], latent_dim=lat, hidden_dim=hid, K=K, multiple_phi=bool(multi),
        gamma=float(z["gamma"]),
        buses=torch.from_numpy(z["buses"]), lines=torch.from_numpy(z["lines"]), gens=torch.from_numpy(z["gens"]),
        v=torch.from_numpy(z["v"]), theta=torch.from_numpy(z["theta"]),
        total_loss=torch.from_numpy(z["total_loss"]), last_loss=torch.from_numpy(z["last_loss"]),
        params={n: torch.from_numpy(z["param/" + n]) for n in names},
        grads={n: torch.from_numpy(z["grad/" + n]) for n in names},
        none_grads=[str(s) for s in z["none_grads"]],
    )


def assert_bus_close(got, want, what):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    err = (got - want).abs()
    bound = TOL_BUS * want.abs().clamp_min(1.0)          # TOL_BUS relative per bus (absolute near zero angles)
    assert bool((err <= bound).all()), f"{what}: max err {err.max():.3e}"


def assert_loss_close(got, want, what):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    rel = ((got - want).abs() / want.abs().clamp_min(1e-12)).max()
    assert rel <= TOL_LOSS, f"{what}: max rel err {rel:.3e}"


def assert_grads_close(got: dict, want: dict, what):
    """Global scaling: |dg| <= 1e-3 * max_all(|g|) + 1e-6."""
    gmax = max(float(w.abs().max()) for w in want.values())
    worst = 0.0
    for n, w in want.items():
        worst = max(worst, float((got[n].detach().cpu().double() - w.double()).abs().max()))
    assert worst <= TOL_GRAD * gmax + 1e-6, f"{what}: max |dgrad| {worst:.3e} vs max|grad| {gmax:.3e}"
    return worst, gmax
