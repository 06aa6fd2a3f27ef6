"""Evaluation metrics of the reference's ``evaluate.py`` (ref GNS/evaluate.py:15-18, 89-148), batched.

The reference script loops over pickled grids, runs pypower's Newton-Raphson and the GNS one
sample at a time, and prints mean / std of |dtheta| and |dv| against the Newton-Raphson solution,
the mean / std of the last-step loss, and percentiles of the active line-flow percentage
difference.  Here the GNS side is ONE batched call of the sm_100a forward kernel and the metric
arithmetic is the same numpy code, vectorised over the samples.  The power-flow solution the GNS is
compared against is an INPUT (arrays in pypower's output convention: ``VM`` in p.u., ``VA`` in
degrees): the product does not ship a Newton-Raphson solver.
"""
from __future__ import annotations

import numpy as np
import torch


def active_line_flow(V, theta, x, src, dst):
    """``1/x * V[src] * V[dst] * sin(theta[src] - theta[dst])`` per line (ref GNS/evaluate.py:15-18),
    for one sample (``V [N]``) or a batch (``V [S, N]``, ``x / src / dst [E]`` or ``[S, E]``);
    ``src`` / ``dst`` are the 1-based bus numbers of the tables."""
    V, theta, x = np.asarray(V), np.asarray(theta), np.asarray(x)
    src = np.asarray(src).astype(int) - 1
    dst = np.asarray(dst).astype(int) - 1
    if V.ndim == 1:
        return 1 / x * (V[src] * V[dst] * np.sin(theta[src] - theta[dst]))
    if src.ndim == 1:
        src, dst = np.broadcast_to(src, (V.shape[0],) + src.shape), np.broadcast_to(dst, (V.shape[0],) + dst.shape)
    take = lambda a, i: np.take_along_axis(a, i, axis=1)
    return 1 / x * (take(V, src) * take(V, dst) * np.sin(take(theta, src) - take(theta, dst)))


def comparison_metrics(gns_v, gns_theta, last_losses, lines, ref_vm, ref_va_deg, ref_branch_x, ref_f_bus, ref_t_bus,
                       reference_degree_quirk: bool = True):
    """The numbers ``evaluate.py`` prints (ref GNS/evaluate.py:89-148) for ``S`` samples.

    gns_v, gns_theta [S, N] (theta in rad), last_losses [S], lines [S, E, 7] (packed GNS line rows:
    ``lines[:, :, 3]`` is x, cols 0 / 1 the bus numbers, ref :87); ref_vm [S, N] in p.u., ref_va_deg [S, N] in
    DEGREES as pypower returns them, ref_branch_x / ref_f_bus / ref_t_bus [S, E] or [E].

    reference_degree_quirk: the reference feeds the Newton-Raphson angles to ``active_line_flow`` in degrees
    (``solved_grid['bus'][:, 8]``, ref :40) although ``np.sin`` expects radians; True reproduces that, False
    converts first.  The theta difference itself is taken in radians like the reference (ref :101).
    """
    gns_v, gns_theta = np.asarray(gns_v, np.float32), np.asarray(gns_theta, np.float32)
    ref_vm, ref_va_deg = np.asarray(ref_vm, np.float32), np.asarray(ref_va_deg, np.float32)
    lines = np.asarray(lines, np.float32)
    ref_theta = np.deg2rad(ref_va_deg)
    theta_diff = np.abs(gns_theta - ref_theta)                    # ref :101-103
    v_diff = np.abs(gns_v - ref_vm)                               # ref :109
    gns_alf = active_line_flow(gns_v, gns_theta, lines[:, :, 3], lines[:, :, 0], lines[:, :, 1]).astype(np.float32)
    nr_angles = ref_va_deg if reference_degree_quirk else ref_theta
    nr_alf = active_line_flow(ref_vm, nr_angles, np.asarray(ref_branch_x), ref_f_bus, ref_t_bus).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        pct = np.abs((nr_alf - gns_alf) / nr_alf) * 100            # ref :124-127
        theta_err = np.abs((gns_theta - ref_theta) / ref_theta) * 100   # ref :117
        v_err = np.abs((gns_v - ref_vm) / ref_vm) * 100                 # ref :119
    low = np.sort(pct, axis=None)[: int(pct.size / 2)]            # "only lowest 50%", ref :129
    return {
        "theta_diff_mean": float(np.mean(theta_diff)), "theta_diff_std": float(np.std(theta_diff)),
        "v_diff_mean": float(np.mean(v_diff)), "v_diff_std": float(np.std(v_diff)),
        "theta_pct_error_mean": float(np.mean(theta_err[np.isfinite(theta_err)])) if np.isfinite(theta_err).any() else float("nan"),
        "v_pct_error_mean": float(np.mean(v_err)),
        "last_loss_mean": float(np.mean(last_losses)), "last_loss_std": float(np.std(last_losses)),
        "line_flow_pct_p20": float(np.percentile(low, 20)), "line_flow_pct_median": float(np.median(low)),
        "line_flow_pct_p80": float(np.percentile(low, 80)),
    }


@torch.no_grad()
def evaluate_model(model, buses, lines, generators, ref_vm, ref_va_deg, ref_branch_x, ref_f_bus, ref_t_bus,
                   reference_degree_quirk: bool = True):
    """Run the GNS on all samples in one batched call (instead of the per-sample loop, ref :73-87) and
    compare with a given power-flow solution.  Returns ``(metrics, (v, theta, last_loss))``."""
    v, theta, _, last = model(buses, lines, generators)
    v, theta, last = v.detach().cpu().numpy(), theta.detach().cpu().numpy(), last.detach().cpu().numpy()
    m = comparison_metrics(v, theta, last, lines.detach().cpu().numpy(), ref_vm, ref_va_deg, ref_branch_x, ref_f_bus,
                           ref_t_bus, reference_degree_quirk)
    return m, (v, theta, last)
