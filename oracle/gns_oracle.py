"""CPU oracle for the Graph Neural Solver hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The shipped path (``opf-graph-neural-solver_b200``)
never imports anything from ``oracle/`` and has no CPU fallback.

It restates, batched over a shared topology and in plain PyTorch (float32 or
float64), the algorithm of the reference implementation:

* 3-layer MLP ``Linear-LeakyReLU(0.01)-Linear-LeakyReLU-Linear``  ref GNS/main.py:17-31
* per-step un-shared nets and their state_dict names               ref GNS/main.py:108-138
* state init, K message-passing steps, discounted loss, clamp      ref GNS/main.py:140-202
* slack redistribution / reactive balancing                        ref GNS/main.py:34-78
* Kirchhoff mismatch                                               ref GNS/main.py:80-104
* packed column maps                                               ref GNS/utils.py:4-13

Parity pin: the reference ships no tests and no expected outputs ("parity
unpinned" by the reference itself).  This restatement is pinned instead by golden
vectors generated from the *live* reference (``tests/golden/make_golden.py`` imports
``/root/reference/GNS/main.py`` unmodified behind a one-line ``torch_scatter``
shim); ``tests/test_oracle_golden.py`` checks the oracle against them.

Quirks of the reference that are part of the contract are kept:
bus numbers re-used as line numbers in the physics gathers (Q1), receiver-only
messages (Q2), single-phi scalar message into latent column 0 (Q3), dQ == 0 up to
cancellation noise (Q4), ``v_f / tau^2`` in the loss term (Q5).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# column maps, ref GNS/utils.py:5-9
BUS_COLS = {"bus_i": 0, "type": 1, "Pd": 2, "Qd": 3, "Gs": 4, "Bs": 5}
LINE_COLS = {"f_bus": 0, "t_bus": 1, "r": 2, "x": 3, "b": 4, "tau": 5, "theta": 6}
GEN_COLS = {"bus_i": 0, "Pmax": 1, "Pmin": 2, "Pg_set": 3, "vg": 4, "qg": 5, "Pg": 6}

LRELU_SLOPE = 0.01  # nn.LeakyReLU() default, ref GNS/main.py:23


def net_names(multiple_phi: bool):
    """Module order as registered by the reference ctor, ref GNS/main.py:113-122."""
    phis = ["phi_v", "phi_theta", "phi_m"] if multiple_phi else ["phi"]
    return phis + ["L_theta", "L_v", "L_m"]


def net_dims(name: str, latent_dim: int, hidden_dim: int):
    """(dim_in, hidden, dim_out) per net, ref GNS/main.py:126-134."""
    if name in ("phi_v", "phi_theta", "phi_m"):
        return 5 + latent_dim, hidden_dim, latent_dim
    if name == "phi":
        return 5 + latent_dim, hidden_dim, 1
    if name in ("L_theta", "L_v"):
        return 4 + 2 * latent_dim, hidden_dim, 1
    if name == "L_m":
        return 4 + 2 * latent_dim, hidden_dim, latent_dim
    raise KeyError(name)


def param_names(K: int, multiple_phi: bool):
    """state_dict key order of the reference module."""
    out = []
    for net in net_names(multiple_phi):
        for k in range(K):
            for lin in ("linear1", "linear2", "linear4"):
                out.append(f"{net}.{k}.{lin}.weight")
                out.append(f"{net}.{k}.{lin}.bias")
    return out


def init_params(latent_dim=10, hidden_dim=10, K=30, multiple_phi=False, seed=0,
                dtype=torch.float32):
    """Reference-identical random init: nn.Linear default init drawn in the
    k-major construction order of ref GNS/main.py:124-134 under ``seed``."""
    torch.manual_seed(seed)
    per_k = (["phi_v", "phi_theta", "phi_m"] if multiple_phi else ["phi"]) + \
        ["L_theta", "L_v", "L_m"]
    p = {}
    for k in range(K):
        for net in per_k:
            din, hid, dout = net_dims(net, latent_dim, hidden_dim)
            for lin, (i, o) in (("linear1", (din, hid)), ("linear2", (hid, hid)),
                                ("linear4", (hid, dout))):
                layer = torch.nn.Linear(i, o)
                p[f"{net}.{k}.{lin}.weight"] = layer.weight.detach().to(dtype)
                p[f"{net}.{k}.{lin}.bias"] = layer.bias.detach().to(dtype)
    return {n: p[n] for n in param_names(K, multiple_phi)}


def _mlp(p, net, k, x):
    h = F.leaky_relu(F.linear(x, p[f"{net}.{k}.linear1.weight"], p[f"{net}.{k}.linear1.bias"]),
                     LRELU_SLOPE)
    h = F.leaky_relu(F.linear(h, p[f"{net}.{k}.linear2.weight"], p[f"{net}.{k}.linear2.bias"]),
                     LRELU_SLOPE)
    return F.linear(h, p[f"{net}.{k}.linear4.weight"], p[f"{net}.{k}.linear4.bias"])


def topology(lines, gens):
    """0-based index vectors of the shared topology (ref GNS/main.py:35-36,144,153)."""
    l0 = lines[0] if lines.dim() == 3 else lines
    g0 = gens[0] if gens.dim() == 3 else gens
    f = l0[:, 0].long() - 1
    t = l0[:, 1].long() - 1
    gb = g0[:, 0].long() - 1
    return f, t, gb


def _bus_sum(vals, idx, n_bus):
    """scatter_add of [S,E(,C)] values onto buses along dim 1."""
    shape = list(vals.shape)
    shape[1] = n_bus
    return torch.zeros(shape, dtype=vals.dtype).index_add_(1, idx, vals)


def physics(v, theta, buses, lines, gens, f, t, gb):
    """One evaluation of the slack redistribution + Kirchhoff mismatch on the
    updated state; returns (dP, dQ, Pg, qg, p_global).  SURVEY.md App. A.3-A.5,
    ref GNS/main.py:34-104.  All tensors are batched [S, .]."""
    N = buses.shape[1]
    Pd, Qd, Gs, Bs = buses[..., 2], buses[..., 3], buses[..., 4], buses[..., 5]
    r, x, b, tau, sh = (lines[..., c] for c in range(2, 7))
    Pmax, Pmin, Pset = gens[..., 1], gens[..., 2], gens[..., 3]

    Y = 1.0 / torch.sqrt(r.pow(2) + x.pow(2))
    D = theta[:, f] - theta[:, t]                      # per-line angle difference
    # alias gathers: per-line vectors indexed by BUS numbers (quirk Q1)
    Yf, tauf, shf, bf, Df = Y[:, f], tau[:, f], sh[:, f], b[:, f], D[:, f]
    Yt, taut, sht, bt = Y[:, t], tau[:, t], sh[:, t], b[:, t]
    Dt = (-D)[:, t]
    vf, vt, thf, tht = v[:, f], v[:, t], theta[:, f], theta[:, t]

    a1 = thf - tht - Df - shf
    a2 = tht - thf - Df + shf
    a3 = tht - thf - Dt - sht

    msg = torch.abs(vf * vt * Yf / tauf * (torch.sin(a1) + torch.sin(a2))
                    + (vf / tauf.pow(2)) * Yf * torch.sin(Df)
                    + vt.pow(2) * Yf * torch.sin(Df))
    p_joule = _bus_sum(msg, t, N).sum(dim=1)
    p_global = Pd.sum(dim=1) + (v.pow(2) * Gs).sum(dim=1) + p_joule

    sPset, sPmin, sPmax = Pset.sum(dim=1), Pmin.sum(dim=1), Pmax.sum(dim=1)
    lam_lo = (p_global - sPmin) / (2 * (sPset - sPmin))
    lam_hi = (p_global - 2 * sPset + sPmax) / (2 * (sPmax - sPset))
    lam = torch.where(p_global < sPset, lam_lo, lam_hi).unsqueeze(1)
    Pg = torch.where(lam < 0.5,
                     Pmin + 2 * (Pset - Pmin) * lam,
                     2 * Pset - Pmax + 2 * (Pmax - Pset) * lam)

    q_from = -vf * vt * Yf / tauf * torch.cos(a1) + (vf / tauf).pow(2) * (Yf * torch.cos(Df) - bf / 2)
    q_to = -vt * vf * Yt / taut * torch.cos(a3) + vt.pow(2) * (Yt * torch.sin(Dt) - bt / 2)
    qg = (Qd - Bs * v.pow(2)) - _bus_sum(q_from, t, N) - _bus_sum(q_to, f, N)

    p_from = vf * vt * Yf / tauf * torch.sin(a1) + (vf / tauf).pow(2) * Yf * torch.sin(Df)
    p_to = vt * vf * Yt / taut * torch.sin(a3) + vt.pow(2) * Yt * torch.sin(Dt)
    dP = _bus_sum(Pg, gb, N) - Pd - Gs * v.pow(2) + _bus_sum(p_from, t, N) + _bus_sum(p_to, f, N)
    dQ = qg - Qd + Bs * v.pow(2) + _bus_sum(q_from, t, N) + _bus_sum(q_to, f, N)
    return dP, dQ, Pg, qg, p_global


def init_state(buses, gens, gb, latent_dim):
    """m, theta, v, dP, dQ before step 0 (ref GNS/main.py:141-152)."""
    S, N = buses.shape[:2]
    dt = buses.dtype
    m = torch.zeros(S, N, latent_dim, dtype=dt)
    theta = torch.zeros(S, N, dtype=dt)
    v = _bus_sum(gens[..., 4], gb, N)
    v = torch.where(v == 0, torch.ones_like(v), v)
    dP = _bus_sum(gens[..., 6], gb, N) - buses[..., 2] - buses[..., 4] * v.pow(2)
    dQ = _bus_sum(gens[..., 5], gb, N) - buses[..., 3] + buses[..., 5] * v.pow(2)
    return m, theta, v, dP, dQ


def gns_forward(params, buses, lines, gens, *, K, latent_dim, gamma=0.9, multiple_phi=False,
                return_trace=False):
    """Batched oracle forward.

    buses [S,N,6], lines [S,E,7], gens [S,Gn,7] (or the un-batched reference shapes);
    all grids share one topology.  Returns (v [S,N], theta [S,N], total_loss [S],
    last_loss [S]); un-batched input gives the reference's un-batched shapes.
    Computation runs in the dtype of ``buses`` with params cast to it.
    """
    single = buses.dim() == 2
    if single:
        buses, lines, gens = buses[None], lines[None], gens[None]
    dt = buses.dtype
    p = {n: w.to(dt) for n, w in params.items()}
    S, N = buses.shape[:2]
    f, t, gb = topology(lines, gens)
    if int(max(f.max(), t.max())) >= lines.shape[1]:
        raise IndexError("reference precondition max(bus)-1 < n_line violated (quirk Q1)")
    m, theta, v, dP, dQ = init_state(buses, gens, gb, latent_dim)
    non_gen = torch.ones(N, dtype=torch.bool)
    non_gen[gb] = False
    feat = lines[..., 2:]
    total = torch.zeros(S, dtype=dt)
    trace = []
    for k in range(K):
        if return_trace:
            trace.append(dict(v=v, theta=theta, m=m, dP=dP, dQ=dQ))
        x_line = torch.cat((m[:, t], feat), dim=2)
        state4 = torch.stack((v, theta, dP, dQ), dim=2)
        if multiple_phi:
            sums = {n: _bus_sum(_mlp(p, "phi_" + n, k, x_line), t, N) for n in ("v", "theta", "m")}
        else:
            msg = _mlp(p, "phi", k, x_line)                          # [S,E,1]
            col0 = _bus_sum(msg, t, N)                               # lands in latent column 0 (Q3)
            padded = torch.cat((col0, torch.zeros(S, N, latent_dim - 1, dtype=dt)), dim=2)
            sums = {"v": padded, "theta": padded, "m": padded}
        d_theta = _mlp(p, "L_theta", k, torch.cat((state4, m, sums["theta"]), dim=2)).squeeze(2)
        d_v = _mlp(p, "L_v", k, torch.cat((state4, m, sums["v"]), dim=2)).squeeze(2)
        d_m = _mlp(p, "L_m", k, torch.cat((state4, m, sums["m"]), dim=2))
        theta = theta + d_theta
        v = torch.where(non_gen, v + d_v, v)
        m = m + d_m
        dP, dQ, _, _, _ = physics(v, theta, buses, lines, gens, f, t, gb)
        total = total + gamma ** (K - k) * (dP.pow(2) + dQ.pow(2)).sum(dim=1) / N
    last = (dP.pow(2) + dQ.pow(2)).sum(dim=1) / N
    v_out = torch.where(v < 0, torch.zeros_like(v), v)
    if single:
        out = (v_out[0], theta[0], total[0], last[0])
    else:
        out = (v_out, theta, total, last)
    if return_trace:
        trace.append(dict(v=v, theta=theta, m=m, dP=dP, dQ=dQ))
        return out + (trace,)
    return out


def gns_loss_and_grads(params, buses, lines, gens, *, K, latent_dim, gamma=0.9, multiple_phi=False):
    """mean(total_loss) over the batch and its parameter gradients
    (training reduction of ref GNS/main.py:284,288).  Unused last-step nets get zeros."""
    leaves = {n: w.detach().clone().to(buses.dtype).requires_grad_(True) for n, w in params.items()}
    v, theta, total, last = gns_forward(leaves, buses, lines, gens, K=K, latent_dim=latent_dim,
                                        gamma=gamma, multiple_phi=multiple_phi)
    loss = total.mean()
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    grads = {n: (torch.zeros_like(w) if g is None else g)
             for (n, w), g in zip(leaves.items(), grads)}
    return (v.detach(), theta.detach(), total.detach(), last.detach()), grads


# ----------------------------------------------------------------------------
# CSR oracle (bit-exact reference for the plan builder)
# ----------------------------------------------------------------------------
def csr_by(key: np.ndarray, n_bus: int):
    """rowptr/ids of lines grouped by ``key`` (0-based bus per line), stable order."""
    key = np.asarray(key, dtype=np.int64)
    order = np.argsort(key, kind="stable").astype(np.int32)
    counts = np.bincount(key, minlength=n_bus)
    rowptr = np.zeros(n_bus + 1, dtype=np.int32)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, order


def degree_order(in_deg: np.ndarray):
    """Internal bus order used by the kernels: in-degree descending, stable."""
    return np.argsort(-np.asarray(in_deg, dtype=np.int64), kind="stable").astype(np.int32)
