"""Newton-Raphson AC power flow, restated from scratch.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference compares the GNS against pypower's ``runpf`` (``ppoption(PF_ALG=1)``,
ref GNS/evaluate.py:25-26,31-40).  pypower 5.1.16 is not installed and cannot be installed
offline, so this file restates the published algorithm pypower implements (MATPOWER's
``newtonpf``: polar-form full Newton, flat start from the generator set-points, tolerance 1e-8 on
the infinity norm of the mismatch, at most 10 iterations) on the same case tables
(``bus`` / ``branch`` / ``gen`` in pypower column order).  It is labelled "restated NR, not
pypower" wherever its timing is reported (``bench.py --impl reference``).  Nothing in the
product path imports it.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve

# pypower column numbers
BUS_I, BUS_TYPE, PD, QD, GS, BS, VM, VA = 0, 1, 2, 3, 4, 5, 7, 8
F_BUS, T_BUS, BR_R, BR_X, BR_B, TAP, SHIFT = 0, 1, 2, 3, 4, 8, 9
GEN_BUS, PG, QG, VG = 0, 1, 2, 5
PQ, PV, REF = 1, 2, 3


def make_ybus(base_mva, bus, branch):
    nb, nl = bus.shape[0], branch.shape[0]
    ys = 1.0 / (branch[:, BR_R] + 1j * branch[:, BR_X])
    bc = branch[:, BR_B]
    tap = np.where(branch[:, TAP] == 0, 1.0, branch[:, TAP]) * np.exp(1j * np.pi / 180.0 * branch[:, SHIFT])
    ytt = ys + 1j * bc / 2
    yff = ytt / (tap * np.conj(tap))
    yft = -ys / np.conj(tap)
    ytf = -ys / tap
    ysh = (bus[:, GS] + 1j * bus[:, BS]) / base_mva
    f = branch[:, F_BUS].astype(int) - 1
    t = branch[:, T_BUS].astype(int) - 1
    rows = np.concatenate([f, f, t, t, np.arange(nb)])
    cols = np.concatenate([f, t, f, t, np.arange(nb)])
    vals = np.concatenate([yff, yft, ytf, ytt, ysh])
    return sp.csr_matrix((vals, (rows, cols)), shape=(nb, nb))


def _ds_dv(ybus, v):
    ibus = ybus @ v
    diag_v = sp.diags(v)
    diag_i = sp.diags(ibus)
    diag_vn = sp.diags(v / np.abs(v))
    ds_dvm = diag_v @ np.conj(ybus @ diag_vn) + np.conj(diag_i) @ diag_vn
    ds_dva = 1j * diag_v @ np.conj(diag_i - ybus @ diag_v)
    return ds_dvm, ds_dva


def newton_pf(case, tol=1e-8, max_it=10):
    """One power flow.  Returns (Vm, Va [rad], converged, iterations)."""
    base, bus, branch, gen = float(case["baseMVA"]), case["bus"], case["branch"], case["gen"]
    nb = bus.shape[0]
    gbus = gen[:, GEN_BUS].astype(int) - 1
    btype = bus[:, BUS_TYPE].astype(int).copy()
    if not (btype == REF).any():
        btype[gbus[0]] = REF
    ref = np.flatnonzero(btype == REF)
    pv = np.flatnonzero(btype == PV)
    pq = np.flatnonzero(btype == PQ)
    ybus = make_ybus(base, bus, branch)
    sbus = -(bus[:, PD] + 1j * bus[:, QD]) / base
    np.add.at(sbus, gbus, (gen[:, PG] + 1j * gen[:, QG]) / base)
    vm = np.ones(nb)
    vm[gbus] = gen[:, VG]
    v = vm.astype(complex)                      # flat start, generator voltage set-points
    pvpq = np.concatenate([pv, pq])
    npv, npq = len(pv), len(pq)

    def mismatch(v):
        mis = v * np.conj(ybus @ v) - sbus
        return np.concatenate([mis[pvpq].real, mis[pq].imag])

    f_vec = mismatch(v)
    converged = np.max(np.abs(f_vec)) < tol
    it = 0
    va, vm = np.angle(v), np.abs(v)
    while not converged and it < max_it:
        it += 1
        ds_dvm, ds_dva = _ds_dv(ybus, v)
        j11 = ds_dva[pvpq][:, pvpq].real
        j12 = ds_dvm[pvpq][:, pq].real
        j21 = ds_dva[pq][:, pvpq].imag
        j22 = ds_dvm[pq][:, pq].imag
        jac = sp.bmat([[j11, j12], [j21, j22]], format="csc")
        dx = -spsolve(jac, f_vec)
        va[pvpq] += dx[:npv + npq]
        vm[pq] += dx[npv + npq:]
        v = vm * np.exp(1j * va)
        f_vec = mismatch(v)
        converged = np.max(np.abs(f_vec)) < tol
    return np.abs(v), np.angle(v), bool(converged), it


def newton_pf_batch(tables, indices):
    """Run NR on the samples `indices` of batched tables (as returned by data.augment)."""
    out = []
    for i in indices:
        case = {"baseMVA": tables["baseMVA"], "bus": tables["bus"][i], "branch": tables["branch"][i],
                "gen": tables["gen"][i]}
        out.append(newton_pf(case))
    return out
