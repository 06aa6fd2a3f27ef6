"""Evaluation metrics: the batched numpy code against a per-sample restatement of the
reference script's own lines (ref GNS/evaluate.py:15-18, 73-131), and an end-to-end run against the
restated Newton-Raphson (oracle/, test infrastructure) on the reference's IEEE-14 case."""
import numpy as np
import pytest
import torch

import opf_graph_neural_solver_b200 as pkg


def _reference_script_metrics(gns_v, gns_theta, last_losses, lines, nr_v, nr_theta_deg, nr_x, nr_f, nr_t):
    """The reference's evaluate.py, sample loop and all (variable names kept)."""
    def active_line_flow(V, theta, x, src, dst):                    # ref :15-18
        src = src.astype(int) - 1
        dst = dst.astype(int) - 1
        return 1 / x * (V[src] * V[dst] * np.sin(theta[src] - theta[dst]))
    S, E = gns_v.shape[0], lines.shape[1]
    NR_active_line_flow = np.zeros((S, E), dtype=np.float32)
    GNS_active_line_flow = np.zeros((S, E), dtype=np.float32)
    for i in range(S):
        NR_active_line_flow[i] = active_line_flow(nr_v[i], nr_theta_deg[i], nr_x[i], nr_f[i], nr_t[i])       # ref :40
        GNS_active_line_flow[i] = active_line_flow(gns_v[i], gns_theta[i], lines[i, :, 3], lines[i, :, 0], lines[i, :, 1])  # ref :87
    NR_theta_out = np.deg2rad(nr_theta_deg)                         # ref :99
    theta_diff_gns_nr = np.abs(gns_theta - NR_theta_out)
    v_diff_gns_nr = np.abs(gns_v - nr_v)
    alf_diff_gns_nr = NR_active_line_flow - GNS_active_line_flow
    percentage_diff_alf_gns_nr = np.abs(alf_diff_gns_nr / NR_active_line_flow) * 100
    percentage_diff_alf_gns_nr = np.sort(percentage_diff_alf_gns_nr, axis=None)[:int(percentage_diff_alf_gns_nr.size / 2)]
    return {
        "theta_diff_mean": np.mean(theta_diff_gns_nr), "theta_diff_std": np.std(theta_diff_gns_nr),
        "v_diff_mean": np.mean(v_diff_gns_nr), "v_diff_std": np.std(v_diff_gns_nr),
        "last_loss_mean": np.mean(last_losses), "last_loss_std": np.std(last_losses),
        "line_flow_pct_p20": np.percentile(percentage_diff_alf_gns_nr, 20),
        "line_flow_pct_median": np.median(percentage_diff_alf_gns_nr),
        "line_flow_pct_p80": np.percentile(percentage_diff_alf_gns_nr, 80),
    }


def test_metrics_match_the_reference_script_arithmetic():
    rng = np.random.default_rng(0)
    S, N, E = 17, 14, 20
    f = rng.integers(1, N + 1, size=E).astype(np.float32)
    t = ((f + rng.integers(0, N - 1, size=E)) % N + 1).astype(np.float32)
    lines = np.zeros((S, E, 7), np.float32)
    lines[:, :, 0], lines[:, :, 1] = f, t
    lines[:, :, 3] = rng.uniform(0.04, 0.6, size=(S, E))
    gns_v = rng.uniform(0.9, 1.1, size=(S, N)).astype(np.float32)
    gns_theta = rng.uniform(-0.3, 0.3, size=(S, N)).astype(np.float32)
    nr_v = (gns_v + rng.normal(0, 0.01, size=(S, N))).astype(np.float32)
    nr_theta_deg = np.rad2deg(gns_theta + rng.normal(0, 0.01, size=(S, N))).astype(np.float32)
    last = rng.uniform(0, 1, size=S).astype(np.float32)
    nr_x = lines[:, :, 3] * rng.uniform(0.99, 1.01, size=(S, E)).astype(np.float32)
    nr_f, nr_t = np.broadcast_to(f, (S, E)), np.broadcast_to(t, (S, E))
    want = _reference_script_metrics(gns_v, gns_theta, last, lines, nr_v, nr_theta_deg, nr_x, nr_f, nr_t)
    got = pkg.evaluate.comparison_metrics(gns_v, gns_theta, last, lines, nr_v, nr_theta_deg, nr_x, nr_f, nr_t)
    for k, w in want.items():
        assert got[k] == pytest.approx(float(w), rel=1e-6, abs=1e-9), k
    # without the degree quirk the Newton-Raphson flows use radians and the percentiles shrink
    fixed = pkg.evaluate.comparison_metrics(gns_v, gns_theta, last, lines, nr_v, nr_theta_deg, nr_x, nr_f, nr_t,
                                            reference_degree_quirk=False)
    assert fixed["line_flow_pct_median"] < got["line_flow_pct_median"]
    # single-sample form of active_line_flow == row of the batched form
    one = pkg.evaluate.active_line_flow(gns_v[3], gns_theta[3], lines[3, :, 3], f, t)
    many = pkg.evaluate.active_line_flow(gns_v, gns_theta, lines[:, :, 3], f, t)
    assert np.array_equal(one, many[3])


@pytest.mark.gpu
def test_evaluate_model_against_restated_newton_raphson(lib):
    from oracle import newton_raphson as nr
    S = 64
    tables = pkg.data.augment(pkg.data.get_case(14)[0], S, seed=5)
    sol = nr.newton_pf_batch(tables, range(S))
    assert all(s[2] for s in sol), "restated Newton-Raphson must converge on the IEEE-14 samples"
    vm = np.stack([s[0] for s in sol]).astype(np.float32)
    va_deg = np.rad2deg(np.stack([s[1] for s in sol])).astype(np.float32)
    buses, lines, gens = pkg.data.pack_grids(tables["bus"], tables["branch"], tables["gen"], tables["baseMVA"])
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    br = tables["branch"]
    m, (v, theta, last) = pkg.evaluate.evaluate_model(model, buses.cuda(), lines.cuda(), gens.cuda(), vm, va_deg,
                                                       br[:, :, nr.BR_X], br[:, :, nr.F_BUS], br[:, :, nr.T_BUS])
    assert v.shape == (S, 14) and all(np.isfinite(list(m.values())))
    assert m["last_loss_mean"] == pytest.approx(float(last.mean()), rel=1e-6)
    # untrained weights: the GNS is far from the power-flow solution, but voltages stay within a few tenths
    assert 0.0 < m["v_diff_mean"] < 1.0
