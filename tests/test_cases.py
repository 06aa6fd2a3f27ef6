"""CPU: the case tables and synthetic topologies behind the measurements."""
import numpy as np
import pytest

import opf_graph_neural_solver_b200 as pkg
from oracle import newton_raphson as nr


def test_case30_table_reproduces_the_published_power_flow():
    """The hand-entered IEEE 30-bus table (pypower's case30, ref GNS/augment_grids.py:1,8) solved by the restated
    Newton-Raphson gives the published runpf answer: losses 2.444 MW, slack 25.97 MW, min |V| 0.961 at bus 8,
    angle extremes -3.96 deg at bus 19 and +1.48 deg at bus 13."""
    case = pkg.data.case30()
    assert case["bus"].shape == (30, 13) and case["branch"].shape == (41, 13) and case["gen"].shape == (6, 21)
    assert abs(case["bus"][:, 2].sum() - 189.2) < 1e-9
    vm, va, conv, it = nr.newton_pf(case)
    assert conv and it <= 4
    assert vm.argmin() == 7 and abs(vm.min() - 0.9606) < 2e-4
    deg = np.degrees(va)
    assert deg.argmin() == 18 and abs(deg.min() + 3.958) < 2e-3
    assert deg.argmax() == 12 and abs(deg.max() - 1.476) < 2e-3
    v = vm * np.exp(1j * va)
    s = v * np.conj(nr.make_ybus(100.0, case["bus"], case["branch"]) @ v)
    assert abs(s.real.sum() * 100 - 2.444) < 2e-3 and abs(s[0].real * 100 - 25.974) < 2e-3


def test_case30_packs_like_the_reference_sizes():
    b, l, g, label = pkg.data.make_batch(30, 5, seed=3)
    assert label == "IEEE case30" and b.shape == (5, 30, 6) and l.shape == (5, 41, 7) and g.shape == (5, 6, 7)
    assert set(g[0, :, 0].tolist()) == {1.0, 2.0, 22.0, 27.0, 23.0, 13.0}


@pytest.mark.parametrize("n_bus", [118, 300])
def test_synthetic_topologies_are_shaped_like_the_ieee_systems_and_solvable(n_bus):
    case, label = pkg.data.get_case(n_bus)
    prof = pkg.data.degree_profile(case)
    want_par, want_hub = pkg.data._IEEE_SHAPE[n_bus]
    assert prof["max_degree"] == want_hub and prof["parallel_lines"] == want_par, prof
    assert abs(prof["mean_degree"] - 2 * pkg.data.IEEE_SIZES[n_bus][0] / n_bus) < 1e-9
    tables = pkg.data.augment(case, 8, seed=5, nominal_taps=True)
    res = nr.newton_pf_batch(tables, range(8))
    assert all(r[2] for r in res), [r[3] for r in res]         # a flat-start power flow exists and converges


def test_nominal_tap_option_only_changes_taps_and_shifts():
    case, _ = pkg.data.get_case(118)
    a = pkg.data.augment(case, 4, seed=2)
    b = pkg.data.augment(case, 4, seed=2, nominal_taps=True)
    assert np.array_equal(a["bus"], b["bus"]) and np.array_equal(a["gen"], b["gen"])
    same = [c for c in range(13) if c not in (8, 9)]
    assert np.array_equal(a["branch"][:, :, same], b["branch"][:, :, same])
    assert np.array_equal(b["branch"][:, :, 8], np.repeat(case["branch"][None, :, 8], 4, 0))
