"""CPU: the oracle restatement against the golden vectors produced by the live reference."""
import numpy as np
import pytest
import torch

from oracle import gns_oracle as orc
from helpers import golden_files, load_golden


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("ref_case14_")[-1][:-4])
def test_oracle_matches_reference_outputs_and_grads(path):
    g = load_golden(path)
    (v, th, tot, last), grads = orc.gns_loss_and_grads(
        g["params"], g["buses"], g["lines"], g["gens"], K=g["K"], latent_dim=g["latent_dim"],
        gamma=g["gamma"], multiple_phi=g["multiple_phi"])
    # measured floor between the per-sample reference and the batched restatement is ~3e-7
    assert (v - g["v"]).abs().max() < 2e-6
    assert (th - g["theta"]).abs().max() < 2e-6
    assert ((tot - g["total_loss"]).abs() / g["total_loss"].abs()).max() < 1e-5
    assert ((last - g["last_loss"]).abs() / g["last_loss"].abs()).max() < 1e-5
    gmax = max(float(w.abs().max()) for w in g["grads"].values())
    for n, w in g["grads"].items():
        assert (grads[n] - w).abs().max() <= 1e-5 * gmax + 1e-7, n
    for n in g["none_grads"]:          # unused last-step nets: None in the reference, zeros here
        assert float(grads[n].abs().max()) == 0.0


def test_reference_init_is_reproduced_bit_for_bit():
    g = load_golden(golden_files()[0])
    p = orc.init_params(g["latent_dim"], g["hidden_dim"], g["K"], g["multiple_phi"], seed=0)
    assert list(p) == list(g["params"])
    for n in p:
        assert torch.equal(p[n], g["params"][n]), n


def test_known_answer_case14_sample1():
    """SURVEY.md 8c known-answer: seed-0 K=4 L=20 H=10 multi-phi on case14 sample 1."""
    g = load_golden([p for p in golden_files() if p.endswith("k4_l20_multi.npz")][0])
    v, th, tot, last = orc.gns_forward(g["params"], g["buses"][0], g["lines"][0], g["gens"][0], K=4,
                                       latent_dim=20, gamma=0.9, multiple_phi=True)
    assert abs(float(tot) - 1.218155) < 2e-6 and abs(float(last) - 0.220033) < 2e-6
    assert abs(float(v[3]) - 0.835368) < 2e-6 and abs(float(th[0]) + 0.659223) < 2e-6


def test_float64_oracle_agrees_with_float32():
    g = load_golden([p for p in golden_files() if p.endswith("k8_l64_multi.npz")][0])
    out32 = orc.gns_forward(g["params"], g["buses"], g["lines"], g["gens"], K=g["K"], latent_dim=g["latent_dim"],
                            gamma=g["gamma"], multiple_phi=True)
    out64 = orc.gns_forward(g["params"], g["buses"].double(), g["lines"].double(), g["gens"].double(), K=g["K"],
                            latent_dim=g["latent_dim"], gamma=g["gamma"], multiple_phi=True)
    assert (out32[0].double() - out64[0]).abs().max() < 5e-6
    assert ((out32[2].double() - out64[2]).abs() / out64[2]).max() < 1e-5


def test_dq_is_cancellation_noise():
    """Quirk Q4: dQ == 0 analytically, so the oracle's dQ is float32 noise."""
    g = load_golden(golden_files()[0])
    f, t, gb = orc.topology(g["lines"], g["gens"])
    m, theta, v, dP, dQ = orc.init_state(g["buses"], g["gens"], gb, g["latent_dim"])
    dP, dQ, *_ = orc.physics(v, theta, g["buses"], g["lines"], g["gens"], f, t, gb)
    assert float(dQ.abs().max()) < 1e-5 and float(dP.abs().max()) > 1e-2


def test_newton_raphson_restatement_reproduces_ieee14_solution():
    """Known answer of the IEEE 14-bus case (MATPOWER/pypower case14 solution): the restated NR is
    only a reported CPU baseline (ref GNS/evaluate.py:31-40), but it should still be a power flow."""
    import opf_graph_neural_solver_b200 as pkg
    from oracle import newton_raphson as nr
    vm, va, ok, it = nr.newton_pf(pkg.data.case14())
    assert ok and it <= 5
    deg = np.degrees(va)
    assert abs(vm[13] - 1.0355) < 2e-4 and abs(deg[13] + 16.03) < 0.01      # bus 14
    assert abs(deg[1] + 4.98) < 0.01 and abs(deg[2] + 12.73) < 0.01          # buses 2, 3
    assert abs(vm[3] - 1.0177) < 2e-4 and abs(vm[8] - 1.0559) < 2e-4         # buses 4, 9
