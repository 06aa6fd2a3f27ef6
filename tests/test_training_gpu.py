"""GPU: the training driver actually trains (ref GNS/main.py:274-309), step for step like the oracle
under torch.optim.Adam on the CPU, and the evaluation metrics of the trained weights (8f-3, ref GNS/evaluate.py:89-148)
agree with the same arithmetic on the oracle's outputs."""
import numpy as np
import pytest
import torch

import opf_graph_neural_solver_b200 as pkg
from oracle import gns_oracle as orc
from oracle import newton_raphson as nr
from helpers import golden_files, load_golden

pytestmark = pytest.mark.gpu


def test_twenty_adam_steps_follow_the_oracle_and_evaluation_agrees(lib):
    g = load_golden([p for p in golden_files() if p.endswith("k4_l20_multi.npz")][0])
    K, L = g["K"], g["latent_dim"]
    buses, lines, gens = g["buses"], g["lines"], g["gens"]               # the 32 case14 samples of the reference
    model = pkg.GNS(latent_dim=L, hidden_dim=g["hidden_dim"], K=K, gamma=g["gamma"], multiple_phi=True)
    model.load_state_dict(g["params"])                                   # the reference's own seed-0 initialisation
    model = model.cuda()
    b, l, gg = buses.cuda(), lines.cuda(), gens.cuda()
    opt = pkg.train.FlatAdam(model, lr=1e-3)
    ours = []
    for _ in range(20):                                                  # ref GNS/main.py:279-291, one batch of 32
        opt.zero_grad()
        _, _, total, _ = model(b, l, gg)
        loss = total.mean()
        loss.backward()
        opt.step()
        ours.append(float(loss.detach()))
    # the oracle under torch.optim.Adam on the CPU, same start, same data
    leaves = {n: w.clone().requires_grad_(True) for n, w in g["params"].items()}
    topt = torch.optim.Adam(list(leaves.values()), lr=1e-3)
    theirs = []
    for _ in range(20):
        topt.zero_grad()
        _, _, total, _ = orc.gns_forward(leaves, buses, lines, gens, K=K, latent_dim=L, gamma=g["gamma"], multiple_phi=True)
        loss = total.mean()
        loss.backward()
        topt.step()
        theirs.append(float(loss.detach()))
    ours, theirs = np.array(ours), np.array(theirs)
    # the loss goes down: to a third in 20 steps, never up by more than 0.5 % (Adam's momentum overshoots once)
    for curve in (ours, theirs):
        assert curve[-1] < 0.4 * curve[0] and np.all(np.diff(curve) < 5e-3 * curve[:-1]), curve
    assert np.max(np.abs(ours - theirs) / theirs) < 1e-3, (ours, theirs)                 # loss curves agree

    # ---- evaluation of the trained checkpoint against a power-flow solution (restated Newton-Raphson) ----
    S = 48
    tables = pkg.data.augment(pkg.data.get_case(14)[0], S, seed=9, nominal_taps=True)
    sol = nr.newton_pf_batch(tables, range(S))
    assert all(s[2] for s in sol)
    vm = np.stack([s[0] for s in sol]).astype(np.float32)
    va_deg = np.rad2deg(np.stack([s[1] for s in sol])).astype(np.float32)
    eb, el, eg = pkg.data.pack_grids(tables["bus"], tables["branch"], tables["gen"], tables["baseMVA"])
    br = tables["branch"]
    sd = {n: w.detach().cpu() for n, w in model.state_dict().items()}
    reloaded = pkg.GNS(latent_dim=L, hidden_dim=g["hidden_dim"], K=K, gamma=g["gamma"], multiple_phi=True)
    reloaded.load_state_dict(sd)                                         # checkpoint round trip (ref GNS/evaluate.py:66)
    got, (v, theta, last) = pkg.evaluate.evaluate_model(reloaded.cuda(), eb.cuda(), el.cuda(), eg.cuda(), vm, va_deg,
                                                        br[:, :, nr.BR_X], br[:, :, nr.F_BUS], br[:, :, nr.T_BUS])
    with torch.no_grad():
        ov, oth, _, olast = orc.gns_forward(sd, eb, el, eg, K=K, latent_dim=L, gamma=g["gamma"], multiple_phi=True)
    want = pkg.evaluate.comparison_metrics(ov.numpy(), oth.numpy(), olast.numpy(), el.numpy(), vm, va_deg,
                                           br[:, :, nr.BR_X], br[:, :, nr.F_BUS], br[:, :, nr.T_BUS])
    assert np.abs(v - ov.numpy()).max() < 1e-4 and np.abs(theta - oth.numpy()).max() < 1e-4
    for k, w in want.items():
        assert got[k] == pytest.approx(w, rel=2e-3, abs=1e-5), (k, got[k], w)
