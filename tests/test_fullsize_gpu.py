"""BASELINE.json full-size batches on the GPU, checked through size-independent properties:
batch-position invariance (bit-exact), oracle agreement on sampled grids, gradient linearity over a
periodic batch, and run-to-run determinism (no contended atomics anywhere in the path: the backward kernel's
`red.global.add` only target accumulators owned by one warp, in program order)."""
import pytest
import torch

import opf_graph_neural_solver_b200 as pkg
from oracle import gns_oracle as orc
from helpers import assert_bus_close, assert_grads_close, assert_loss_close

pytestmark = pytest.mark.gpu
BLG = pkg.get_BLG()


def _periodic_batch(n_bus, period, S, seed):
    b, l, g, _ = pkg.data.make_batch(n_bus, period, seed=seed)
    rep = S // period
    return b, l, g, tuple(t.repeat(rep, 1, 1).contiguous().cuda() for t in (b, l, g))


def test_case300_inference_batch_65536(lib):
    """configs[3]: case300 K=4 inference, 65,536 grids."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    period, S = 2048, 65536
    b, l, g, (B_, L_, G_) = _periodic_batch(300, period, S, seed=21)
    with torch.no_grad():
        v, th, tot, last = model(B_, L_, G_, *BLG)
        v2 = model(B_, L_, G_, *BLG)[0]
    assert torch.equal(v, v2)                                             # deterministic
    for t in (v, th, tot, last):                                          # a grid's result does not depend on its position
        folded = t.view(S // period, period, *t.shape[1:])
        assert torch.equal(folded[0], folded[-1]) and torch.equal(folded[0], folded[S // period // 2])
    idx = torch.tensor([0, 1, 777, 2047])
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    want = orc.gns_forward(params, b[idx].double(), l[idx].double(), g[idx].double(), K=4, latent_dim=20,
                           gamma=0.9, multiple_phi=True)
    assert_bus_close(v[idx.cuda()], want[0], "v")
    assert_bus_close(th[idx.cuda()], want[1], "theta")
    assert_loss_close(tot[idx.cuda()], want[2], "total")
    assert bool(torch.isfinite(v).all() and torch.isfinite(tot).all())


def test_case118_training_batch_16384_gradient_linearity(lib):
    """configs[2]: case118 fwd+bwd, 16,384 grids: the gradient of the batch mean over a periodic batch
    equals the gradient over one period (linearity of the batch reduction, ref GNS/main.py:284)."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    period, S = 256, 16384
    b, l, g, (B_, L_, G_) = _periodic_batch(118, period, S, seed=22)
    model(B_, L_, G_, *BLG)[2].mean().backward()
    big = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    model(B_, L_, G_, *BLG)[2].mean().backward()
    for n, p in model.named_parameters():
        assert torch.equal(p.grad, big[n]), n                             # run-to-run determinism of the gradient
    model.zero_grad(set_to_none=True)
    model(B_[:period], L_[:period], G_[:period], *BLG)[2].mean().backward()
    small = {n: p.grad.cpu() for n, p in model.named_parameters()}
    assert_grads_close(big, small, "periodic batch")
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    _, want = orc.gns_loss_and_grads(params, b[:32].double(), l[:32].double(), g[:32].double(), K=4, latent_dim=20,
                                     gamma=0.9, multiple_phi=True)
    model.zero_grad(set_to_none=True)
    model(B_[:32], L_[:32], G_[:32], *BLG)[2].mean().backward()
    assert_grads_close({n: p.grad for n, p in model.named_parameters()}, want, "oracle on 32 grids")


def test_case30_training_batch_4096(lib):
    """configs[1]: case30 training step, batch 4096, vs the oracle's autograd on the whole batch."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    b, l, g, _ = pkg.data.make_batch(30, 4096, seed=23)
    out = model(b.cuda(), l.cuda(), g.cuda(), *BLG)
    out[2].mean().backward()
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (ov, oth, otot, olast), want = orc.gns_loss_and_grads(params, b, l, g, K=4, latent_dim=20, gamma=0.9, multiple_phi=True)
    assert_bus_close(out[0], ov, "v")
    assert_loss_close(out[2], otot, "total")
    assert_grads_close({n: p.grad for n, p in model.named_parameters()}, want, "case30 batch 4096")
