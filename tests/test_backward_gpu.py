"""GPU parity of the backward kernel (hand-written BPTT) against the reference's autograd
gradients (golden vectors) and against the oracle's autograd on larger cases."""
import pytest
import torch

import opf_graph_neural_solver_b200 as pkg
from oracle import gns_oracle as orc
from helpers import assert_grads_close, assert_loss_close, golden_files, load_golden

pytestmark = pytest.mark.gpu
BLG = pkg.get_BLG()


def _model_from(params, latent_dim, hidden_dim, K, gamma, multiple_phi):
    m = pkg.GNS(latent_dim=latent_dim, hidden_dim=hidden_dim, K=K, gamma=gamma, multiple_phi=multiple_phi)
    m.load_state_dict(params)
    return m.cuda()


def _grads(model):
    return {n: (p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}


def per_tensor_report(got, want):
    gmax = max(float(w.abs().max()) for w in want.values())
    rows = []
    for n, w in want.items():
        err = float((got[n].detach().cpu().double() - w.double()).abs().max())
        rows.append((err / gmax, n, err, float(w.abs().max())))
    rows.sort(reverse=True)
    return "\n".join(f"{r[1]:32s} err {r[2]:.3e}  max|g| {r[3]:.3e}" for r in rows[:12])


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("ref_case14_")[-1][:-4])
def test_gradients_match_reference_autograd(lib, path):
    g = load_golden(path)
    model = _model_from(g["params"], g["latent_dim"], g["hidden_dim"], g["K"], g["gamma"], g["multiple_phi"])
    v, th, tot, last = model(g["buses"].cuda(), g["lines"].cuda(), g["gens"].cuda(), *BLG)
    tot.mean().backward()                       # training reduction, ref GNS/main.py:284-288
    got = _grads(model)
    try:
        assert_grads_close(got, g["grads"], g["name"])
    except AssertionError:
        print(per_tensor_report(got, g["grads"]))
        raise
    for n in g["none_grads"]:                   # unused last-step nets: exactly zero here (None in the reference)
        assert float(got[n].abs().max()) == 0.0, n


@pytest.mark.parametrize("n_bus,S,multi", [(14, 5, True), (30, 33, True), (118, 9, True), (300, 3, True),
                                            (30, 17, False), (300, 2, False)])
def test_gradients_vs_oracle_autograd(lib, n_bus, S, multi):
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=multi).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(n_bus, S, seed=7)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (_, _, otot, _), want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=4,
                                                   latent_dim=20, gamma=0.9, multiple_phi=multi)
    out = model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)
    out[2].mean().backward()
    assert_loss_close(out[2], otot, "total")
    got = _grads(model)
    try:
        assert_grads_close(got, want, f"case{n_bus}")
    except AssertionError:
        print(per_tensor_report(got, want))
        raise


@pytest.mark.parametrize("n_bus,S,multi,L,K", [(300, 3, True, 20, 4), (300, 2, False, 20, 4), (118, 9, True, 20, 4),
                                               (300, 150, True, 10, 3)])
def test_warp_specialised_backward_kernel_matches_oracle(lib, monkeypatch, n_bus, S, multi, L, K):
    """GNS_BWD2=1: the second, independent implementation of the adjoint (producer / consumer warps, per-grid block
    checkpoints, slope bit masks; csrc/gns_backward2.cuh) gives the same gradients, bit-reproducibly."""
    monkeypatch.setenv("GNS_BWD2", "1")
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=multi).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(n_bus, S, seed=7)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (_, _, otot, _), want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=K,
                                                   latent_dim=L, gamma=0.9, multiple_phi=multi)
    b, l, g = buses.cuda(), lines.cuda(), gens.cuda()
    info = model.plan_for(l, g, n_bus).launch_info(S, K, L, 10, multi, backward=True)
    assert info["grids_per_cta"] == 1 and info["vector_width"] == 2, info      # the warp-specialised geometry is in use
    runs = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        out = model(b, l, g, *BLG)
        out[2].mean().backward()
        runs.append({n: p.grad.detach().clone() for n, p in model.named_parameters()})
    assert_loss_close(out[2], otot, "total")
    try:
        assert_grads_close(runs[0], want, f"case{n_bus} bwd2")
    except AssertionError:
        print(per_tensor_report(runs[0], want))
        raise
    assert all(torch.equal(runs[0][n], runs[1][n]) for n in runs[0])


@pytest.mark.parametrize("tiles", ["1", "2"])
@pytest.mark.parametrize("n_bus,S,L,K", [(300, 3, 20, 4), (118, 9, 20, 4), (300, 150, 10, 3)])
def test_fragment_space_backward_kernel_matches_oracle(lib, monkeypatch, n_bus, S, L, K, tiles):
    """GNS_BWD3=1: the third implementation of the adjoint (every layer of the MLP adjoint as 3xTF32 mma.sync products on
    16-item tiles, csrc/gns_backward3.cuh), in both launch geometries (GNS_BWD3_TILES=1: one bus tile per warp, 2: two),
    gives the same gradients, bit-reproducibly."""
    monkeypatch.setenv("GNS_BWD3", "1")
    monkeypatch.setenv("GNS_BWD3_TILES", tiles)
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(n_bus, S, seed=7)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (_, _, otot, _), want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=K,
                                                   latent_dim=L, gamma=0.9, multiple_phi=True)
    b, l, g = buses.cuda(), lines.cuda(), gens.cuda()
    info = model.plan_for(l, g, n_bus).launch_info(S, K, L, 10, True, backward=True)
    assert info["grids_per_cta"] == 1 and info["vector_width"] == 16 * int(tiles), info   # the fragment-space geometry is in use
    runs = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        out = model(b, l, g, *BLG)
        out[2].mean().backward()
        runs.append({n: p.grad.detach().clone() for n, p in model.named_parameters()})
    assert_loss_close(out[2], otot, "total")
    try:
        assert_grads_close(runs[0], want, f"case{n_bus} bwd3")
    except AssertionError:
        print(per_tensor_report(runs[0], want))
        raise
    assert all(torch.equal(runs[0][n], runs[1][n]) for n in runs[0])


def test_deepcopied_model_still_receives_gradients(lib):
    """copy.deepcopy drops the flat leaf's hook: the copy must rebuild it instead of training nothing."""
    import copy
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=10, hidden_dim=10, K=2, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(14, 4, seed=1)
    b, l, g = buses.cuda(), lines.cuda(), gens.cuda()
    model(b, l, g, *BLG)[2].mean().backward()
    want = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    twin = copy.deepcopy(model)
    twin.zero_grad(set_to_none=True)
    twin(b, l, g, *BLG)[2].mean().backward()
    for n, p in twin.named_parameters():
        assert p.grad is not None, n
        assert torch.equal(p.grad, want[n]), n


def test_flat_adam_with_frozen_parameters_matches_torch_adam(lib):
    """with frozen parameters the gradient alias is shorter than the flat buffer; FlatAdam must not pair
    gradients with the wrong parameters, and must leave the frozen ones untouched like torch.optim.Adam."""
    torch.manual_seed(1)
    model = pkg.GNS(latent_dim=10, hidden_dim=10, K=2, gamma=0.9, multiple_phi=True).cuda()
    ref = pkg.GNS(latent_dim=10, hidden_dim=10, K=2, gamma=0.9, multiple_phi=True).cuda()
    ref.load_state_dict(model.state_dict())
    for m in (model, ref):
        for n, p in m.named_parameters():
            if n.startswith("phi_v.0") or n.startswith("L_m.1.linear4"):
                p.requires_grad_(False)
    buses, lines, gens, _ = pkg.data.make_batch(14, 8, seed=2)
    b, l, g = buses.cuda(), lines.cuda(), gens.cuda()
    ref.per_parameter_autograd = True
    model.per_parameter_autograd = True
    opt = pkg.train.FlatAdam(model, lr=1e-2)
    topt = torch.optim.Adam([p for p in ref.parameters() if p.requires_grad], lr=1e-2)
    for _ in range(3):
        opt.zero_grad(); topt.zero_grad(set_to_none=True)
        model(b, l, g, *BLG)[2].mean().backward()
        ref(b, l, g, *BLG)[2].mean().backward()
        opt.step(); topt.step()
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        diff = (p.detach() - q.detach()).abs()
        if not q.requires_grad:                       # frozen: bit-identical to the reference copy (never stepped)
            assert float(diff.max()) == 0.0, n
            continue
        # Adam turns rounding-noise gradients into +-lr steps of arbitrary sign: compare those loosely
        solid = q.grad.abs() > 1e-4
        assert float(diff[solid].max() if solid.any() else 0.0) < 5e-5, n
        assert float(diff.max()) < 4 * 3 * 1e-2, n


def test_stress_config_k8_l64_gradients_case300(lib):
    """BASELINE configs[4]: case300, K=8, latent 64 (m / adj m rows live in the global scratch)."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=64, hidden_dim=10, K=8, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(300, 3, seed=3)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (_, _, otot, _), want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=8,
                                                   latent_dim=64, gamma=0.9, multiple_phi=True)
    out = model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)
    out[2].mean().backward()
    assert_loss_close(out[2], otot, "total")
    got = _grads(model)
    try:
        assert_grads_close(got, want, "case300 K8 L64")
    except AssertionError:
        print(per_tensor_report(got, want))
        raise


def test_gradients_of_an_arbitrary_loss_on_all_four_outputs(lib):
    """v, theta, total_loss, last_loss all carry gradients (supervised / mixed losses)."""
    torch.manual_seed(3)
    model = pkg.GNS(latent_dim=10, hidden_dim=10, K=3, gamma=0.8, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(30, 11, seed=2)
    w = [torch.randn(11), torch.randn(11), torch.randn(11, 30), torch.randn(11, 30)]
    leaves = {n: p.detach().cpu().double().requires_grad_(True) for n, p in model.named_parameters()}
    ov, oth, otot, olast = orc.gns_forward(leaves, buses.double(), lines.double(), gens.double(), K=3,
                                           latent_dim=10, gamma=0.8, multiple_phi=True)
    oloss = (otot * w[0]).sum() + (olast * w[1]).sum() + (ov * w[2]).sum() + (oth * w[3]).sum()
    og = torch.autograd.grad(oloss, list(leaves.values()), allow_unused=True)
    want = {n: (torch.zeros_like(p) if gr is None else gr) for (n, p), gr in zip(leaves.items(), og)}
    v, th, tot, last = model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)
    loss = (tot * w[0].cuda()).sum() + (last * w[1].cuda()).sum() + (v * w[2].cuda()).sum() + (th * w[3].cuda()).sum()
    loss.backward()
    got = _grads(model)
    try:
        assert_grads_close(got, want, "mixed loss")
    except AssertionError:
        print(per_tensor_report(got, want))
        raise


def test_reference_style_per_sample_loop_accumulates_like_a_batch(lib):
    """ref GNS/main.py:279-288: per-sample forward, mean of losses, one backward."""
    g = load_golden([p for p in golden_files() if p.endswith("k4_l20_multi.npz")][0])
    model = _model_from(g["params"], 20, 10, 4, 0.9, True)
    losses = []
    for i in range(8):
        _, _, loss, _ = model(buses=g["buses"][i], lines=g["lines"][i], generators=g["gens"][i],
                              B=BLG[0], L=BLG[1], G=BLG[2])
        losses.append(loss)
    torch.stack(losses).mean().backward()
    loop = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad()
    out = model(g["buses"][:8].cuda(), g["lines"][:8].cuda(), g["gens"][:8].cuda(), *BLG)
    out[2].mean().backward()
    assert_grads_close(loop, {n: p.grad.cpu() for n, p in model.named_parameters()}, "loop vs batch")


def test_adam_step_matches_torch(lib):
    torch.manual_seed(0)
    n = 14768
    p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 4):
        ref.grad = g.clone()
        opt.step()
        rc = lib.gns_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8,
                               step, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        g = g * 0.5 + 0.1
    torch.cuda.synchronize()
    assert (p - ref.detach()).abs().max() < 1e-6


def test_flat_adam_training_follows_torch_adam(lib):
    """ref GNS/main.py:243,284-291: three Adam steps on the batch-mean loss, fused flat Adam vs torch.optim.Adam."""
    buses, lines, gens, _ = pkg.data.make_batch(14, 64, seed=11)
    b, l, g = buses.cuda(), lines.cuda(), gens.cuda()
    torch.manual_seed(0)
    m1 = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    torch.manual_seed(0)
    m2 = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    o1 = pkg.train.FlatAdam(m1, lr=1e-3)
    o2 = torch.optim.Adam(m2.parameters(), lr=1e-3)
    for _ in range(3):
        o1.zero_grad(); m1(b, l, g)[2].mean().backward(); o1.step()
        o2.zero_grad(); m2(b, l, g)[2].mean().backward(); o2.step()
    # Adam turns a rounding-noise gradient (e.g. L_theta.{K-1}.linear4.bias, analytically zero: uniform
    # angle shift is a gauge symmetry) into +-lr steps of arbitrary sign, so elements whose gradient
    # is noise are compared loosely and everything else tightly.
    for (n, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
        solid = p2.grad.abs() > 1e-4
        diff = (p1.detach() - p2.detach()).abs()
        assert float(diff[solid].max() if solid.any() else 0.0) < 5e-6, n
        assert float(diff.max()) < 4 * 3 * 1e-3, n
    hist = pkg.train.fit(m1, b, l, g, epochs=3, batch_size=32, log=lambda *_: None)
    assert len(hist) == 3 and all(h == h for h in hist)
    assert pkg.train.checkpoint_name(14, m1) == "best_model_c14_K4_L20_H10_True_optimAdam.pth"


def test_large_grid_gradients_wide_cta_variant(lib):
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=10, hidden_dim=10, K=2, gamma=0.9, multiple_phi=True).cuda()
    case = pkg.data.synthetic_case(400, 520, 50, seed=2)     # > 384 slots: 1024-thread variant, still fits shared memory
    aug = pkg.data.augment(case, 3, seed=1)
    buses, lines, gens = pkg.data.pack_grids(aug["bus"], aug["branch"], aug["gen"], aug["baseMVA"])
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (_, _, otot, _), want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=2,
                                                   latent_dim=10, gamma=0.9, multiple_phi=True)
    out = model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)
    out[2].mean().backward()
    assert_loss_close(out[2], otot, "total")
    assert_grads_close(_grads(model), want, "400-bus grid")
    assert model._last_plan.launch_info(3, 2, 10, 10, True, backward=True)["tmax"] == 1024


def test_backward_refuses_grids_that_do_not_fit_shared_memory(lib):
    """No silent fallback: a 700-bus grid trains nowhere else, so the call must raise."""
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=2, multiple_phi=True).cuda()
    case = pkg.data.synthetic_case(700, 900, 80, seed=2)
    aug = pkg.data.augment(case, 2, seed=1)
    buses, lines, gens = pkg.data.pack_grids(aug["bus"], aug["branch"], aug["gen"], aug["baseMVA"])
    with pytest.raises(RuntimeError, match="no launch geometry fits"):
        model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)


def test_flat_leaf_and_per_parameter_autograd_agree(lib):
    """Default gradient delivery (one flat autograd leaf + hook) vs every Parameter in the graph."""
    g = load_golden([p for p in golden_files() if p.endswith("k4_l20_multi.npz")][0])
    b, l, ge = g["buses"][:6].cuda(), g["lines"][:6].cuda(), g["gens"][:6].cuda()
    m1 = _model_from(g["params"], 20, 10, 4, 0.9, True)
    m2 = _model_from(g["params"], 20, 10, 4, 0.9, True)
    m2.per_parameter_autograd = True
    for m in (m1, m2):
        m(b, l, ge, *BLG)[2].mean().backward()
        m(b[:3], l[:3], ge[:3], *BLG)[2].sum().backward()          # second backward accumulates
    for (n, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
        assert torch.allclose(p1.grad, p2.grad, rtol=1e-6, atol=1e-7), n
    flat = pkg.parallel.flat_gradient(list(m1.parameters()))
    assert flat is not None and flat.numel() == 14768               # one buffer behind all .grad tensors
    m1.zero_grad(set_to_none=True)
    m1(b, l, ge, *BLG)[2].mean().backward()                           # fresh after zero_grad: overwrite, not add
    m2.zero_grad(set_to_none=True)
    m2(b, l, ge, *BLG)[2].mean().backward()
    for p1, p2 in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(p1.grad, p2.grad, rtol=1e-6, atol=1e-7)
    grads = torch.autograd.grad(m2(b, l, ge, *BLG)[2].mean(), list(m2.parameters()), allow_unused=True)
    assert grads[0] is not None                                       # parameters are graph inputs in this mode


def _case_grads(n_bus, S, seed=11):
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(n_bus, S, seed=seed)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    _, want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=4, latent_dim=20,
                                     gamma=0.9, multiple_phi=True)
    model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)[2].mean().backward()
    return _grads(model), want


def test_tensor_core_weight_gradients_keep_fp32_accuracy(lib):
    """The weight-gradient tiles run on mma.sync TF32 with a 3-term operand split.  Plain TF32 would sit at
    ~1e-3 of max|g|; the split has to stay two orders of magnitude below the parity tolerance."""
    got, want = _case_grads(118, 48)
    worst, gmax = assert_grads_close(got, want, "case118 split accuracy")
    print(f"max |dgrad| {worst:.3e} vs max|grad| {gmax:.3e} -> {worst / gmax:.2e} relative")
    assert worst <= 2e-5 * gmax + 1e-6, f"3xTF32 split lost accuracy: {worst:.3e} vs {gmax:.3e}"


def test_shared_accumulator_mode_matches_within_tolerance(lib, monkeypatch):
    """GNS_DETERMINISTIC=0: the warps of a CTA add into one accumulator block (reproducible to rounding only)."""
    monkeypatch.setenv("GNS_DETERMINISTIC", "0")
    got, want = _case_grads(300, 40)
    assert_grads_close(got, want, "case300 shared accumulators")
    monkeypatch.delenv("GNS_DETERMINISTIC")
    got2, _ = _case_grads(300, 40)
    worst = max(float((got[n] - got2[n]).abs().max()) for n in got)
    gmax = max(float(w.abs().max()) for w in want.values())
    assert worst <= 1e-5 * gmax, f"shared vs per-warp accumulators differ by {worst:.3e} (max|g| {gmax:.3e})"


@pytest.mark.parametrize("K", [1, 2])
def test_short_recurrences(lib, K):
    """K = 1 has no checkpointed state at all (the only step starts from the initial state), K = 2 one."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=K, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(30, 21, seed=4)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (_, _, otot, _), want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=K,
                                                   latent_dim=20, gamma=0.9, multiple_phi=True)
    out = model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)
    out[2].mean().backward()
    assert_loss_close(out[2], otot, f"K={K} total")
    assert_grads_close(_grads(model), want, f"K={K}")


def test_odd_batch_on_the_two_grids_per_thread_forward(lib):
    """S = 299 is large enough for the VG=2 forward (two grids per thread) and odd: the last CTA-batch holds one
    real grid and one replicated tail grid, whose activations are written but must not reach the gradient."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(30, 299, seed=6)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    (_, _, otot, _), want = orc.gns_loss_and_grads(params, buses.double(), lines.double(), gens.double(), K=4,
                                                   latent_dim=20, gamma=0.9, multiple_phi=True)
    out = model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)
    info = model._last_plan.launch_info(299, 4, 20, 10, True)
    assert info["vector_width"] == 2, info
    out[2].mean().backward()
    assert_loss_close(out[2], otot, "total")
    assert_grads_close(_grads(model), want, "odd batch")
