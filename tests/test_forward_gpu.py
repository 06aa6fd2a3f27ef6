"""GPU parity of the forward kernel, through the module -> ctypes -> C ABI path."""
import numpy as np
import pytest
import torch

import opf_graph_neural_solver_b200 as pkg
from oracle import gns_oracle as orc
from helpers import assert_bus_close, assert_loss_close, golden_files, load_golden

pytestmark = pytest.mark.gpu
BLG = pkg.get_BLG()


def _model_from(params, latent_dim, hidden_dim, K, gamma, multiple_phi):
    m = pkg.GNS(latent_dim=latent_dim, hidden_dim=hidden_dim, K=K, gamma=gamma, multiple_phi=multiple_phi)
    m.load_state_dict(params)
    return m.cuda()


def _check_against_oracle(model, buses, lines, gens, what):
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    want = orc.gns_forward(params, buses.double(), lines.double(), gens.double(), K=model.K,
                           latent_dim=model.latent_dim, gamma=model.gamma, multiple_phi=model.multiple_phis)
    with torch.no_grad():
        got = model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)
    assert_bus_close(got[0], want[0], what + " v")
    assert_bus_close(got[1], want[1], what + " theta")
    assert_loss_close(got[2], want[2], what + " total_loss")
    assert_loss_close(got[3], want[3], what + " last_loss")
    return got


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("ref_case14_")[-1][:-4])
def test_forward_matches_reference_golden(lib, path):
    g = load_golden(path)
    model = _model_from(g["params"], g["latent_dim"], g["hidden_dim"], g["K"], g["gamma"], g["multiple_phi"])
    with torch.no_grad():
        v, th, tot, last = model(g["buses"].cuda(), g["lines"].cuda(), g["gens"].cuda(), *BLG)
    assert_bus_close(v, g["v"], "v")
    assert_bus_close(th, g["theta"], "theta")
    assert_loss_close(tot, g["total_loss"], "total_loss")
    assert_loss_close(last, g["last_loss"], "last_loss")


def test_single_grid_call_has_reference_shapes_and_cpu_round_trip(lib):
    g = load_golden([p for p in golden_files() if p.endswith("k4_l20_multi.npz")][0])
    model = _model_from(g["params"], 20, 10, 4, 0.9, True)
    with torch.no_grad():   # CPU tensors in (the reference's usage) -> CPU tensors out
        v, th, tot, last = model(g["buses"][1], g["lines"][1], g["gens"][1], *BLG)
    assert v.shape == (14,) and th.shape == (14,) and tot.shape == () and last.shape == ()
    assert v.device.type == "cpu"
    assert_bus_close(v, g["v"][1], "v")
    assert_loss_close(tot, g["total_loss"][1], "total")
    # keyword call like ref GNS/main.py:281
    with torch.no_grad():
        out = model(buses=g["buses"][1], lines=g["lines"][1], generators=g["gens"][1], B=BLG[0], L=BLG[1], G=BLG[2])
    assert torch.equal(out[0], v)


@pytest.mark.parametrize("n_bus,S", [(14, 1), (14, 37), (30, 101), (118, 19), (300, 7), (300, 297)])
def test_forward_synthetic_cases_vs_oracle(lib, n_bus, S):
    """ragged batch sizes (tail CTA batches), every IEEE size, seed-0 weights."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(n_bus, S, seed=5)
    _check_against_oracle(model, buses, lines, gens, f"case{n_bus} S={S}")


@pytest.mark.parametrize("vg,ngq", [(1, 1), (2, 1), (1, 4), (2, 8), (2, 16)])
def test_forward_is_independent_of_launch_geometry(lib, monkeypatch, vg, ngq):
    monkeypatch.setenv("GNS_FWD_VG", str(vg))
    monkeypatch.setenv("GNS_FWD_NGQ", str(ngq))
    torch.manual_seed(1)
    model = pkg.GNS(latent_dim=10, hidden_dim=10, K=3, gamma=0.9, multiple_phi=False).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(14, 53, seed=2)
    _check_against_oracle(model, buses, lines, gens, f"VG={vg} NGQ={ngq}")


def test_both_lambda_arms_are_exercised(lib):
    """ref GNS/main.py:47-57: low load takes the `if` arms, nominal load the `else` arms."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(30, 64, seed=9)
    lo = buses.clone(); lo[:, :, 2] *= 0.3
    hi = buses.clone(); hi[:, :, 2] *= 3.0
    mixed = torch.cat([lo[:32], hi[32:]])
    _check_against_oracle(model, mixed, lines, gens, "mixed lambda arms")


def test_stress_config_k8_l64(lib):
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=64, hidden_dim=10, K=8, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(300, 5, seed=3)
    _check_against_oracle(model, buses, lines, gens, "case300 K=8 L=64")


def test_duplicate_generator_buses_sum_like_the_reference(lib):
    """quirk Q7: two generators on one bus add their vg / Pg (ref GNS/main.py:146-151)."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=2, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(14, 6, seed=4)
    gens = gens.clone(); gens[:, 1, 0] = gens[:, 0, 0]
    _check_against_oracle(model, buses, lines, gens, "duplicate gen bus")


def test_plan_rejects_a_batch_of_another_topology(lib):
    """One plan = one topology: the check the forward relies on (mixed batches are then split by topology, see
    test_heterogeneous_topology_batch_is_grouped_and_matches_per_sample_calls)."""
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=2, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(14, 4, seed=4)
    lines = lines.clone(); lines[2, 3, 1] = 9.0
    with pytest.raises(ValueError, match="share one topology"):
        model.plan_for(lines.cuda(), gens.cuda(), 14)


def test_unsupported_dims_fail_loudly(lib):
    with pytest.raises(ValueError, match="no fallback"):          # at construction, not at the first forward
        pkg.GNS(latent_dim=7, hidden_dim=3, K=2)
    # the C ABI refuses them as well
    assert lib.gns_dims_supported(7, 3) == 0


def test_negative_voltage_is_clamped_like_the_reference(lib):
    """ref GNS/main.py:201: v<0 -> 0 on output only."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=2, gamma=0.9, multiple_phi=True).cuda()
    with torch.no_grad():
        model.L_v["1"].linear4.bias.fill_(-5.0)      # drives non-generator voltages negative
    buses, lines, gens, _ = pkg.data.make_batch(14, 3, seed=4)
    got = _check_against_oracle(model, buses, lines, gens, "clamp")
    assert float(got[0].min()) == 0.0


def test_infer_host_streams_chunks_and_matches_forward(lib):
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(30, 1000, seed=6)
    with torch.no_grad():
        want = model(buses.cuda(), lines.cuda(), gens.cuda(), *BLG)
    hb, hl, hg = buses.pin_memory(), lines.pin_memory(), gens.pin_memory()
    got = model.infer_host(hb, hl, hg, chunk=192)   # ragged last chunk
    # chunk by chunk the pipeline runs exactly the kernel a device-resident call of that chunk runs: bit-identical
    with torch.no_grad():
        bounds = pkg.model.chunk_bounds(1000, 192)
        assert bounds[0][0] == 0 and bounds[-1][1] == 1000 and all(x[1] == y[0] for x, y in zip(bounds, bounds[1:]))
        for a, b in bounds:
            per = model(buses[a:b].cuda(), lines[a:b].cuda(), gens[a:b].cuda(), *BLG)
            for g, w in zip(got, per):
                assert g.device.type == "cpu" and torch.equal(g[a:b], w.cpu())
    for g, w in zip(got, want):   # the whole batch may pick another launch geometry: same math, other summation order
        assert torch.allclose(g, w.cpu(), rtol=1e-4, atol=1e-5)
    # compact input format (only the columns that vary between samples travel): same results, bit for bit
    var, const = pkg.data.pack_varying(hb, hl, hg)
    got2 = model.infer_host_compact(tuple(t.pin_memory() for t in var), const, chunk=192)
    for g, w in zip(got2, got):
        assert torch.equal(g, w)
    # the device-side expansion of the compact format (gns_expand_inputs) rebuilds the packed rows bit for bit
    eb, el, eg = pkg.data.expand_varying_device(var, const)
    assert torch.equal(eb.cpu(), buses) and torch.equal(el.cpu(), lines) and torch.equal(eg.cpu(), gens)
    # a batch whose topology changes after the first chunk is rejected (every chunk is checked)
    bad = hl.clone()
    bad[700, 3, 0], bad[700, 3, 1] = bad[700, 4, 0], bad[700, 4, 1]
    with pytest.raises(ValueError):
        model.infer_host(hb, bad.pin_memory(), hg, chunk=192)


def test_reference_default_constructor_k30(lib):
    """GNS() defaults of the reference (latent 10, hidden 10, K=30, single phi).  At random init the
    30-step recurrence is ill-conditioned, so the weights are damped to keep the
    float32/float64 comparison meaningful."""
    torch.manual_seed(0)
    model = pkg.GNS().cuda()
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(0.5)
    buses, lines, gens, _ = pkg.data.make_batch(14, 9, seed=8)
    _check_against_oracle(model, buses, lines, gens, "defaults K=30")


def test_large_grid_uses_the_wide_cta_variant(lib):
    """700 buses / 900 lines: more than 384 threads per grid -> the 1024-thread kernel variant."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=10, hidden_dim=10, K=2, gamma=0.9, multiple_phi=True).cuda()
    case = pkg.data.synthetic_case(700, 900, 80, seed=2)
    aug = pkg.data.augment(case, 5, seed=1)
    buses, lines, gens = pkg.data.pack_grids(aug["bus"], aug["branch"], aug["gen"], aug["baseMVA"])
    _check_against_oracle(model, buses, lines, gens, "700-bus grid")
    info = model._last_plan.launch_info(5, 2, 10, 10, True)
    assert info["threads"] > 384 and info["tmax"] == 1024


def test_device_side_augmenter_feeds_the_kernel(lib):
    """perturb + pack on the GPU, no host round trip; the batch must be a valid input."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    case, _ = pkg.data.get_case(118)
    b, l, g = pkg.data.augment_pack_device(case, 300, seed=3, device="cuda")
    assert b.is_cuda and b.shape == (300, 118, 6)
    _check_against_oracle(model, b.cpu(), l.cpu(), g.cpu(), "device-augmented case118")


def test_heterogeneous_topology_batch_is_grouped_and_matches_per_sample_calls(lib):
    """A batch that mixes topologies (here: generator sets and one re-routed line) runs group by group; results and
    gradients equal the reference-style per-sample loop (ref GNS/main.py:153 rebuilds the indices per sample)."""
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=3, gamma=0.9, multiple_phi=True).cuda()
    buses, lines, gens, _ = pkg.data.make_batch(14, 9, seed=8)
    lines, gens = lines.clone(), gens.clone()
    lines[2::3, 5, 1] = 7.0            # every third grid: line 5 ends at bus 7 instead
    gens[1::3, 4, 0] = 9.0             # every third grid (shifted): the last generator sits at bus 9
    b, l, g = buses.cuda(), lines.cuda(), gens.cuda()
    out = model(b, l, g, *BLG)
    out[2].mean().backward()
    got = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    singles = [model(b[i], l[i], g[i], *BLG) for i in range(9)]
    torch.stack([s[2] for s in singles]).mean().backward()
    for k in range(4):
        assert torch.allclose(out[k], torch.stack([s[k] for s in singles]), rtol=1e-5, atol=1e-6)
    for n, p in model.named_parameters():
        assert torch.allclose(got[n], p.grad, rtol=1e-4, atol=1e-6), n
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    for i in (0, 1, 2):                # and each kind against the oracle
        ov, oth, otot, olast = orc.gns_forward(params, buses[i], lines[i], gens[i], K=3, latent_dim=20, gamma=0.9,
                                               multiple_phi=True)
        assert_bus_close(out[0][i], ov, f"v[{i}]"); assert_loss_close(out[2][i:i + 1], otot.reshape(1), f"total[{i}]")


def test_non_contiguous_bus_numbers_run_after_renumbering(lib):
    """The real IEEE-300 table numbers its buses up to 9533, on which the reference raises IndexError at ``m[dst]``
    (ref GNS/main.py:153, quirk Q8); ``data.renumber_buses`` maps them to 1..N and the GPU path runs."""
    case = pkg.data.case30()
    ext = case["bus"][:, 0] * 37 + 1000                      # external numbering with gaps
    lut = {int(i + 1): float(e) for i, e in enumerate(ext)}
    wild = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in case.items()}
    wild["bus"][:, 0] = ext
    wild["branch"][:, 0] = [lut[int(x)] for x in case["branch"][:, 0]]
    wild["branch"][:, 1] = [lut[int(x)] for x in case["branch"][:, 1]]
    wild["gen"][:, 0] = [lut[int(x)] for x in case["gen"][:, 0]]
    with pytest.raises(IndexError):                           # like the reference: ids outside 1..N cannot be indexed
        aug = pkg.data.augment(wild, 2, seed=1)
        b, l, g = pkg.data.pack_grids(aug["bus"], aug["branch"], aug["gen"], aug["baseMVA"])
        pkg.TopologyPlan.from_tensors(l, g, 30, device=-1)
    fixed, ids = pkg.data.renumber_buses(wild)
    assert np.array_equal(ids, ext.astype(np.int64))
    aug = pkg.data.augment(fixed, 5, seed=1)
    b, l, g = pkg.data.pack_grids(aug["bus"], aug["branch"], aug["gen"], aug["baseMVA"])
    ref = pkg.data.augment(case, 5, seed=1)
    rb, rl, rg = pkg.data.pack_grids(ref["bus"], ref["branch"], ref["gen"], ref["baseMVA"])
    assert torch.equal(b, rb) and torch.equal(l, rl) and torch.equal(g, rg)
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True).cuda()
    _check_against_oracle(model, b, l, g, "renumbered case30")
