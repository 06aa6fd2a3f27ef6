"""CPU: host-side logic - packing transform, perturbation recipe, CSR plan, module surface,
and that the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import opf_graph_neural_solver_b200 as pkg
from oracle import gns_oracle as orc
from helpers import GOLDEN, golden_files, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(lib):
    text = open(os.path.join(ROOT, "include", "gns_b200.h")).read()
    declared = set(re.findall(r"\b(gns_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gns_b200.h but not exported"
    assert set(pkg._lib.SYMBOLS) == declared
    assert b"sm_100a" in lib.gns_version()


def test_param_count_matches_reference_formula(lib):
    assert lib.gns_param_count(4, 20, 10, 1) == 14768
    assert lib.gns_param_count(8, 64, 10, 1) == 76704
    assert lib.gns_param_count(4, 20, 10, 0) == 9212
    assert lib.gns_dims_supported(20, 10) == 1 and lib.gns_dims_supported(7, 3) == 0


def test_pack_grids_is_bit_exact_with_reference_prepare_grid():
    z = np.load(os.path.join(GOLDEN, "case14_raw_1_4.npz"))
    b, l, g = pkg.data.pack_grids(z["bus"], z["branch"], z["gen"], float(z["baseMVA"]))
    assert np.array_equal(b.numpy(), z["buses"])
    assert np.array_equal(l.numpy(), z["lines"])
    assert np.array_equal(g.numpy(), z["gens"])
    b1, l1, g1 = pkg.data.pack_grids(z["bus"][0], z["branch"][0], z["gen"][0], float(z["baseMVA"]))
    assert b1.shape == (14, 6) and l1.shape == (20, 7) and g1.shape == (5, 7)


def test_case14_table_equals_reference_pickle():
    z = np.load(os.path.join(GOLDEN, "case14_base.npz"))
    c = pkg.data.case14()
    assert np.array_equal(c["bus"][:, :6], z["bus"][:, :6])
    assert np.array_equal(c["branch"][:, [0, 1, 2, 3, 4, 8, 9]], z["branch"][:, [0, 1, 2, 3, 4, 8, 9]])
    assert np.array_equal(c["gen"][:, [0, 8, 9, 1, 5, 2]], z["gen"][:, [0, 8, 9, 1, 5, 2]])


def test_augment_follows_the_reference_recipe():
    case = pkg.data.case14()
    a = pkg.data.augment(case, 64, seed=3)
    br0, br = case["branch"], a["branch"]
    nz = br0[:, 2] > 0
    assert ((br[:, nz, 2] / br0[nz, 2] >= 0.9) & (br[:, nz, 2] / br0[nz, 2] <= 1.1)).all()
    assert ((br[:, :, 8] >= 0.8) & (br[:, :, 8] <= 1.2)).all()
    assert (np.abs(br[:, :, 9]) <= 0.2).all()
    span = case["gen"][:, 8] - case["gen"][:, 9]
    assert ((a["gen"][:, :, 1] >= 0.25 * span - 1e-9) & (a["gen"][:, :, 1] <= 0.75 * span + 1e-9)).all()
    np.testing.assert_allclose(a["bus"][:, :, 2].sum(1), a["gen"][:, :, 1].sum(1), rtol=1e-12)   # sum Pd == sum Pg
    assert np.array_equal(a["branch"][:, :, :2], np.repeat(br0[None, :, :2], 64, 0))              # topology untouched
    again = pkg.data.augment(case, 64, seed=3)
    assert np.array_equal(again["bus"], a["bus"])                                                   # seeded


@pytest.mark.parametrize("n_bus", [30, 118, 300])
def test_synthetic_case_sizes_and_preconditions(n_bus):
    c = pkg.data.synthetic_case(n_bus, seed=0)
    E, Gn = pkg.data.IEEE_SIZES[n_bus]
    assert c["bus"].shape == (n_bus, 13) and c["branch"].shape == (E, 13) and c["gen"].shape == (Gn, 21)
    f, t = c["branch"][:, 0].astype(int), c["branch"][:, 1].astype(int)
    assert f.min() >= 1 and t.max() <= n_bus and (f != t).all()
    assert len(set(c["gen"][:, 0].astype(int))) == Gn
    assert set(np.concatenate([f, t])) == set(range(1, n_bus + 1))        # connected spanning tree touches all


def _host_plan(f, t, gb, n_bus):
    return pkg.TopologyPlan(f, t, gb, n_bus, device=-1)     # host-only plan: no CUDA needed


@pytest.mark.parametrize("n_bus", [14, 30, 118, 300])
def test_plan_csr_is_bit_exact_against_numpy(lib, n_bus):
    case, _ = pkg.data.get_case(n_bus)
    f = case["branch"][:, 0].astype(np.int32) - 1
    t = case["branch"][:, 1].astype(np.int32) - 1
    gb = case["gen"][:, 0].astype(np.int32) - 1
    plan = _host_plan(f, t, gb, n_bus)
    for key, name in ((t, "in"), (f, "out")):
        rowptr, ids = orc.csr_by(key, n_bus)
        assert np.array_equal(plan.export(name + "_rowptr"), rowptr)
        assert np.array_equal(plan.export(name + "_lines"), ids)
    rowptr, ids = orc.csr_by(gb, n_bus)
    assert np.array_equal(plan.export("gen_rowptr"), rowptr) and np.array_equal(plan.export("gen_ids"), ids)
    in_deg = np.diff(plan.export("in_rowptr"))
    order = orc.degree_order(in_deg)
    assert np.array_equal(plan.export("bus_order"), order)
    assert np.array_equal(plan.export("bus_rank")[order], np.arange(n_bus))


def test_plan_rejects_what_the_reference_cannot_index(lib):
    with pytest.raises(IndexError):        # bus id beyond n_bus (reference: IndexError at m[dst])
        _host_plan([0, 1, 5], [1, 2, 0], [0], 3)
    with pytest.raises(IndexError):        # n_bus > n_line: y_ij[src] out of range in the reference
        _host_plan([0, 1], [1, 2], [0], 3)
    with pytest.raises(IndexError):
        _host_plan([], [], [], 0)
    with pytest.raises(KeyError):
        _host_plan([0, 1, 2], [1, 2, 0], [0], 3).export("nope")


@pytest.mark.parametrize("multi", [True, False])
def test_module_surface_and_state_dict_contract(multi):
    torch.manual_seed(0)
    m = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=multi)
    assert m.multiple_phis is multi and m.latent_dim == 20 and m.K == 4 and m.gamma == 0.9
    names = [n for n, _ in m.named_parameters()]
    assert names == orc.param_names(4, multi)
    ref = orc.init_params(20, 10, 4, multi, seed=0)
    for n, p in m.named_parameters():
        assert torch.equal(p.detach(), ref[n]), n            # same RNG stream as the reference ctor
    assert sum(p.numel() for p in m.parameters()) == (14768 if multi else 9212)
    d = pkg.GNS()                                              # reference defaults
    assert (d.latent_dim, d.hidden_dim, d.K, d.gamma, d.multiple_phis) == (10, 10, 30, 0.9, False)


def test_state_dict_round_trip_with_reference_checkpoint():
    g = load_golden([p for p in golden_files() if p.endswith("k4_l20_multi.npz")][0])
    m = pkg.GNS(latent_dim=20, hidden_dim=10, K=4, gamma=0.9, multiple_phi=True)
    missing = m.load_state_dict(g["params"], strict=True)     # keys saved by the reference load as they are
    assert not missing.missing_keys and not missing.unexpected_keys
    flat = m.flat_parameters()
    assert flat.numel() == 14768 and m._flat_ok()
    off = 0
    for n, p in m.named_parameters():                          # flat buffer is in state_dict order
        assert torch.equal(flat[off:off + p.numel()].view(p.shape), g["params"][n])
        off += p.numel()
    m.load_state_dict(g["params"])                             # in-place copy keeps the views
    assert m._flat_ok()


def test_forward_without_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = pkg.GNS(latent_dim=20, hidden_dim=10, K=2, multiple_phi=True)
    g = load_golden(golden_files()[0])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(g["buses"][0], g["lines"][0], g["gens"][0], *pkg.get_BLG())


def test_block_forward_is_not_a_side_door():
    with pytest.raises(RuntimeError):
        pkg.LearningBlock(3, 4, 5)(torch.zeros(1, 3))


@pytest.mark.parametrize("n_bus", [14, 30, 118, 300])
def test_slot_tables_split_high_degree_buses_consistently(lib, n_bus):
    """Every bus owns 1, 2 or 4 adjacent, aligned slots whose line ranges tile its in-list."""
    case, _ = pkg.data.get_case(n_bus)
    f = case["branch"][:, 0].astype(np.int32) - 1
    t = case["branch"][:, 1].astype(np.int32) - 1
    gb = case["gen"][:, 0].astype(np.int32) - 1
    plan = _host_plan(f, t, gb, n_bus)
    sb, sp, b0, b1, gs = (plan.export(k) for k in ("slot_bus", "slot_primary", "slot_in_begin", "slot_in_end", "slot_gsz"))
    order = plan.export("bus_order")
    deg = np.diff(plan.export("in_rowptr"))
    assert np.array_equal(sb[sp == np.arange(len(sb))], order)           # primaries appear in bus order
    pos = 0
    covered = 0
    for b in order:
        g = gs[pos]
        assert g in (1, 2, 4) and pos % g == 0                           # aligned twin group
        assert (sb[pos:pos + g] == b).all() and (sp[pos:pos + g] == pos).all() and (gs[pos:pos + g] == g).all()
        assert b0[pos] == covered and b1[pos + g - 1] == covered + deg[b]
        assert (b0[pos + 1:pos + g] == b1[pos:pos + g - 1]).all()        # consecutive sub-ranges
        assert (b1[pos:pos + g] - b0[pos:pos + g]).max() == -(-deg[b] // g)      # even split
        assert g == min(4, 1 << max(0, int(np.ceil(np.log2(max(1, -(-deg[b] // 3)))))))   # cap of 3 lines per slot, <= 4 twins
        covered += deg[b]
        pos += g
    assert pos == len(sb) and covered == len(t)


def test_renumber_buses_makes_noncontiguous_ids_usable():
    """quirk Q8: real case300-style bus numbers; the reference would raise IndexError at m[dst]."""
    c = pkg.data.case14()
    ext_ids = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 9011, 9012, 9533, 14])
    lut = {i + 1: int(e) for i, e in enumerate(ext_ids)}
    c2 = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in c.items()}
    c2["bus"][:, 0] = ext_ids
    c2["branch"][:, 0] = [lut[int(x)] for x in c["branch"][:, 0]]
    c2["branch"][:, 1] = [lut[int(x)] for x in c["branch"][:, 1]]
    c2["gen"][:, 0] = [lut[int(x)] for x in c["gen"][:, 0]]
    with pytest.raises(IndexError):
        _host_plan(c2["branch"][:, 0].astype(int) - 1, c2["branch"][:, 1].astype(int) - 1,
                   c2["gen"][:, 0].astype(int) - 1, 14)
    back, ext = pkg.data.renumber_buses(c2)
    assert np.array_equal(ext, ext_ids)
    for k in ("bus", "branch", "gen"):
        assert np.array_equal(back[k], c[k])
    bad = dict(c2); bad["branch"] = c2["branch"].copy(); bad["branch"][0, 0] = 777
    with pytest.raises(IndexError):
        pkg.data.renumber_buses(bad)


def test_device_augmenter_follows_recipe_on_cpu_device():
    """augment_pack_device is torch-only: exercised here on the CPU device (the GPU test repeats it)."""
    case = pkg.data.case14()
    b, l, g = pkg.data.augment_pack_device(case, 64, seed=5, device="cpu")
    assert b.shape == (64, 14, 6) and l.shape == (64, 20, 7) and g.shape == (64, 5, 7)
    assert torch.equal(l[:, :, :2], torch.as_tensor(case["branch"][:, :2], dtype=torch.float32).expand(64, -1, -1))
    assert ((l[:, :, 5] >= 0.8) & (l[:, :, 5] <= 1.2)).all()
    assert (l[:, :, 6].abs() <= np.deg2rad(0.2) + 1e-7).all()
    assert torch.allclose(b[:, :, 2].sum(1), g[:, :, 3].sum(1), rtol=1e-5)          # sum Pd == sum Pg (p.u.)
    assert torch.equal(g[:, :, 3], g[:, :, 6])
    assert torch.allclose(b[:, :, 4], torch.full_like(b[:, :, 4], 0.01)) and torch.allclose(b[:, :, 5], torch.full_like(b[:, :, 5], -0.01))
    b2, _, _ = pkg.data.augment_pack_device(case, 64, seed=5, device="cpu")
    assert torch.equal(b, b2)


@pytest.mark.parametrize("L,multi", [(20, 1), (20, 0), (10, 1), (64, 1)])
def test_gradient_layout_maps_cover_every_parameter_once(lib, L, multi):
    """The backward kernel accumulates weight gradients in MMA-fragment order; `frag` maps the packed layout
    onto it.  Every canonical parameter must be reachable exactly once: either through a fragment cell, or as one
    of the entries the fused block's chain rule derives (linear4 of the phi nets and the S-slice of the L-nets'
    linear1, ref GNS/main.py:157-171)."""
    import ctypes as C
    K, H = 3, 10
    def export(name):
        n = lib.gns_layout_export(name.encode(), K, L, H, multi, None, 0)
        assert n > 0, pkg._lib.last_error()
        buf = (C.c_int32 * n)()
        assert lib.gns_layout_export(name.encode(), K, L, H, multi, buf, n) == n
        return np.frombuffer(buf, dtype=np.int32).copy()
    pack, frag = export("pack"), export("frag")
    assert len(pack) == lib.gns_param_count(K, L, H, multi)
    assert len(np.unique(pack)) == len(pack)                       # canonical -> packed is injective
    wstep = len(frag)
    used = frag[frag >= 0]
    assert len(np.unique(used)) == len(used)                       # no two packed entries share a fragment cell
    # names of the canonical parameters, in order
    m = pkg.GNS(latent_dim=L, hidden_dim=H, K=K, multiple_phi=bool(multi))
    names = []
    for n, p in m.named_parameters():
        names += [n] * p.numel()
    assert len(names) == len(pack)
    sizes = {n: p.shape for n, p in m.named_parameters()}
    offs, seen = {}, {}
    for i, n in enumerate(names):
        offs.setdefault(n, i)
    derived = 0
    for i, n in enumerate(names):
        step_local = pack[i] % wstep
        if frag[step_local] >= 0:
            continue
        derived += 1
        net, _, layer, kind = n.split(".")
        if net.startswith("phi"):
            assert layer == "linear4", n                           # W4 / b4 come from dM / dc
        else:
            assert layer == "linear1" and kind == "weight", n      # ... and so does W1[:, 4+L:]
            col = (i - offs[n]) % sizes[n][1]
            assert col >= 4 + L, (n, col)
    assert derived > 0


def test_fragment_maps_are_injective(lib):
    """Every packed weight the backward kernels write maps to its own accumulator cell, for both kernels' layouts."""
    for name in (b"frag", b"frag2", b"frag3"):
        for multi in (1, 0):
            if name == b"frag3" and not multi:
                continue                      # the fragment-space kernel is built for multiple_phi only
            n = lib.gns_layout_export(name, 4, 20, 10, multi, None, 0)
            assert n > 0
            out = np.zeros(n, dtype=np.int32)
            lib.gns_layout_export(name, 4, 20, 10, multi, out.ctypes.data, n)
            used = out[out >= 0]
            assert len(np.unique(used)) == len(used) and used.min() >= 0
    # the two layouts carry the same set of packed entries
    a = np.zeros(n, dtype=np.int32); b = np.zeros(n, dtype=np.int32)
    lib.gns_layout_export(b"frag", 4, 20, 10, 0, a.ctypes.data, n)
    lib.gns_layout_export(b"frag2", 4, 20, 10, 0, b.ctypes.data, n)
    assert np.array_equal(a >= 0, b >= 0)
    n3 = lib.gns_layout_export(b"frag3", 4, 20, 10, 1, None, 0)
    c = np.zeros(n3, dtype=np.int32); a1 = np.zeros(n3, dtype=np.int32)
    lib.gns_layout_export(b"frag3", 4, 20, 10, 1, c.ctypes.data, n3)
    lib.gns_layout_export(b"frag", 4, 20, 10, 1, a1.ctypes.data, n3)
    assert np.array_equal(a1 >= 0, c >= 0)


def test_pack_varying_round_trip_and_rejection():
    b, l, g, _ = pkg.data.make_batch(30, 6, seed=4)
    var, const = pkg.data.pack_varying(b, l, g)
    assert var[0].shape == (6, 30, 2) and var[1].shape == (6, 41, 5) and var[2].shape == (6, 6, 2)
    b2, l2, g2 = pkg.data.expand_varying(var, const)
    assert torch.equal(b, b2) and torch.equal(l, l2) and torch.equal(g, g2)
    bad = l.clone(); bad[3, 0, 1] += 1.0           # another topology in grid 3
    with pytest.raises(ValueError):
        pkg.data.pack_varying(b, bad, g)


def test_topology_check_on_host_tensors():
    b, l, g, _ = pkg.data.make_batch(14, 3, seed=0)
    f, t, gb = l[0, :, 0].numpy().astype(int) - 1, l[0, :, 1].numpy().astype(int) - 1, g[0, :, 0].numpy().astype(int) - 1
    plan = _host_plan(f, t, gb, 14)
    assert plan.matches_host(l, g) and plan.matches_host(l[1], g[1])
    bad = l.clone(); bad[2, 5, 0] = 3.0 if bad[2, 5, 0] != 3.0 else 4.0
    assert not plan.matches_host(bad, g)


def test_copied_model_rebuilds_its_gradient_leaf():
    import copy
    m = pkg.GNS(latent_dim=10, hidden_dim=10, K=2, multiple_phi=True)
    m.flat_parameters()
    leaf = m._leaf()
    assert m._leaf() is leaf
    twin = copy.deepcopy(m)
    twin.flat_parameters()
    assert twin._leaf() is not leaf and twin._leaf_owner == id(twin)


def test_unsupported_dims_fail_at_construction(lib):
    with pytest.raises(ValueError):
        pkg.GNS(latent_dim=7, hidden_dim=3, K=2)
