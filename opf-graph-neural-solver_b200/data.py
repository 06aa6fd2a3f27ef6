"""Grid tables -> packed tensors, and the synthetic load-perturbed samples used for
measurement (host-side PyTorch / numpy; not on the GPU hot path).

* ``pack_grids``   restates ``prepare_grid``  (ref GNS/utils.py:17-41), batched and bit-exact
* ``augment``      restates the perturbation recipe of ref GNS/augment_grids.py:25-54 with a
                   seeded ``numpy.random.default_rng`` (the reference uses unseeded ``np.random``)
* ``case14``       the IEEE 14-bus table (standard public test case; same numbers as
                   ``data/case14/augmented_case14_0.pkl`` of the reference)
* ``case30``       the IEEE 30-bus table (MATPOWER / pypower ``case30``), pinned by a Newton-Raphson known answer
* ``synthetic_case`` an IEEE-*sized* random topology for 118 / 300 buses (hub degree and parallel lines like the
                   published systems): the real
                   tables ship with pypower, which is not available offline.
                   Results obtained on these are labelled "IEEE-sized synthetic topology".
"""
from __future__ import annotations

import numpy as np
import torch

# n_bus -> (n_line, n_gen), ref GNS/utils.py:45-56
IEEE_SIZES = {14: (20, 5), 30: (41, 6), 118: (186, 54), 300: (411, 69)}

# pypower column numbers (caseformat): bus 13 cols, branch 13 cols, gen 21 cols
_BUS_I, _BUS_TYPE, _PD, _QD, _GS, _BS = 0, 1, 2, 3, 4, 5
_F, _T, _R, _X, _B, _TAP, _SHIFT = 0, 1, 2, 3, 4, 8, 9
_GBUS, _PG, _QG, _VG, _PMAX, _PMIN = 0, 1, 2, 5, 8, 9


def _tables(bus6, br7, gen6, base_mva=100.0):
    bus = np.zeros((len(bus6), 13)); branch = np.zeros((len(br7), 13)); gen = np.zeros((len(gen6), 21))
    bus[:, :6] = bus6
    branch[:, [_F, _T, _R, _X, _B, _TAP, _SHIFT]] = br7
    gen[:, [_GBUS, _PMAX, _PMIN, _PG, _VG, _QG]] = gen6
    return {"baseMVA": float(base_mva), "bus": bus, "branch": branch, "gen": gen}


def case14():
    """IEEE 14-bus test case (bus: id,type,Pd,Qd,Gs,Bs; branch: f,t,r,x,b,ratio,angle;
    gen: bus,Pmax,Pmin,Pg,Vg,Qg)."""
    bus = [[1, 3, 0, 0, 0, 0], [2, 2, 21.7, 12.7, 0, 0], [3, 2, 94.2, 19, 0, 0], [4, 1, 47.8, -3.9, 0, 0],
           [5, 1, 7.6, 1.6, 0, 0], [6, 2, 11.2, 7.5, 0, 0], [7, 1, 0, 0, 0, 0], [8, 2, 0, 0, 0, 0],
           [9, 1, 29.5, 16.6, 0, 19], [10, 1, 9, 5.8, 0, 0], [11, 1, 3.5, 1.8, 0, 0], [12, 1, 6.1, 1.6, 0, 0],
           [13, 1, 13.5, 5.8, 0, 0], [14, 1, 14.9, 5, 0, 0]]
    br = [[1, 2, 0.01938, 0.05917, 0.0528, 0, 0], [1, 5, 0.05403, 0.22304, 0.0492, 0, 0],
          [2, 3, 0.04699, 0.19797, 0.0438, 0, 0], [2, 4, 0.05811, 0.17632, 0.034, 0, 0],
          [2, 5, 0.05695, 0.17388, 0.0346, 0, 0], [3, 4, 0.06701, 0.17103, 0.0128, 0, 0],
          [4, 5, 0.01335, 0.04211, 0, 0, 0], [4, 7, 0, 0.20912, 0, 0.978, 0], [4, 9, 0, 0.55618, 0, 0.969, 0],
          [5, 6, 0, 0.25202, 0, 0.932, 0], [6, 11, 0.09498, 0.1989, 0, 0, 0], [6, 12, 0.12291, 0.25581, 0, 0, 0],
          [6, 13, 0.06615, 0.13027, 0, 0, 0], [7, 8, 0, 0.17615, 0, 0, 0], [7, 9, 0, 0.11001, 0, 0, 0],
          [9, 10, 0.03181, 0.0845, 0, 0, 0], [9, 14, 0.12711, 0.27038, 0, 0, 0], [10, 11, 0.08205, 0.19207, 0, 0, 0],
          [12, 13, 0.22092, 0.19988, 0, 0, 0], [13, 14, 0.17093, 0.34802, 0, 0, 0]]
    gen = [[1, 332.4, 0, 232.4, 1.06, -16.9], [2, 140, 0, 40, 1.045, 42.4], [3, 100, 0, 0, 1.01, 23.4],
           [6, 100, 0, 0, 1.07, 12.2], [8, 100, 0, 0, 1.09, 17.4]]
    return _tables(np.array(bus, float), np.array(br, float), np.array(gen, float))


def case30():
    """IEEE 30-bus test case as shipped with MATPOWER / pypower (``case30``: 30 buses, 41 branches, 6 generators at
    buses 1, 2, 22, 27, 23, 13; Alsac & Stott data) - what ``pypower.api.case30()`` returns in the reference
    (ref GNS/augment_grids.py:1,8).  Entered by hand (pypower is not installable offline) and pinned by a
    Newton-Raphson known-answer test: losses 2.444 MW, slack 25.97 MW, min |V| 0.961 at bus 8, angles -3.96 deg at
    bus 19 / +1.48 deg at bus 13 (the published ``runpf`` result)."""
    pd = {2: (21.7, 12.7), 3: (2.4, 1.2), 4: (7.6, 1.6), 7: (22.8, 10.9), 8: (30, 30), 10: (5.8, 2), 12: (11.2, 7.5),
          14: (6.2, 1.6), 15: (8.2, 2.5), 16: (3.5, 1.8), 17: (9, 5.8), 18: (3.2, 0.9), 19: (9.5, 3.4), 20: (2.2, 0.7),
          21: (17.5, 11.2), 23: (3.2, 1.6), 24: (8.7, 6.7), 26: (3.5, 2.3), 29: (2.4, 0.9), 30: (10.6, 1.9)}
    bus = np.zeros((30, 6))
    bus[:, 0] = np.arange(1, 31)
    bus[:, 1] = 1
    for b in (2, 13, 22, 23, 27):
        bus[b - 1, 1] = 2
    bus[0, 1] = 3
    for b, (p, q) in pd.items():
        bus[b - 1, 2], bus[b - 1, 3] = p, q
    bus[4, 5], bus[23, 5] = 0.19, 0.04
    br = [[1, 2, 0.02, 0.06, 0.03], [1, 3, 0.05, 0.19, 0.02], [2, 4, 0.06, 0.17, 0.02], [3, 4, 0.01, 0.04, 0],
          [2, 5, 0.05, 0.2, 0.02], [2, 6, 0.06, 0.18, 0.02], [4, 6, 0.01, 0.04, 0], [5, 7, 0.05, 0.12, 0.01],
          [6, 7, 0.03, 0.08, 0.01], [6, 8, 0.01, 0.04, 0], [6, 9, 0, 0.21, 0], [6, 10, 0, 0.56, 0], [9, 11, 0, 0.21, 0],
          [9, 10, 0, 0.11, 0], [4, 12, 0, 0.26, 0], [12, 13, 0, 0.14, 0], [12, 14, 0.12, 0.26, 0], [12, 15, 0.07, 0.13, 0],
          [12, 16, 0.09, 0.2, 0], [14, 15, 0.22, 0.2, 0], [16, 17, 0.08, 0.19, 0], [15, 18, 0.11, 0.22, 0],
          [18, 19, 0.06, 0.13, 0], [19, 20, 0.03, 0.07, 0], [10, 20, 0.09, 0.21, 0], [10, 17, 0.03, 0.08, 0],
          [10, 21, 0.03, 0.07, 0], [10, 22, 0.07, 0.15, 0], [21, 22, 0.01, 0.02, 0], [15, 23, 0.1, 0.2, 0],
          [22, 24, 0.12, 0.18, 0], [23, 24, 0.13, 0.27, 0], [24, 25, 0.19, 0.33, 0], [25, 26, 0.25, 0.38, 0],
          [25, 27, 0.11, 0.21, 0], [28, 27, 0, 0.4, 0], [27, 29, 0.22, 0.42, 0], [27, 30, 0.32, 0.6, 0],
          [29, 30, 0.24, 0.45, 0], [8, 28, 0.06, 0.2, 0.02], [6, 28, 0.02, 0.06, 0.01]]
    br7 = np.zeros((41, 7))
    br7[:, :5] = br
    # bus, Pmax, Pmin, Pg, Vg, Qg
    gen = [[1, 80, 0, 23.54, 1, 0], [2, 80, 0, 60.97, 1, 0], [22, 50, 0, 21.59, 1, 0], [27, 55, 0, 26.91, 1, 0],
           [23, 30, 0, 19.2, 1, 0], [13, 40, 0, 37, 1, 0]]
    case = _tables(bus, br7, np.array(gen, float))
    case["bus"][:, 7] = 1.0
    return case


def degree_profile(case: dict):
    """Degree statistics of a case's topology (what drives the twin-slot split and the barrier skew of the kernels)."""
    f, t = case["branch"][:, 0].astype(int), case["branch"][:, 1].astype(int)
    n = case["bus"].shape[0]
    deg = np.bincount(np.concatenate([f, t]), minlength=n + 1)[1:]
    indeg = np.bincount(t, minlength=n + 1)[1:]
    pairs = {}
    for a, b in zip(np.minimum(f, t), np.maximum(f, t)):
        pairs[(a, b)] = pairs.get((a, b), 0) + 1
    return {"max_degree": int(deg.max()), "max_in_degree": int(indeg.max()), "mean_degree": float(deg.mean()),
            "parallel_lines": int(sum(v - 1 for v in pairs.values())), "in_degree_hist": np.bincount(indeg).tolist()}


# Published statistics of the IEEE tables the generator imitates (the tables themselves ship with pypower, which is
# not available offline): the 118-bus system has 186 branches over 179 distinct bus pairs (7 parallel lines) and its
# busiest bus (49) carries 12 branches; the 300-bus system's busiest buses carry about a dozen branches and it has a
# handful of parallel lines.  (n_parallel, hub_degree) per size.
_IEEE_SHAPE = {118: (7, 12), 300: (4, 12)}


def synthetic_case(n_bus: int, n_line: int | None = None, n_gen: int | None = None, seed: int = 0,
                   n_parallel: int | None = None, hub_degree: int | None = None):
    """IEEE-sized synthetic topology: seeded random spanning tree, one hub grown to ``hub_degree`` branches,
    ``n_parallel`` duplicated lines, then random chords (no self loops), contiguous bus ids 1..n_bus, generators on
    distinct buses, base values drawn from case14-like ranges.  Hub degree and parallel-line count
    default to the statistics of the IEEE system of that size (``_IEEE_SHAPE``)."""
    if n_line is None or n_gen is None:
        n_line, n_gen = IEEE_SIZES[n_bus]
    if n_line < n_bus:
        raise ValueError("need n_line >= n_bus (the reference indexes line vectors by bus number)")
    if n_parallel is None or hub_degree is None:
        dp, dh = _IEEE_SHAPE.get(n_bus, (0, 0))
        n_parallel = dp if n_parallel is None else n_parallel
        hub_degree = dh if hub_degree is None else hub_degree
    rng = np.random.default_rng(seed)
    order = rng.permutation(n_bus)
    f, t = [], []
    for i in range(1, n_bus):                       # spanning tree
        a, b = order[i], order[rng.integers(0, i)]
        f.append(min(a, b)); t.append(max(a, b))
    if hub_degree:                                  # grow the busiest bus to the published maximum degree
        deg = np.bincount(np.array(f + t), minlength=n_bus)
        hub = int(deg.argmax())
        while deg[hub] < hub_degree and len(f) < n_line - n_parallel:
            o = int(rng.integers(0, n_bus))
            if o != hub and (min(o, hub), max(o, hub)) not in set(zip(f, t)):
                f.append(min(o, hub)); t.append(max(o, hub)); deg[hub] += 1
    for _ in range(min(n_parallel, max(0, n_line - len(f)))):   # parallel lines (the real 118 / 300 tables have them)
        j = int(rng.integers(0, len(f)))
        while hub_degree and hub in (f[j], t[j]):
            j = int(rng.integers(0, len(f)))
        f.append(f[j]); t.append(t[j])
    have = set(zip(f, t))
    hub_id = hub if hub_degree else -1
    while len(f) < n_line:                          # chords: new bus pairs, away from the hub
        a, b = rng.integers(0, n_bus, size=2)
        lo, hi = int(min(a, b)), int(max(a, b))
        if lo != hi and (lo, hi) not in have and hub_id not in (lo, hi):
            f.append(lo); t.append(hi); have.add((lo, hi))
    idx = np.lexsort((t, f))                        # sorted by (f, t) like the IEEE tables
    f, t = np.array(f)[idx] + 1, np.array(t)[idx] + 1
    br = np.zeros((n_line, 7))
    br[:, 0], br[:, 1] = f, t
    br[:, 2] = rng.uniform(0.0, 0.24, n_line)
    br[:, 3] = rng.uniform(0.04, 0.61, n_line)
    br[:, 4] = rng.uniform(0.0, 0.06, n_line)
    gen_bus = np.sort(rng.choice(n_bus, size=n_gen, replace=False)) + 1
    bus = np.zeros((n_bus, 6))
    bus[:, 0] = np.arange(1, n_bus + 1)
    bus[:, 1] = 1
    bus[gen_bus - 1, 1] = 2
    bus[gen_bus[0] - 1, 1] = 3
    # Injections per bus shrink with the size of the system (x 1.2 / sqrt(n_bus)): with case14-like per-bus power the
    # backbone lines of a 300-bus random tree would carry several p.u. and no power flow exists (a flat-start
    # Newton-Raphson then diverges; with this scaling it converges in 4-7 iterations on nominal taps).
    ps = min(1.0, 1.2 / np.sqrt(n_bus))
    bus[:, 2] = ps * rng.uniform(0.0, 100.0, n_bus) * (rng.uniform(size=n_bus) < 0.8)   # Pd (rescaled by augment)
    bus[:, 3] = bus[:, 2] * rng.uniform(-0.1, 0.5, n_bus)                                # Qd
    gen = np.zeros((n_gen, 6))
    gen[:, 0] = gen_bus
    gen[:, 1] = ps * rng.uniform(100.0, 330.0, n_gen)    # Pmax
    gen[:, 2] = 0.0                                       # Pmin
    gen[:, 3] = gen[:, 1] * rng.uniform(0.2, 0.7, n_gen)
    gen[:, 4] = rng.uniform(1.0, 1.1, n_gen)             # Vg
    gen[:, 5] = ps * rng.uniform(-17.0, 42.0, n_gen)     # Qg
    return _tables(bus, br, gen)


def get_case(n_bus: int, seed: int = 0):
    """case14 / case30 -> the IEEE tables; 118 / 300 -> IEEE-sized synthetic topology (labelled)."""
    if n_bus == 14:
        return case14(), "IEEE case14"
    if n_bus == 30 and seed == 0:
        return case30(), "IEEE case30"
    return synthetic_case(n_bus, seed=seed), f"IEEE-sized synthetic topology ({n_bus} buses)"


def augment(case: dict, n_samples: int, seed: int = 0, nominal_taps: bool = False):
    """Vectorised restatement of the perturbation loop of ref GNS/augment_grids.py:35-53.
    Returns float64 tables ``bus [S,N,13]``, ``branch [S,E,13]``, ``gen [S,Gn,21]``.
    ``nominal_taps=True`` keeps the case's own tap ratios / phase shifts instead of drawing U(0.8, 1.2) / U(-0.2, 0.2)
    for EVERY line like the reference does (ref :43-45): random off-nominal taps on all lines of a meshed grid leave
    no solvable power flow, so the Newton-Raphson baseline is timed on the nominal-tap variant (labelled)."""
    rng = np.random.default_rng(seed)
    S = int(n_samples)
    bus = np.repeat(np.asarray(case["bus"], dtype=np.float64)[None], S, axis=0)
    branch = np.repeat(np.asarray(case["branch"], dtype=np.float64)[None], S, axis=0)
    gen = np.repeat(np.asarray(case["gen"], dtype=np.float64)[None], S, axis=0)
    E, Gn, N = branch.shape[1], gen.shape[1], bus.shape[1]
    branch[:, :, _R] *= rng.uniform(0.9, 1.1, (S, E))
    branch[:, :, _X] *= rng.uniform(0.9, 1.1, (S, E))
    branch[:, :, _B] *= rng.uniform(0.9, 1.1, (S, E))
    tap, shift = rng.uniform(0.8, 1.2, (S, E)), rng.uniform(-0.2, 0.2, (S, E))     # drawn either way: same random stream
    if not nominal_taps:
        branch[:, :, _TAP], branch[:, :, _SHIFT] = tap, shift
    gen[:, :, _VG] *= rng.uniform(0.95, 1.05, (S, Gn))
    span = gen[:, :, _PMAX] - gen[:, :, _PMIN]
    # the reference draws Pg from U(Pmin + 0.25 span, 0.75 span)  (augment_grids.py:47-49)
    gen[:, :, _PG] = rng.uniform(gen[:, :, _PMIN] + 0.25 * span, 0.75 * span)
    bus[:, :, _PD] *= rng.uniform(0.5, 1.5, (S, N))
    bus[:, :, _PD] *= (gen[:, :, _PG].sum(axis=1) / bus[:, :, _PD].sum(axis=1))[:, None]
    bus[:, :, _QD] *= rng.uniform(0.5, 1.5, (S, N))
    return {"baseMVA": float(case["baseMVA"]), "bus": bus, "branch": branch, "gen": gen}


def pack_grids(bus, branch, gen, base_mva: float):
    """Batched ``prepare_grid`` (ref GNS/utils.py:17-41), same operation order in float32:
    cast to f32, override Gs=1 / Bs=-1, divide P,Q,Gs,Bs by baseMVA, tau==0 -> 1,
    shift degrees -> radians, generator columns [bus, Pmax, Pmin, Pg, Vg, Qg] + copy of Pg."""
    bus = torch.as_tensor(np.asarray(bus), dtype=torch.float32)
    branch = torch.as_tensor(np.asarray(branch), dtype=torch.float32)
    gen = torch.as_tensor(np.asarray(gen), dtype=torch.float32)
    single = bus.dim() == 2
    if single:
        bus, branch, gen = bus[None], branch[None], gen[None]
    buses = bus[:, :, [0, 1, 2, 3, 4, 5]].clone()
    buses[:, :, 4] = 1.0
    buses[:, :, 5] = -1.0
    buses[:, :, [2, 3, 4, 5]] /= base_mva
    lines = branch[:, :, [0, 1, 2, 3, 4, 8, 9]].clone()
    lines[:, :, 5] = torch.where(lines[:, :, 5] == 0, torch.ones_like(lines[:, :, 5]), lines[:, :, 5])
    lines[:, :, 6] = torch.deg2rad(lines[:, :, 6])
    generators = gen[:, :, [0, 8, 9, 1, 5, 2]].clone()
    generators = torch.cat((generators, generators[:, :, 3:4]), dim=2)
    generators[:, :, [1, 2, 3, 5, 6]] /= base_mva
    if single:
        return buses[0], lines[0], generators[0]
    return buses, lines, generators


def pack_varying(buses: torch.Tensor, lines: torch.Tensor, generators: torch.Tensor):
    """Compact form of a packed batch of ONE case: only Pd,Qd | r,x,b,tau,shift | vg,Pg differ between the samples
    the reference generates (ref GNS/augment_grids.py:35-53); bus_i,type,Gs,Bs | f_bus,t_bus | bus_i,Pmax,Pmin,qg
    are constants of the case and Pg_set is a copy of Pg (ref GNS/utils.py:38).  Returns
    ``(bus_var [S,N,2], line_var [S,E,5], gen_var [S,Gn,2]), (bus_const [N,4], line_const [E,2], gen_const [Gn,4])``
    - 2N+5E+2Gn floats per grid instead of 6N+7E+7Gn (11.2 KB instead of 20.6 KB for case300).  Raises if the
    batch does not have that structure (``GNS.infer_host_compact`` expands it on the device)."""
    b, l, g = buses, lines, generators
    cb, cl, cg = b[0][:, [0, 1, 4, 5]], l[0][:, [0, 1]], g[0][:, [0, 1, 2, 5]]
    if not (torch.equal(b[:, :, [0, 1, 4, 5]], cb.expand(b.shape[0], -1, -1)) and
            torch.equal(l[:, :, [0, 1]], cl.expand(l.shape[0], -1, -1)) and
            torch.equal(g[:, :, [0, 1, 2, 5]], cg.expand(g.shape[0], -1, -1)) and torch.equal(g[:, :, 3], g[:, :, 6])):
        raise ValueError("pack_varying: the batch is not a set of perturbed samples of one case "
                         "(constant columns differ between grids, or Pg_set != Pg)")
    var = (b[:, :, 2:4].contiguous(), l[:, :, 2:7].contiguous(), g[:, :, [4, 6]].contiguous())
    return var, (cb.contiguous(), cl.contiguous(), cg.contiguous())


def expand_varying(var, const):
    """Host inverse of ``pack_varying`` (what ``gns_expand_inputs`` does on the device)."""
    (bv, lv, gv), (cb, cl, cg) = var, const
    S = bv.shape[0]
    b = torch.empty(S, cb.shape[0], 6); l = torch.empty(S, cl.shape[0], 7); g = torch.empty(S, cg.shape[0], 7)
    b[:, :, [0, 1, 4, 5]] = cb; b[:, :, 2:4] = bv
    l[:, :, [0, 1]] = cl; l[:, :, 2:7] = lv
    g[:, :, [0, 1, 2, 5]] = cg; g[:, :, 4] = gv[:, :, 0]; g[:, :, 3] = gv[:, :, 1]; g[:, :, 6] = gv[:, :, 1]
    return b, l, g


def expand_varying_device(var, const, device="cuda"):
    """``expand_varying`` on the device with the hand-written kernel behind ``gns_expand_inputs`` (include/gns_b200.h):
    compact device (or host) tensors -> the reference's packed rows on the GPU, e.g. to train on a compact data set."""
    from . import _lib
    lib = _lib.load_library()
    dev = torch.device(device)
    bv, lv, gv = (t.to(device=dev, dtype=torch.float32).contiguous() for t in var)
    cb, cl, cg = (t.to(device=dev, dtype=torch.float32).contiguous() for t in const)
    S, N, E, Gn = bv.shape[0], cb.shape[0], cl.shape[0], cg.shape[0]
    b = torch.empty(S, N, 6, device=dev); l = torch.empty(S, E, 7, device=dev); g = torch.empty(S, Gn, 7, device=dev)
    rc = lib.gns_expand_inputs(bv.data_ptr(), lv.data_ptr(), gv.data_ptr(), cb.data_ptr(), cl.data_ptr(), cg.data_ptr(), S, N, E,
                               Gn, b.data_ptr(), l.data_ptr(), g.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "gns_expand_inputs")
    return b, l, g


def renumber_buses(case: dict):
    """Map arbitrary external bus numbers (e.g. the real IEEE-300 table goes up to 9533) to the
    contiguous 1..N the path requires; the reference has no such step and raises IndexError at
    ``m[dst]`` (ref GNS/main.py:153, quirk Q8).  Returns (new case, external ids in internal order)."""
    bus = np.array(case["bus"], dtype=np.float64, copy=True)
    branch = np.array(case["branch"], dtype=np.float64, copy=True)
    gen = np.array(case["gen"], dtype=np.float64, copy=True)
    ext = bus[:, _BUS_I].astype(np.int64)
    if len(set(ext.tolist())) != len(ext):
        raise ValueError("duplicate bus numbers")
    lut = {int(e): i + 1 for i, e in enumerate(ext)}
    try:
        branch[:, _F] = [lut[int(x)] for x in branch[:, _F]]
        branch[:, _T] = [lut[int(x)] for x in branch[:, _T]]
        gen[:, _GBUS] = [lut[int(x)] for x in gen[:, _GBUS]]
    except KeyError as e:
        raise IndexError(f"branch / generator refers to unknown bus {e}") from None
    bus[:, _BUS_I] = np.arange(1, len(ext) + 1)
    out = dict(case)
    out.update(bus=bus, branch=branch, gen=gen)
    return out, ext


def augment_pack_device(case: dict, n_samples: int, seed: int = 0, device="cuda"):
    """`augment` + `pack_grids` fused on the device with torch ops: the perturbed,
    packed batch is produced directly in GPU memory, so building a large synthetic batch costs no
    host time or PCIe traffic.  Same recipe and packing transform; the random stream differs from
    the numpy generator of `augment` (the reference's own stream is unseeded anyway)."""
    dev = torch.device(device)
    gen_ = torch.Generator(device=dev).manual_seed(int(seed))
    S = int(n_samples)
    f64 = dict(dtype=torch.float64, device=dev)
    bus = torch.as_tensor(np.asarray(case["bus"]), **f64)
    branch = torch.as_tensor(np.asarray(case["branch"]), **f64)
    gen = torch.as_tensor(np.asarray(case["gen"]), **f64)
    N, E, Gn = bus.shape[0], branch.shape[0], gen.shape[0]
    base = float(case["baseMVA"])

    def U(lo, hi, *shape):
        return lo + (hi - lo) * torch.rand(*shape, generator=gen_, **f64)

    r = branch[:, _R] * U(0.9, 1.1, S, E)
    x = branch[:, _X] * U(0.9, 1.1, S, E)
    b = branch[:, _B] * U(0.9, 1.1, S, E)
    tap = U(0.8, 1.2, S, E)
    shift = U(-0.2, 0.2, S, E)
    vg = gen[:, _VG] * U(0.95, 1.05, S, Gn)
    span = gen[:, _PMAX] - gen[:, _PMIN]
    lo, hi = gen[:, _PMIN] + 0.25 * span, 0.75 * span
    pg = lo + (hi - lo) * torch.rand(S, Gn, generator=gen_, **f64)
    pd = bus[:, _PD] * U(0.5, 1.5, S, N)
    pd = pd * (pg.sum(1) / pd.sum(1))[:, None]
    qd = bus[:, _QD] * U(0.5, 1.5, S, N)
    f32 = torch.float32
    buses = torch.empty(S, N, 6, dtype=f32, device=dev)
    buses[:, :, 0] = bus[:, _BUS_I].to(f32)
    buses[:, :, 1] = bus[:, _BUS_TYPE].to(f32)
    buses[:, :, 2] = pd.to(f32) / base
    buses[:, :, 3] = qd.to(f32) / base
    buses[:, :, 4] = torch.tensor(1.0, dtype=f32, device=dev) / base
    buses[:, :, 5] = torch.tensor(-1.0, dtype=f32, device=dev) / base
    lines = torch.empty(S, E, 7, dtype=f32, device=dev)
    lines[:, :, 0] = branch[:, _F].to(f32)
    lines[:, :, 1] = branch[:, _T].to(f32)
    lines[:, :, 2], lines[:, :, 3], lines[:, :, 4] = r.to(f32), x.to(f32), b.to(f32)
    tap32 = tap.to(f32)
    lines[:, :, 5] = torch.where(tap32 == 0, torch.ones_like(tap32), tap32)
    lines[:, :, 6] = torch.deg2rad(shift.to(f32))
    gens = torch.empty(S, Gn, 7, dtype=f32, device=dev)
    gens[:, :, 0] = gen[:, _GBUS].to(f32)
    gens[:, :, 1] = gen[:, _PMAX].to(f32) / base
    gens[:, :, 2] = gen[:, _PMIN].to(f32) / base
    gens[:, :, 3] = pg.to(f32) / base
    gens[:, :, 4] = vg.to(f32)
    gens[:, :, 5] = gen[:, _QG].to(f32) / base
    gens[:, :, 6] = gens[:, :, 3]
    return buses, lines, gens


def make_batch(n_bus: int, n_samples: int, seed: int = 0, topo_seed: int = 0):
    """Synthetic load-perturbed batch of the named case, packed: (buses, lines, generators, label)."""
    case, label = get_case(n_bus, seed=topo_seed)
    aug = augment(case, n_samples, seed=seed)
    b, l, g = pack_grids(aug["bus"], aug["branch"], aug["gen"], aug["baseMVA"])
    return b, l, g, label
