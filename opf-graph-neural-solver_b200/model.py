"""``GNS`` - the reference's ``nn.Module`` surface on the B200 kernels.

Mirrors ref GNS/main.py:17-31 (``LearningBlock``) and :107-202 (``GNS``):

* constructor ``GNS(latent_dim=10, hidden_dim=10, K=30, gamma=0.9, multiple_phi=False)``,
  attributes ``multiple_phis``, ``latent_dim``, ``gamma``, ``K`` and the ``ModuleDict``s
  ``phi_v / phi_theta / phi_m`` (or ``phi``) and ``L_theta / L_v / L_m`` keyed ``str(k)``;
* ``state_dict`` keys ``"{net}.{k}.linear{1,2,4}.{weight,bias}"``, and the
  k-major construction order so ``torch.manual_seed(s)`` reproduces the reference init;
* ``forward(buses, lines, generators, B, L, G) -> (v, theta, total_loss, last_loss)``.

New capability: the three tensors may carry a leading batch dimension
(``[S,N,6] / [S,E,7] / [S,Gn,7]``, one shared topology); outputs are then ``[S,N]``,
``[S,N]``, ``[S]``, ``[S]``.  All arithmetic happens in the sm_100a kernels behind the C
ABI of ``include/gns_b200.h``; this file only owns parameters, shapes and autograd glue.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .plan import TopologyPlan


def get_BLG():
    """Column maps of the packed bus / line / generator tensors (ref GNS/utils.py:4-13)."""
    B = {"bus_i": 0, "type": 1, "Pd": 2, "Qd": 3, "Gs": 4, "Bs": 5}
    L = {"f_bus": 0, "t_bus": 1, "r": 2, "x": 3, "b": 4, "tau": 5, "theta": 6}
    G = {"bus_i": 0, "Pmax": 1, "Pmin": 2, "Pg_set": 3, "vg": 4, "qg": 5, "Pg": 6}
    return B, L, G


class LearningBlock(nn.Module):
    """Parameter container of one 3-layer MLP (ref GNS/main.py:17-31).  The layer names
    ``linear1, linear2, linear4`` are part of the checkpoint contract.  Its arithmetic runs
    inside the fused kernels; calling the block on its own is not part of the hot path."""

    def __init__(self, dim_in, hidden_dim, dim_out):
        super().__init__()
        self.linear1 = nn.Linear(dim_in, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.linear4 = nn.Linear(hidden_dim, dim_out)

    def forward(self, x):  # pragma: no cover - never used by GNS.forward
        raise RuntimeError("LearningBlock is evaluated inside the fused GNS kernels; call GNS.forward")


# kernels launched through the C ABI by this process (bench.py reports the count of its timed regions):
# gns_forward = pack + fuse + persistent forward; gns_backward = persistent backward + reduce + gather + unfuse + unpack
COUNTERS = {"kernels": 0, "forward_calls": 0, "backward_calls": 0}


def chunk_bounds(S: int, chunk: int):
    """[start, stop) of the chunks of the host pipeline: the first ones are small (chunk/8, chunk/8, chunk/4, chunk/2)
    so that the kernels start after a short first copy, the rest full size (per-chunk launch overhead amortised)."""
    bounds, a = [], 0
    for frac in (8, 8, 4, 2):
        step = max(1, chunk // frac)
        if S - a > chunk:
            bounds.append((a, a + step))
            a += step
    while a < S:
        bounds.append((a, min(S, a + chunk)))
        a += chunk
    return bounds


def _run_forward(module, plan, need_grad, buses, lines, gens, flat, const=None):
    """One ``gns_forward`` call on device tensors; returns (v, theta, total, last, workspace).
    ``const`` = (bus_const, gen_const) device tensors: the three inputs are then in the compact format
    (``gns_forward_compact``, inference only)."""
    lib = _lib.load_library()
    S, N = buses.shape[0], buses.shape[1]
    dev = buses.device
    K, Ld, Hd, multi = module.K, module.latent_dim, module.hidden_dim, int(module.multiple_phis)
    nbytes = lib.gns_workspace_bytes(plan.handle, S, K, Ld, Hd, multi, int(need_grad))
    if nbytes < 0:
        raise RuntimeError("gns_workspace_bytes: " + _lib.last_error())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(2 * S * N + 2 * S, dtype=torch.float32, device=dev)      # one allocation for the four outputs
    v, theta = out[:S * N].view(S, N), out[S * N:2 * S * N].view(S, N)
    total, last = out[2 * S * N:2 * S * N + S], out[2 * S * N + S:]
    stream = torch.cuda.current_stream(dev).cuda_stream
    if const is not None:
        assert not need_grad
        rc = lib.gns_forward_compact(plan.handle, flat.data_ptr(), buses.data_ptr(), lines.data_ptr(), gens.data_ptr(),
                                     const[0].data_ptr(), const[1].data_ptr(), S, K, Ld, Hd, multi, float(module.gamma),
                                     v.data_ptr(), theta.data_ptr(), total.data_ptr(), last.data_ptr(),
                                     ws.data_ptr(), nbytes, stream)
    else:
        rc = lib.gns_forward(plan.handle, flat.data_ptr(), buses.data_ptr(), lines.data_ptr(), gens.data_ptr(),
                             S, K, Ld, Hd, multi, float(module.gamma),
                             v.data_ptr(), theta.data_ptr(), total.data_ptr(), last.data_ptr(),
                             ws.data_ptr(), nbytes, int(need_grad), stream)
    _lib.check(rc, "gns_forward")
    COUNTERS["kernels"] += 3
    COUNTERS["forward_calls"] += 1
    return v, theta, total, last, ws


class _GNSFunction(torch.autograd.Function):
    """autograd glue: one ``gns_forward`` / one ``gns_backward`` call per step."""

    @staticmethod
    def forward(ctx, module, plan, need_grad, buses, lines, gens, flat, *params):
        v, theta, total, last, ws = _run_forward(module, plan, need_grad, buses, lines, gens, flat)
        if need_grad:
            ctx.module, ctx.plan, ctx.ws = module, plan, ws
            ctx.save_for_backward(buses, lines, gens, flat, v)
            ctx.shapes = [p.shape for p in params]     # empty in flat-leaf mode
        return v, theta, total, last

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_v, g_theta, g_total, g_last):
        lib = _lib.load_library()
        module, plan, ws = ctx.module, ctx.plan, ctx.ws
        buses, lines, gens, flat, v = ctx.saved_tensors
        S = buses.shape[0]
        dev = buses.device
        K, Ld, Hd, multi = module.K, module.latent_dim, module.hidden_dim, int(module.multiple_phis)

        def ptr(t):
            return None if t is None else t.contiguous().data_ptr()

        g_total = torch.zeros(S, dtype=torch.float32, device=dev) if g_total is None else g_total.contiguous()
        keep = [g_total]
        for t in (g_last, g_v, g_theta):
            keep.append(None if t is None else t.contiguous())
        if keep[2] is not None:   # v is clamped at 0 on output (ref GNS/main.py:201): no gradient where clamped
            keep[2] = torch.where(v > 0, keep[2], torch.zeros_like(keep[2]))
        grad_flat = torch.empty_like(flat)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.gns_backward(plan.handle, flat.data_ptr(), buses.data_ptr(), lines.data_ptr(), gens.data_ptr(),
                              S, K, Ld, Hd, multi, float(module.gamma),
                              keep[0].data_ptr(), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]),
                              grad_flat.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "gns_backward")
        COUNTERS["kernels"] += 5
        COUNTERS["backward_calls"] += 1
        if not ctx.shapes:      # flat-leaf mode: one gradient for the flat parameter buffer
            return (None, None, None, None, None, None, grad_flat)
        # Per-parameter gradients alias ONE flat buffer (a single all-reduce / fused Adam can use it), but
        # are handed to autograd as storage aliases rather than views: AccumulateGrad clones view
        # gradients, which cost one tiny copy kernel per parameter tensor (144 of them for K=4).
        grads, off, storage = [], 0, grad_flat.untyped_storage()
        base = grad_flat.storage_offset()
        for shp in ctx.shapes:
            n = shp.numel()
            g = torch.empty(0, dtype=torch.float32, device=dev).set_(storage, base + off, shp)
            grads.append(g)
            off += n
        module._last_grad_flat = grad_flat
        return (None, None, None, None, None, None, None) + tuple(grads)


class GNS(nn.Module):
    def __init__(self, latent_dim=10, hidden_dim=10, K=30, gamma=0.9, multiple_phi=False):
        super().__init__()
        self.multiple_phis = multiple_phi
        if self.multiple_phis:
            self.phi_v = nn.ModuleDict()
            self.phi_theta = nn.ModuleDict()
            self.phi_m = nn.ModuleDict()
        else:
            self.phi = nn.ModuleDict()
        self.L_theta = nn.ModuleDict()
        self.L_v = nn.ModuleDict()
        self.L_m = nn.ModuleDict()
        # k-major construction keeps the RNG stream of the reference ctor (ref GNS/main.py:124-134)
        for k in range(K):
            key = str(k)
            if self.multiple_phis:
                self.phi_v[key] = LearningBlock(5 + latent_dim, hidden_dim, latent_dim)
                self.phi_theta[key] = LearningBlock(5 + latent_dim, hidden_dim, latent_dim)
                self.phi_m[key] = LearningBlock(5 + latent_dim, hidden_dim, latent_dim)
            else:
                self.phi[key] = LearningBlock(5 + latent_dim, hidden_dim, 1)
            self.L_theta[key] = LearningBlock(4 + 2 * latent_dim, hidden_dim, 1)
            self.L_v[key] = LearningBlock(4 + 2 * latent_dim, hidden_dim, 1)
            self.L_m[key] = LearningBlock(4 + 2 * latent_dim, hidden_dim, latent_dim)
        self.latent_dim = latent_dim
        self.hidden_dim = hidden_dim
        self.gamma = gamma
        self.K = K
        # Kernels are instantiated for hidden_dim 10 and latent_dim 10 / 20 / 64 and K <= 64 (every value the
        # reference's own scripts use); fail here rather than at the first forward.  (Skipped when the library
        # has not been built yet: constructing / loading checkpoints needs no kernel.)
        try:
            lib = _lib.load_library()
        except RuntimeError:
            lib = None
        if lib is not None and (not lib.gns_dims_supported(int(latent_dim), int(hidden_dim)) or not 1 <= K <= 64):
            raise ValueError(f"GNS (B200 build): no sm_100a kernel for latent_dim={latent_dim}, hidden_dim={hidden_dim}, "
                             f"K={K}; built: hidden_dim=10, latent_dim in (10, 20, 64), 1 <= K <= 64. There is no "
                             "fallback path.")
        # --- not part of the reference surface ---
        self.validate_topology = True     # device-side check that a batch shares the plan's topology
        self._flat = None                 # flat float32 parameter storage, state_dict order
        self._param_list = None
        self._last_grad_flat = None       # flat gradient buffer of the most recent backward
        # Gradient delivery.  Default: ONE autograd leaf (the flat buffer); a post-accumulate hook then
        # exposes the result as per-parameter ``.grad`` aliases of one persistent flat gradient buffer.
        # A reference-style loop (128 forward calls, one backward, ref GNS/main.py:279-288) otherwise
        # pays 144 AccumulateGrad nodes per call.  Set True to put every Parameter into the autograd
        # graph instead (needed for per-parameter hooks, torch.autograd.grad w.r.t. parameters, DDP).
        self.per_parameter_autograd = False
        self._flat_leaf = None
        self._leaf_owner = None           # id() of the module that registered the leaf's hook (copies lose hooks)
        self._grad_flat = None
        self._grad_aliases = None
        self._topo_ok_key = None          # (ptr, version, shape) of the last device tensors that passed the topology check
        self._plans = {}                  # topology key -> TopologyPlan
        self._last_plan = None

    def __getstate__(self):
        # copies / pickles carry the parameters only: topology plans hold native handles, and the gradient leaf and
        # its aliases belong to this instance (a copied leaf has no hook); all of them are rebuilt on first use
        state = self.__dict__.copy()
        state.update(_plans={}, _last_plan=None, _flat_leaf=None, _leaf_owner=None, _grad_flat=None, _grad_aliases=None,
                     _topo_ok_key=None, _last_grad_flat=None)
        return state

    # ------------------------------------------------------------------ parameters
    def _flat_ok(self):
        f = self._flat
        if f is None:
            return False
        if self._param_list is None:
            self._param_list = [(p, p.numel()) for p in self.parameters()]
        base, off = f.data_ptr(), 0
        for p, n in self._param_list:      # a parameter that moved (.to / .cuda / re-assignment) has another address
            if p.data_ptr() != base + 4 * off:
                return False
            off += n
        return off == f.numel() and self._param_list[0][0].device == f.device

    def flatten_parameters(self):
        """Re-home every parameter as a view of one flat float32 buffer in ``state_dict`` order
        (what the C ABI consumes).  Called lazily; ``load_state_dict`` / in-place optimizer
        updates keep the views intact, ``.to()`` / ``.cuda()`` trigger a re-flatten."""
        params = list(self.parameters())
        dev = params[0].device
        with torch.no_grad():
            flat = torch.cat([p.detach().reshape(-1).to(torch.float32) for p in params])
            off = 0
            for p in params:
                n = p.numel()
                p.data = flat[off:off + n].view(p.shape)
                off += n
        self._flat = flat
        self._param_list = [(p, p.numel()) for p in params]
        self._flat_leaf = self._grad_flat = self._grad_aliases = None
        assert flat.device == dev
        return flat

    def flat_parameters(self) -> torch.Tensor:
        if not self._flat_ok():
            self.flatten_parameters()
        return self._flat

    def _leaf(self) -> torch.Tensor:
        """The flat buffer as the single autograd leaf of the fast gradient path."""
        # copy.deepcopy / pickling copy the cached leaf WITHOUT its hook (and the gradient aliases without their
        # owner): a copy therefore rebuilds them on first use instead of silently training nothing
        if self._flat_leaf is None or self._leaf_owner != id(self):
            leaf = self._flat.detach().requires_grad_(True)      # shares storage with the parameters
            leaf.register_post_accumulate_grad_hook(self._deliver_gradients)
            self._flat_leaf, self._leaf_owner = leaf, id(self)
            self._grad_flat = self._grad_aliases = None
        return self._flat_leaf

    def _deliver_gradients(self, leaf):
        g, leaf.grad = leaf.grad, None
        params = self._param_list
        if self._grad_flat is None:
            self._grad_flat = torch.empty_like(g)
            storage, off, aliases = self._grad_flat.untyped_storage(), 0, []
            for p, n in params:
                aliases.append(torch.empty(0, dtype=torch.float32, device=g.device).set_(storage, off, p.shape))
                off += n
            self._grad_aliases = aliases
        fresh = True
        for (p, _), alias in zip(params, self._grad_aliases):
            if p.grad is not None and p.grad.data_ptr() == alias.data_ptr():
                fresh = False
                break
        if fresh:
            self._grad_flat.copy_(g)
        else:
            self._grad_flat.add_(g)
        off = 0
        for (p, n), alias in zip(params, self._grad_aliases):
            if p.requires_grad:
                if p.grad is None:
                    p.grad = alias
                elif p.grad.data_ptr() != alias.data_ptr():
                    # a foreign gradient tensor (set by the user, or moved by Module.to()): accumulate into it
                    p.grad.add_(g[off:off + n].view(p.shape).to(p.grad.device))
            off += n
        self._last_grad_flat = self._grad_flat

    # ------------------------------------------------------------------ topology plans
    def plan_for(self, lines: torch.Tensor, generators: torch.Tensor, n_bus: int, host=None) -> TopologyPlan:
        """Plan of the batch's topology.  With ``validate_topology`` every grid of the batch is checked against it:
        on the host when the caller's tensors live there (``host`` = the caller's (lines, generators); reference-style
        per-sample loops then never synchronise the stream), else by one small kernel + flag read, skipped when the
        same device tensors (address, version, shape) already passed."""
        dev_index = lines.device.index if lines.device.index is not None else torch.cuda.current_device()

        def ok(plan):
            if not self.validate_topology:
                return True
            if host is not None:
                return plan.matches_host(*host)
            key = (lines.data_ptr(), lines._version, tuple(lines.shape), generators.data_ptr(), generators._version, id(plan))
            if key == self._topo_ok_key:
                return True
            if plan.matches(lines, generators):
                self._topo_ok_key = key
                return True
            return False

        plan = self._last_plan
        if plan is not None and plan.device == dev_index and \
                (plan.n_bus, plan.n_line, plan.n_gen) == (n_bus, lines.shape[1], generators.shape[1]):
            if ok(plan):
                return plan
        plan = TopologyPlan.from_tensors(lines, generators, n_bus, dev_index)
        plan = self._plans.setdefault((dev_index,) + plan.key(), plan)
        if not ok(plan):
            raise ValueError("all grids of a batch must share one topology (f_bus, t_bus, generator buses)")
        self._last_plan = plan
        return plan

    # ------------------------------------------------------------------ host-resident batches
    @torch.no_grad()
    def infer_host(self, buses, lines, generators, out=None, chunk=8192, device=None):
        """Inference on a HOST-resident batch (ideally pinned): the batch is cut into chunks and the
        host->device copy of chunk i+1, the kernel of chunk i and the device->host copy of chunk i-1
        run on three streams, so the end-to-end rate is max(PCIe, compute) instead of their sum.
        ``out`` = optional (v, theta, total_loss, last_loss) host tensors to fill (pinned for speed).
        Returns host tensors.  (Not in the reference: it has no batching.)"""
        return self._infer_pipeline((buses, lines, generators), None, out, chunk, device)

    @torch.no_grad()
    def infer_host_compact(self, var, const, out=None, chunk=8192, device=None):
        """Same pipeline on the compact format of ``data.pack_varying``: the host ships only the columns that vary
        between samples (Pd,Qd | r,x,b,tau,shift | vg,Pg: 11.2 KB instead of 20.6 KB per case300 grid) and one
        constant block per case; the forward kernel reads the compact rows directly (``gns_forward_compact``) and
        takes the constant columns from the per-case block."""
        return self._infer_pipeline(tuple(var), tuple(const), out, chunk, device)

    def _infer_pipeline(self, host_in, const, out, chunk, device):
        if not torch.cuda.is_available():
            raise RuntimeError("GNS (B200 build) needs a CUDA device: there is no CPU fallback path")
        lib = _lib.load_library()
        p0 = next(self.parameters())
        if p0.device.type != "cuda":
            self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))
            p0 = next(self.parameters())
        dev = p0.device
        S, N = host_in[0].shape[0], host_in[0].shape[1]
        E, Gn = host_in[1].shape[1], host_in[2].shape[1]
        if out is None:
            out = (torch.empty(S, N).pin_memory(), torch.empty(S, N).pin_memory(),
                   torch.empty(S).pin_memory(), torch.empty(S).pin_memory())
        with torch.cuda.device(dev):
            comp = torch.cuda.current_stream(dev)
            h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            NS = 3                                     # device staging slots
            dbuf = [[torch.empty((chunk,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev) for t in host_in]
                    for _ in range(NS)]
            cdev = None
            if const is not None:
                # The forward kernel reads the compact rows directly (gns_forward_compact); the constants of the case
                # are checked against the plan once: the plan is built from them.
                cb, cl, cg = (t.to(device=dev, dtype=torch.float32).contiguous() for t in const)
                cdev = (cb, cg)
                lines0 = torch.zeros(1, E, 7, dtype=torch.float32, device=dev); lines0[0, :, :2] = cl
                gens0 = torch.zeros(1, Gn, 7, dtype=torch.float32, device=dev); gens0[0, :, 0] = cg[:, 0]
                plan = self.plan_for(lines0, gens0, N)
            else:
                plan = None
            ready = [torch.cuda.Event() for _ in range(NS)]
            free = [torch.cuda.Event() for _ in range(NS)]
            keep = []
            start = torch.cuda.Event(); start.record(comp)
            h2d.wait_event(start)
            flat = self.flat_parameters()
            bad = torch.zeros(1, dtype=torch.int32, device=dev)     # set by the per-chunk topology checks
            for i, (a, b) in enumerate(chunk_bounds(S, chunk)):
                slot = i % NS
                with torch.cuda.stream(h2d):
                    if i >= NS:
                        h2d.wait_event(free[slot])
                    for dst, src in zip(dbuf[slot], host_in):
                        dst[:b - a].copy_(src[a:b], non_blocking=True)
                    ready[slot].record(h2d)
                comp.wait_event(ready[slot])
                d = [t[:b - a] for t in dbuf[slot]]
                if plan is None:
                    plan = self.plan_for(d[1], d[2], N)
                elif self.validate_topology and const is None:
                    plan.check_async(d[1], d[2], bad)           # every chunk, without stalling the pipeline
                    COUNTERS["kernels"] += 1
                res = _run_forward(self, plan, False, d[0], d[1], d[2], flat, const=cdev)[:4]
                free[slot].record(comp)          # the chunk's device buffers may be refilled
                done = torch.cuda.Event(); done.record(comp)
                d2h.wait_event(done)
                with torch.cuda.stream(d2h):
                    for dst, src in zip(out, res):
                        src.record_stream(d2h)
                        dst[a:b].copy_(src, non_blocking=True)
                keep.append(res)
            d2h.synchronize()
            if self.validate_topology and int(bad.item()):
                raise ValueError("all grids of a batch must share one topology (f_bus, t_bus, generator buses)")
        return out

    # ------------------------------------------------------------------ forward
    def forward(self, buses, lines, generators, B=None, L=None, G=None):
        """Same call as ref GNS/main.py:140.  ``B, L, G`` are accepted for compatibility; the
        column order is fixed exactly as in the reference, which hard-codes ``lines[:, 1]`` and
        ``lines[:, 2:]`` (ref GNS/main.py:153,155)."""
        if not torch.cuda.is_available():
            raise RuntimeError("GNS (B200 build) needs a CUDA device: there is no CPU fallback path")
        single = buses.dim() == 2
        if single:
            buses, lines, generators = buses[None], lines[None], generators[None]
        if buses.dim() != 3 or buses.shape[2] != 6 or lines.shape[2] != 7 or generators.shape[2] != 7:
            raise ValueError("expected buses [S,N,6], lines [S,E,7], generators [S,Gn,7]")
        if not (buses.shape[0] == lines.shape[0] == generators.shape[0]):
            raise ValueError("batch sizes of buses / lines / generators differ")
        if buses.requires_grad or lines.requires_grad or generators.requires_grad:
            raise NotImplementedError("gradients w.r.t. the grid tensors are not provided (the reference never uses them)")
        p0 = next(self.parameters())
        if p0.device.type != "cuda":
            self.to(torch.device("cuda", torch.cuda.current_device()))
            p0 = next(self.parameters())
        dev, in_dev = p0.device, buses.device

        def prep(t):
            return t.detach().to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()

        buses_d, lines_d, gens_d = prep(buses), prep(lines), prep(generators)
        with torch.cuda.device(dev):
            flat = self.flat_parameters()
            params = [p for p, _ in self._param_list]
            need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)

            def run(plan, b, l, g):
                if need_grad and not self.per_parameter_autograd:
                    return _GNSFunction.apply(self, plan, True, b, l, g, self._leaf())
                if need_grad:
                    return _GNSFunction.apply(self, plan, True, b, l, g, flat, *params)
                return _run_forward(self, plan, False, b, l, g, flat)[:4]     # inference: no autograd bookkeeping at all

            host = (lines, generators) if (in_dev.type == "cpu" and lines.dtype == torch.float32) else None
            try:
                plan = self.plan_for(lines_d, gens_d, buses_d.shape[1], host=host)
            except ValueError:
                plan = None
            if plan is not None:
                v, theta, total, last = run(plan, buses_d, lines_d, gens_d)
            else:
                # Heterogeneous batch (the reference rebuilds its index tensors per sample, ref GNS/main.py:153, so a
                # per-sample loop over different topologies just works there): group the grids by topology, run every
                # group as one batched call on its own plan, and put the results back in input order.
                S = buses_d.shape[0]
                keys = torch.cat([lines_d[:, :, :2].reshape(S, -1), gens_d[:, :, 0]], dim=1)
                _, inverse = torch.unique(keys, dim=0, return_inverse=True)
                parts, order = [], []
                for gi in range(int(inverse.max()) + 1):
                    idx = (inverse == gi).nonzero().flatten()
                    b, l, g = buses_d[idx], lines_d[idx], gens_d[idx]
                    parts.append(run(self.plan_for(l, g, b.shape[1]), b, l, g))
                    order.append(idx)
                back = torch.argsort(torch.cat(order))
                v, theta, total, last = (torch.cat([p[i] for p in parts])[back] for i in range(4))
        if in_dev != dev:
            v, theta, total, last = (t.to(in_dev) for t in (v, theta, total, last))
        if single:
            return v[0], theta[0], total[0], last[0]
        return v, theta, total, last
