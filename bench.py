#!/usr/bin/env python
"""bench.py - GNS K-step message passing throughput on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic load-perturbed grids:
  forward  (inference, no checkpoints)                                  -> `value`, grids/s
  fwd+bwd  (training step: forward, backward of mean(total_loss), NCCL all-reduce of the flat
            gradient when N > 1; optimizer excluded)                    -> `fwd_bwd`, `roofline.fwd_bwd_frac`
  end to end (pinned host buffers -> public API -> pinned host outputs) -> `e2e`
Workload (config.workload): BASELINE.json configs[3] - case300 (IEEE-sized synthetic topology 300/411/69), K=4,
latent 20, hidden 10, multiple_phi, 65536 grids per GPU (weak scaling: every rank its own batch, no data-path
collective in inference).  The line also carries what BASELINE.json states beside it: `strong` (configs[2]: case118
fwd+bwd with 16384 grids IN TOTAL split over the ranks, all-reduce inside the timed region; configs[3]: case300
forward with 65536 grids in total), `stress` (configs[4]: case300 K=8 latent 64 fwd+bwd, batch 32768, N=1 only) and,
for N > 1, `grad_parity` (all-reduced gradient of N shards against the 1-GPU gradient of the whole batch).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference ...                           # CPU arm (host cores)
    torchrun ... bench.py --gpus N ...                             # one rank per GPU

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IEEE = {14: (20, 5), 30: (41, 6), 118: (186, 54), 300: (411, 69)}
REF_PER_CORE = 77.0     # grids/s/core of the real reference on case300-sized grids, forward, 1 thread


def flops_per_grid(n_bus, n_line, K, L, H, multi=True):
    """Algorithmic MLP FLOPs of one forward: 2*K*(E*MAC_line + N*MAC_bus)."""
    mac_line = 3 * ((5 + L) * H + H * H + H * L) if multi else ((5 + L) * H + H * H + H)
    mac_bus = 2 * ((4 + 2 * L) * H + H * H + H) + ((4 + 2 * L) * H + H * H + H * L)
    return 2 * K * (n_line * mac_line + n_bus * mac_bus)


def io_bytes_per_grid(n_bus, n_line, n_gen):
    return 4 * (6 * n_bus + 7 * n_line + 7 * n_gen), 4 * (2 * n_bus + 2)


def compact_bytes_per_grid(n_bus, n_line, n_gen):
    return 4 * (2 * n_bus + 5 * n_line + 2 * n_gen)


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed like the reference runs (per-sample Python loop,
# ref GNS/main.py:279-283), one single-threaded worker per host core, PERSISTENT pool.
# ------------------------------------------------------------------------------------------
_W = {}


def _cpu_init():
    import torch
    torch.set_num_threads(1)
    from oracle import gns_oracle  # noqa: F401  (imported once per worker, outside every timed region)


def _cpu_job(job):
    import torch
    from oracle import gns_oracle as orc
    params, buses, lines, gens, K, L, train = job
    n = buses.shape[0]
    if train:
        leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        losses = []
        for i in range(n):
            out = orc.gns_forward(leaves, buses[i], lines[i], gens[i], K=K, latent_dim=L, gamma=0.9, multiple_phi=True)
            losses.append(out[2])
        torch.stack(losses).mean().backward()
    else:
        with torch.no_grad():
            for i in range(n):
                orc.gns_forward(params, buses[i], lines[i], gens[i], K=K, latent_dim=L, gamma=0.9, multiple_phi=True)
    return n


def _nr_job(job):
    from oracle import newton_raphson as nr
    tables, idx = job
    res = nr.newton_pf_batch(tables, idx)
    return len(idx), sum(int(r[2]) for r in res)


class CpuArm:
    """Persistent worker pool (created and warmed before any clock starts) over the bounded sample."""

    def __init__(self, case, K, L, sample, workers):
        import multiprocessing as mp
        import opf_graph_neural_solver_b200 as pkg
        from oracle import gns_oracle as orc
        self.case, self.K, self.L, self.sample = case, K, L, sample
        self.workers = max(1, min(workers, sample))
        self.params = orc.init_params(L, 10, K, True, seed=0)
        self.buses, self.lines, self.gens, self.label = pkg.data.make_batch(case, sample, seed=1)
        self.tables = pkg.data.augment(pkg.data.get_case(case)[0], sample, seed=1, nominal_taps=True)
        self.pool = mp.get_context("fork").Pool(self.workers, initializer=_cpu_init)
        per = (sample + self.workers - 1) // self.workers
        self.slices = [(i, min(sample, i + per)) for i in range(0, sample, per)]
        self.pool.map(_cpu_job, [self._job(a, min(b, a + 1), False) for a, b in self.slices])   # warm every worker

    def _job(self, a, b, train):
        return (self.params, self.buses[a:b], self.lines[a:b], self.gens[a:b], self.K, self.L, train)

    def run(self, train):
        t0 = time.perf_counter()
        n = sum(self.pool.map(_cpu_job, [self._job(a, b, train) for a, b in self.slices]))
        wall = time.perf_counter() - t0
        return n / wall, wall

    def newton_raphson(self):
        """Restated Newton-Raphson (not pypower, which is not installable offline; ref GNS/evaluate.py:31-40) on the
        same topology and load / generation perturbation with NOMINAL taps (the reference's recipe draws off-nominal
        taps for every line, ref GNS/augment_grids.py:43-45, which leaves no solvable power flow on a meshed grid)."""
        chunks = [list(range(a, b)) for a, b in self.slices]
        t0 = time.perf_counter()
        res = self.pool.map(_nr_job, [(self.tables, c) for c in chunks])
        wall = time.perf_counter() - t0
        n, conv = sum(r[0] for r in res), sum(r[1] for r in res)
        return {"value": n / wall, "unit": "grids/s", "cores": self.workers, "kind": "restated NR, not pypower",
                "sample": f"{n} grids (nominal taps), tol 1e-8, max 10 iterations, flat start", "converged": conv,
                "all_converged": conv == n}

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(args):
    """`--impl reference`: rank 0 only; other ranks exit 0 without work."""
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # the other ranks' interpreter start-up (import torch on a fresh box) must not run under rank 0's clock: meet at
        # a gloo barrier first, then let them leave
        import torch.distributed as dist
        dist.init_process_group("gloo")
        dist.barrier()
        dist.destroy_process_group()
    if int(os.environ.get("RANK", "0")) != 0:
        return
    try:
        # all host cores the container allows (the GPU arm pins itself to the GPU's NUMA node before it runs this leg as a
        # child process; the CPU arm gets every core back)
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
        workers = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):  # pragma: no cover
        workers = os.cpu_count() or 1
    sample = max(args.cpu_sample or 256, 256)
    arm = CpuArm(args.case, args.K, args.latent, sample, workers)
    for _ in range(args.warmup):
        arm.run(False)
    vals = [arm.run(False) for _ in range(args.steps)]
    gps = statistics.median(v[0] for v in vals)
    tr = [arm.run(True) for _ in range(max(1, min(args.steps, 3)))]
    tr_gps = statistics.median(v[0] for v in tr)
    try:
        nrb = arm.newton_raphson()
    except Exception as ex:  # pragma: no cover
        nrb = {"value": None, "error": str(ex)}
    arm.close()
    cb = {"value": gps, "unit": "grids/s", "cores": arm.workers, "kind": "port",
          "sample": f"{sample} grids per step, per-sample loop like ref GNS/main.py:279-283, one 1-thread worker per core, "
                    f"persistent pool warmed before the clock; oracle port (the Python reference cannot travel to the GPU box)",
          "per_core": gps / arm.workers, "reference_per_core_survey": REF_PER_CORE,
          "fwd_bwd_value": tr_gps, "fwd_bwd_per_core": tr_gps / arm.workers, "newton_raphson": nrb}
    line = {
        "impl": "reference", "metric": f"case{args.case}_K{args.K}_fwd_grids_per_s", "value": gps, "unit": "grids/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * statistics.median(v[1] for v in vals), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"case{args.case} ({arm.label}) K={args.K} latent={args.latent} hidden=10 multiple_phi, "
                               f"{sample} grids per step (bounded sample of the 65536-grid batch)"},
        "fwd_bwd": {"value": tr_gps, "unit": "grids/s"},
        "cpu_baseline": cb,
        "e2e": {"value": gps, "unit": "grids/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import opf_graph_neural_solver_b200 as pkg
    from opf_graph_neural_solver_b200 import model as gmodel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU for the default arm (there is no CPU fallback path)")
    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 meanwhile (NCCL prints its version
    # banner there) is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    numa_bound = pkg.parallel.bind_to_gpu_numa_node(local)      # pinned buffers on the GPU's NUMA node (host copies)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = pkg.load_library()
    BLG = pkg.get_BLG()
    counters = gmodel.COUNTERS

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """W warm-up steps, then exactly `steps` steps between barrier + synchronize, CUDA events, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        k0 = counters["kernels"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        timed.launches += counters["kernels"] - k0
        return ms / steps

    timed.launches = 0

    def make_model(K, L):
        torch.manual_seed(0)
        m = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=True).to(dev)
        m.validate_topology = False          # checked once per workload below, outside the timed regions
        return m

    def device_batch(case, S, seed):
        base = min(S, 8192)
        b, l, g, label = pkg.data.make_batch(case, base, seed=seed)
        rep = (S + base - 1) // base
        host = [t.repeat(rep, 1, 1)[:S].contiguous() for t in (b, l, g)]
        return host, label

    # ---------------- headline workload: configs[3], weak scaling ----------------
    case, K, L, S = args.case, args.K, args.latent, args.batch
    E, Gn = IEEE[case]
    model = make_model(K, L)
    host, label = device_batch(case, S, 1 + rank)
    host = [t.pin_memory() for t in host]
    buses, lines, gens = (t.to(dev) for t in host)
    plan = model.plan_for(lines, gens, case)
    assert plan.matches(lines, gens)
    var, const = pkg.data.pack_varying(*host)
    var = tuple(t.pin_memory() for t in var)
    S_train = min(S, args.train_batch)
    tb, tl, tg = buses[:S_train], lines[:S_train], gens[:S_train]

    def fwd_step():
        with torch.no_grad():
            return model(buses, lines, gens, *BLG)

    def train_step():
        model.zero_grad(set_to_none=True)
        out = model(tb, tl, tg, *BLG)
        (out[2].sum() / (S_train * world)).backward()
        if world > 1:   # one all-reduce of the flat gradient (the views share one buffer)
            pkg.parallel.allreduce_gradients(model.parameters())
        return out

    out_host = [torch.empty(S, case).pin_memory(), torch.empty(S, case).pin_memory(),
                torch.empty(S).pin_memory(), torch.empty(S).pin_memory()]

    def e2e_step():        # public API on host buffers, compact input format: chunked H2D / expand / kernel / D2H pipeline
        model.infer_host_compact(var, const, out=out_host, chunk=args.e2e_chunk)

    def e2e_full_step():   # same with the reference's full rows
        model.infer_host(host[0], host[1], host[2], out=out_host, chunk=args.e2e_chunk)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_fwd = timed(fwd_step, args.steps, args.warmup)
    ms_train = timed(train_step, args.steps, args.warmup)
    model.zero_grad(set_to_none=True)
    ms_e2e = timed(e2e_step, max(2, args.steps // 2), 3)
    ms_e2e_full = timed(e2e_full_step, max(2, args.steps // 2), 3)
    clocks = sampler.stop() if rank == 0 else None

    # dominant-kernel durations, measured live with CUDA events around the C-ABI calls on the launching stream
    # (forward call = pack + fuse + persistent kernel, the first two ~2 us; training = the two persistent kernels + folds)
    def one(fn):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        fn(); torch.cuda.synchronize()
        ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1])
    kern_ms = one(fwd_step)
    kern_train_ms = one(train_step)
    model.zero_grad(set_to_none=True)

    # ---------------- pinned H2D ceiling of this box with all ranks copying at once ----------------
    probe_bytes = 512 << 20
    ph = torch.empty(probe_bytes, dtype=torch.uint8).pin_memory()
    pd = torch.empty(probe_bytes, dtype=torch.uint8, device=dev)
    pd.copy_(ph, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        pd.copy_(ph, non_blocking=True)
    e1.record()
    barrier()
    h2d_gbs = 4 * probe_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    # ... and with the results travelling the other way at the same time, in the workload's byte ratio (the copies of
    # the two directions share the link: measured here, they do not simply overlap)
    in_g, out_g = compact_bytes_per_grid(case, E, Gn), io_bytes_per_grid(case, E, Gn)[1]
    n_out = int(probe_bytes * out_g / in_g) // 16 * 16
    qh = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    qd = torch.empty(n_out, dtype=torch.uint8, device=dev)
    s_out = torch.cuda.Stream(dev)
    barrier()
    e0.record()
    s_out.wait_event(e0)
    for _ in range(4):
        pd.copy_(ph, non_blocking=True)
        with torch.cuda.stream(s_out):
            qh.copy_(qd, non_blocking=True)
    e2 = torch.cuda.Event(enable_timing=True)
    e2.record(s_out)
    torch.cuda.current_stream(dev).wait_event(e2)
    e1.record()
    barrier()
    bidir_grids_per_s = 4 * (probe_bytes / in_g) / (e0.elapsed_time(e1) * 1e-3)      # this rank: grids/s the link can carry
    del qh, qd
    if world > 1:
        t = torch.tensor([h2d_gbs, bidir_grids_per_s], device=dev)
        tl_ = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl_, t)
        h2d_all = [float(x[0]) for x in tl_]
        bidir_all = [float(x[1]) for x in tl_]
    else:
        h2d_all, bidir_all = [h2d_gbs], [bidir_grids_per_s]
    del ph, pd

    # ---------------- strong scaling of the stated totals (BASELINE.json configs[2], configs[3]) ----------------
    strong = {}
    lo, hi = pkg.parallel.shard_range(args.strong_fwd_total, rank, world)
    sb, sl, sg = buses[:hi - lo], lines[:hi - lo], gens[:hi - lo]

    def strong_fwd():
        with torch.no_grad():
            model(sb, sl, sg, *BLG)
    ms = timed(strong_fwd, args.steps, args.warmup)
    strong["configs3_case300_fwd"] = {"total_grids": args.strong_fwd_total, "grids_per_gpu": hi - lo, "ms_per_step": ms,
                                      "value": args.strong_fwd_total / (ms * 1e-3), "unit": "grids/s"}
    c2 = 118
    m118 = make_model(4, 20)
    h118, label118 = device_batch(c2, args.strong_train_total, 11)
    lo, hi = pkg.parallel.shard_range(args.strong_train_total, rank, world)
    b118, l118, g118 = (t[lo:hi].to(dev) for t in h118)
    p118 = m118.plan_for(l118, g118, c2)
    assert p118.matches(l118, g118)

    def strong_train():
        m118.zero_grad(set_to_none=True)
        out = m118(b118, l118, g118, *BLG)
        (out[2].sum() / args.strong_train_total).backward()
        if world > 1:
            pkg.parallel.allreduce_gradients(m118.parameters())
    ms = timed(strong_train, args.steps, args.warmup)
    fl118 = flops_per_grid(c2, IEEE[c2][0], 4, 20, 10)
    strong["configs2_case118_fwd_bwd"] = {"total_grids": args.strong_train_total, "grids_per_gpu": hi - lo, "ms_per_step": ms,
                                          "value": args.strong_train_total / (ms * 1e-3), "unit": "grids/s",
                                          "workload": f"case118 ({label118}) K=4 latent=20, gradient all-reduce inside the timed region",
                                          "tflops_algorithmic_per_gpu": args.strong_train_total / world / (ms * 1e-3) * 3 * fl118 / 1e12}
    del m118, b118, l118, g118

    # ---------------- multi-GPU gradient parity: N shards + all-reduce == 1 GPU on the whole batch ----------------
    grad_parity = None
    if world > 1:
        Sp = 1024 * world
        hp, _ = device_batch(case, Sp, 77)          # same seed on every rank: identical global batch
        mp_ = make_model(K, L)
        lo, hi = pkg.parallel.shard_range(Sp, rank, world)
        mp_.zero_grad(set_to_none=True)
        out = mp_(hp[0][lo:hi].to(dev), hp[1][lo:hi].to(dev), hp[2][lo:hi].to(dev), *BLG)
        (out[2].sum() / Sp).backward()
        pkg.parallel.allreduce_gradients(mp_.parameters())
        g_sharded = torch.cat([p.grad.reshape(-1) for p in mp_.parameters()]).clone()
        if rank == 0:
            mp_.zero_grad(set_to_none=True)
            out = mp_(hp[0].to(dev), hp[1].to(dev), hp[2].to(dev), *BLG)
            out[2].mean().backward()
            g_single = torch.cat([p.grad.reshape(-1) for p in mp_.parameters()])
            gmax = float(g_single.abs().max())
            diff = float((g_sharded - g_single).abs().max())
            grad_parity = {"grids": Sp, "max_abs_diff": diff, "max_abs_grad": gmax, "rel": diff / gmax,
                           "ok": diff <= 1e-5 * gmax, "bound": "1e-5 * max|g|"}
        barrier()
        del mp_

    # ---------------- stress config: case300 K=8 latent 64 fwd+bwd, batch 32768 (N = 1 only) ----------------
    stress = None
    if world == 1 and not args.no_stress:
        Ks, Ls, Ss, micro = 8, 64, args.stress_batch, args.stress_micro
        ms64 = make_model(Ks, Ls)
        nb = (Ss + micro - 1) // micro

        def stress_step():     # one training step of batch Ss in micro-batches (the checkpoints of 32768 grids are 79 GB)
            ms64.zero_grad(set_to_none=True)
            for i in range(nb):
                a, b_ = i * micro, min(Ss, (i + 1) * micro)
                out = ms64(buses[a:b_], lines[a:b_], gens[a:b_], *BLG)
                (out[2].sum() / Ss).backward()          # gradients accumulate into the flat buffer
        try:
            ms = timed(stress_step, 2, 1)
            fls = flops_per_grid(case, E, Ks, Ls, 10)
            stress = {"workload": f"case300 K={Ks} latent={Ls} multiple_phi fwd+bwd, batch {Ss} in {nb} micro-batches of {micro}",
                      "ms_per_step": ms, "value": Ss / (ms * 1e-3), "unit": "grids/s",
                      "tflops_algorithmic": Ss / (ms * 1e-3) * 3 * fls / 1e12}
        except RuntimeError as ex:  # pragma: no cover
            stress = {"error": str(ex)[:200]}
        del ms64

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    fl = flops_per_grid(case, E, K, L, 10)
    in_b, out_b = io_bytes_per_grid(case, E, Gn)
    cin_b = compact_bytes_per_grid(case, E, Gn)
    peak_ffma = max(lib.gns_measure_ffma_flops(local, 20000), lib.gns_measure_ffma2_flops(local, 20000))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    gps_fwd = world * S / (ms_fwd * 1e-3)
    gps_train = world * S_train / (ms_train * 1e-3)
    gps_e2e = world * S / (ms_e2e * 1e-3)
    gps_e2e_full = world * S / (ms_e2e_full * 1e-3)
    ach = S * fl / (kern_ms * 1e-3) / 1e12
    ach_train = S_train * 3 * fl / (kern_train_ms * 1e-3) / 1e12
    if stress and "tflops_algorithmic" in stress:
        stress["frac_of_fp32_peak"] = stress["tflops_algorithmic"] * 1e12 / peak_ffma
    for v in strong.values():
        if "tflops_algorithmic_per_gpu" in v:
            v["frac_of_fp32_peak"] = v["tflops_algorithmic_per_gpu"] * 1e12 / peak_ffma
    traffic_ncu = None
    try:
        traffic_ncu = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    info = plan.launch_info(S, K, L, 10, True)
    info_b = plan.launch_info(S_train, K, L, 10, True, backward=True)
    h2d_ceiling = sum(h2d_all) * 1e9 / cin_b
    cpu = None
    if not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N=1 only; same code path as --impl reference
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--case", str(case),
                                "--K", str(K), "--latent", str(L), "--cpu-sample", str(args.cpu_sample or 256),
                                "--steps", "3", "--warmup", "1"],
                               capture_output=True, text=True, timeout=900,
                               env={**os.environ, "CUDA_VISIBLE_DEVICES": "", "RANK": "0", "WORLD_SIZE": "1"})
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:  # pragma: no cover
            cpu = {"value": None, "unit": "grids/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    line = {
        "metric": f"case{case}_K{K}_fwd_grids_per_s", "value": gps_fwd, "unit": "grids/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_fwd, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"case{case} ({label}) K={K} latent={L} hidden=10 multiple_phi gamma=0.9, "
                               f"{S} grids per GPU per step (BASELINE.json configs[3])",
                   "l2": "inputs larger than L2 (%.0f MB per step)" % (S * in_b / 1e6),
                   "parallelism": f"dp{world} (batch sharded, no data-path collective)", "launch": info,
                   "launch_backward": info_b, "numa_bound": numa_bound},
        "fwd_bwd": {"metric": f"case{case}_K{K}_fwd_bwd_grids_per_s", "value": gps_train, "unit": "grids/s",
                    "ms_per_step": ms_train, "grids_per_gpu_per_step": S_train,
                    "tflops_algorithmic": gps_train * 3 * fl / 1e12 / world,
                    "frac_of_fp32_peak": gps_train * 3 * fl / world / peak_ffma,
                    "includes": "forward with checkpoints, backward, gradient all-reduce (N>1); optimizer excluded"},
        "roofline": {"bound": "fp32_ffma", "achieved": ach, "peak": peak_ffma / 1e12, "unit": "TFLOP/s",
                     "frac": ach * 1e12 / peak_ffma, "traffic": None,
                     "traffic_ncu": traffic_ncu,
                     "fwd_bwd_achieved": ach_train, "fwd_bwd_frac": ach_train * 1e12 / peak_ffma,
                     "fwd_bwd_kernel_ms": kern_train_ms,
                     "peak_source": "measured here: max of the register-only FFMA and packed FFMA2 probes "
                                    "(MEASURED_PEAKS.json has no FP32 entry); theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5",
                     "kernel_ms": kern_ms, "flop_per_grid": fl,
                     "flop_note": "flop_per_grid is the reference's algorithmic MLP count (SURVEY.md 8a/8d); the kernels execute "
                                  "0.62 of those MACs on case300 K=4 (0.68 before the dead net was dropped): fused W4^T.W1 block, receiver latent hoisted out of the line loop, and "
                                  "the last step's m net (its result is never read, ref main.py:176-202) is not evaluated",
                     "hbm": {"achieved_gbs": S * (in_b + out_b) / (kern_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                             "frac": S * (in_b + out_b) / (kern_ms * 1e-3) / 1e9 / hbm_peak, "of": "measured"}},
        "strong": strong,
        "stress": stress,
        "grad_parity": grad_parity,
        "cpu_baseline": cpu,
        "e2e": {"value": gps_e2e, "unit": "grids/s", "h2d_bytes_per_step": S * cin_b, "d2h_bytes_per_step": S * out_b,
                "ms_per_step": ms_e2e,
                "path": "pinned host tensors in the compact format (data.pack_varying: Pd,Qd | r,x,b,tau,shift | vg,Pg) -> "
                        "GNS.infer_host_compact (%d-grid chunks:" % args.e2e_chunk + " H2D, gns_forward_compact, D2H on three streams) -> "
                        "pinned host outputs",
                "h2d_gbs_per_gpu_concurrent": h2d_all, "h2d_ceiling_grids_per_s": h2d_ceiling,
                "frac_of_h2d_ceiling": gps_e2e / h2d_ceiling,
                "link_ceiling_grids_per_s": sum(bidir_all), "frac_of_link_ceiling": gps_e2e / sum(bidir_all),
                "link_ceiling_note": "all ranks copying a grid's compact inputs host->device and its outputs device->host at the "
                                     "same time (pinned, 512 MiB probes): the ceiling of ANY pipeline on this box",

                "full_rows": {"value": gps_e2e_full, "ms_per_step": ms_e2e_full, "h2d_bytes_per_step": S * in_b,
                              "path": "GNS.infer_host on the reference's packed rows",
                              "frac_of_h2d_ceiling": gps_e2e_full / (sum(h2d_all) * 1e9 / in_b)}},
        # kernels launched through the C ABI inside the timed regions, counted at the call sites (model.COUNTERS):
        # forward = pack, fuse, gns_forward; backward = gns_backward, reduce, gather, unfuse, unpack; + topology check per chunk (full rows)
        "gpu_launches": timed.launches,
        "clocks": clocks,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--case", type=int, default=300, choices=sorted(IEEE))
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--latent", type=int, default=20)
    ap.add_argument("--batch", type=int, default=65536, help="grids per GPU per forward step")
    ap.add_argument("--train-batch", type=int, default=16384, help="grids per GPU per training step")
    ap.add_argument("--strong-fwd-total", type=int, default=65536, help="configs[3]: total grids split over the ranks")
    ap.add_argument("--strong-train-total", type=int, default=16384, help="configs[2]: total case118 grids split over the ranks")
    ap.add_argument("--stress-batch", type=int, default=32768)
    ap.add_argument("--stress-micro", type=int, default=8192)
    ap.add_argument("--no-stress", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=4096, help="grids per chunk of the host pipeline")
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
