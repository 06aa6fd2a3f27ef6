#!/usr/bin/env python
"""bench.py - GNS K-step message passing throughput on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic load-perturbed grids:
  forward  (inference, no checkpoints)           -> `value`, grids/s
  fwd+bwd  (training step: forward, backward of mean(total_loss), NCCL all-reduce of the
            flat gradient when N > 1; optimizer excluded)              -> `fwd_bwd.value`
Workload (config.workload): BASELINE.json configs[3]/[metric] - case300 (IEEE-sized synthetic
topology, 300/411/69), K=4, latent 20, hidden 10, multiple_phi, 65536 grids per GPU (weak
scaling: every rank processes its own batch, no data-path collective in inference).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference ...                           # CPU arm (oracle port, host cores)
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


def flops_per_grid(n_bus, n_line, K, L, H, multi=True):
    """Algorithmic MLP FLOPs of one forward: 2*K*(E*MAC_line + N*MAC_bus)."""
    mac_line = 3 * ((5 + L) * H + H * H + H * L) if multi else ((5 + L) * H + H * H + H)
    mac_bus = 2 * ((4 + 2 * L) * H + H * H + H) + ((4 + 2 * L) * H + H * H + H * L)
    return 2 * K * (n_line * mac_line + n_bus * mac_bus)


def io_bytes_per_grid(n_bus, n_line, n_gen):
    return 4 * (6 * n_bus + 7 * n_line + 7 * n_gen), 4 * (2 * n_bus + 2)


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed like the reference runs (per-sample Python loop,
# ref GNS/main.py:279-283), one single-threaded worker per host core.
# ------------------------------------------------------------------------------------------
def _cpu_worker(job):
    import torch
    torch.set_num_threads(1)
    from oracle import gns_oracle as orc
    params, buses, lines, gens, K, L, train = job
    t0 = time.perf_counter()
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
    return n, time.perf_counter() - t0


def cpu_reference_run(case, K, L, sample, train, workers):
    import multiprocessing as mp
    import torch
    import opf_graph_neural_solver_b200 as pkg
    from oracle import gns_oracle as orc
    params = orc.init_params(L, 10, K, True, seed=0)
    buses, lines, gens, label = pkg.data.make_batch(case, sample, seed=1)
    workers = max(1, min(workers, sample))
    per = (sample + workers - 1) // workers
    jobs = [(params, buses[i:i + per], lines[i:i + per], gens[i:i + per], K, L, train)
            for i in range(0, sample, per)]
    t0 = time.perf_counter()
    if len(jobs) == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res) / wall, wall, len(jobs), label


def _nr_worker(job):
    from oracle import newton_raphson as nr
    tables, idx = job
    t0 = time.perf_counter()
    res = nr.newton_pf_batch(tables, idx)
    return len(idx), sum(r[2] for r in res), time.perf_counter() - t0


def nr_baseline(case, sample, workers):
    """Restated Newton-Raphson (not pypower, which is not installable offline) on the same
    synthetic samples, one single-threaded worker per host core (ref GNS/evaluate.py:31-40)."""
    import multiprocessing as mp
    import opf_graph_neural_solver_b200 as pkg
    tables = pkg.data.augment(pkg.data.get_case(case)[0], sample, seed=1)
    workers = max(1, min(workers, sample))
    chunks = [list(range(i, sample, workers)) for i in range(workers)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(workers) as pool:
        res = pool.map(_nr_worker, [(tables, c) for c in chunks])
    wall = time.perf_counter() - t0
    n, conv = sum(r[0] for r in res), sum(r[1] for r in res)
    return {"value": n / wall, "unit": "grids/s", "cores": workers, "kind": "restated NR, not pypower",
            "sample": f"{n} grids, tol 1e-8, max 10 iterations, flat start", "converged": conv}


def run_reference_arm(args):
    """`--impl reference`: rank 0 only; other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    workers = os.cpu_count() or 1
    sample = args.cpu_sample or max(workers * 4, 64)
    vals = []
    for _ in range(args.warmup if args.cpu_sample is None else 0):
        cpu_reference_run(args.case, args.K, args.latent, min(sample, workers), False, workers)
    steps = args.steps if args.cpu_sample is None else 1
    for _ in range(steps):
        gps, wall, used, label = cpu_reference_run(args.case, args.K, args.latent, sample, False, workers)
        vals.append((gps, wall))
    gps = statistics.median(v[0] for v in vals)
    tr_gps, _, _, _ = cpu_reference_run(args.case, args.K, args.latent, max(sample // 2, used), True, workers)
    E, Gn = IEEE[args.case]
    line = {
        "impl": "reference", "metric": f"case{args.case}_K{args.K}_fwd_grids_per_s", "value": gps, "unit": "grids/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(v[1] for v in vals),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"case{args.case} ({label}) K={args.K} latent={args.latent} hidden=10 multiple_phi, "
                               f"{sample} grids per step (bounded sample of the 65536-grid batch)"},
        "fwd_bwd": {"value": tr_gps, "unit": "grids/s"},
        "cpu_baseline": {"value": gps, "unit": "grids/s", "cores": used, "kind": "port",
                         "sample": f"{sample} grids, per-sample loop like ref GNS/main.py:279-283, one 1-thread worker per core; "
                                   f"oracle port (the Python reference cannot travel to the GPU box)",
                         "fwd_bwd_value": tr_gps},
        "e2e": {"value": gps, "unit": "grids/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        line["cpu_baseline"]["newton_raphson"] = nr_baseline(args.case, min(sample, max(workers * 2, 32)), workers)
    except Exception as ex:  # pragma: no cover
        line["cpu_baseline"]["newton_raphson"] = {"value": None, "error": str(ex)}
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
    numa_bound = pkg.parallel.bind_to_gpu_numa_node(local) if world > 1 else False
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = pkg.load_library()

    case, K, L, S = args.case, args.K, args.latent, args.batch
    E, Gn = IEEE[case]
    torch.manual_seed(0)
    model = pkg.GNS(latent_dim=L, hidden_dim=10, K=K, gamma=0.9, multiple_phi=True).to(dev)
    model.validate_topology = False          # checked once below, outside the timed region
    base = min(S, 8192)
    b, l, g, label = pkg.data.make_batch(case, base, seed=1 + rank)
    rep = (S + base - 1) // base
    host = [t.repeat(rep, 1, 1)[:S].contiguous().pin_memory() for t in (b, l, g)]
    buses, lines, gens = (t.to(dev) for t in host)
    BLG = pkg.get_BLG()
    plan = model.plan_for(lines, gens, case)
    assert plan.matches(lines, gens)
    S_train = min(S, args.train_batch)
    tb, tl, tg = buses[:S_train], lines[:S_train], gens[:S_train]
    flat_grad = None

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

    def e2e_step():   # public API on host buffers: chunked H2D / kernel / D2H pipeline
        model.infer_host(host[0], host[1], host[2], out=out_host, chunk=8192)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
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
        return ms / steps

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_fwd = timed(fwd_step, args.steps, args.warmup)
    ms_train = timed(train_step, args.steps, args.warmup)
    ms_e2e = timed(e2e_step, max(2, args.steps // 2), 3)
    clocks = sampler.stop() if rank == 0 else None

    # dominant-kernel duration, measured live with CUDA events around the C-ABI forward call on the
    # launching stream (pack kernel + persistent kernel; the pack kernel is ~2 us)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.no_grad():
        fwd_step(); torch.cuda.synchronize()
        ev[0].record(); fwd_step(); ev[1].record(); torch.cuda.synchronize()
    kern_ms = ev[0].elapsed_time(ev[1])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    fl = flops_per_grid(case, E, K, L, 10)
    in_b, out_b = io_bytes_per_grid(case, E, Gn)
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
    ach = S * fl / (kern_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("forward_dram_bytes_per_launch")
    except Exception:
        pass
    info = plan.launch_info(S, K, L, 10, True)
    cpu = None
    if not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N=1 only
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--case", str(case),
                                "--K", str(K), "--latent", str(L), "--cpu-sample", str(args.cpu_sample or 256)],
                               capture_output=True, text=True, timeout=600,
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
                   "parallelism": f"dp{world} (batch sharded, no data-path collective)", "launch": info},
        "fwd_bwd": {"metric": f"case{case}_K{K}_fwd_bwd_grids_per_s", "value": gps_train, "unit": "grids/s",
                    "ms_per_step": ms_train, "grids_per_gpu_per_step": S_train,
                    "tflops_algorithmic": gps_train * 3 * fl / 1e12 / world,
                    "frac_of_fp32_peak": gps_train * 3 * fl / world / peak_ffma,
                    "includes": "forward with checkpoints, backward, gradient all-reduce (N>1); optimizer excluded"},
        "roofline": {"bound": "fp32_ffma", "achieved": ach, "peak": peak_ffma / 1e12, "unit": "TFLOP/s",
                     "frac": ach * 1e12 / peak_ffma, "traffic": traffic,
                     "peak_source": "measured here: max of the register-only FFMA and packed FFMA2 probes "
                                    "(MEASURED_PEAKS.json has no FP32 entry); theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5",
                     "kernel_ms": kern_ms, "flop_per_grid": fl,
                     "hbm": {"achieved_gbs": S * (in_b + out_b) / (kern_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                             "frac": S * (in_b + out_b) / (kern_ms * 1e-3) / 1e9 / hbm_peak, "of": "measured"}},
        "cpu_baseline": cpu,
        "e2e": {"value": gps_e2e, "unit": "grids/s", "h2d_bytes_per_step": S * in_b, "d2h_bytes_per_step": S * out_b,
                "ms_per_step": ms_e2e, "path": "pinned host tensors -> GNS.infer_host (8192-grid chunks, copy/compute overlap) -> pinned host outputs"},
        # our kernels per call: forward = pack, fuse, gns_forward; backward = gns_backward, reduce, gather, unfuse, unpack
        "gpu_launches": 3 * args.steps + 8 * args.steps + max(2, args.steps // 2) * 3 * ((S + 8191) // 8192),
        "clocks": clocks,
    }
    line["config"]["numa_bound"] = numa_bound
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
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
