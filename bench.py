#!/usr/bin/env python
"""Benchmark of the LP_MP message-passing hot path (contract: see the task prompt / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C3|C2|C4|C5|C5h|C1] [--dtype f32|f64]
                  [--scaling strong|weak] [--coupling halo|halo_nccl|pipelined] [--no-extras]

A "step" is ONE ITERATION of the hot path = forward pass + backward pass (LP::ComputePass, include/LP_MP.h:869-887)
+ the lower-bound reduction (LP::LowerBound, :1507-1518) over one synthetic MRF.  Metric: pairwise message updates/s
(anisotropic: one receive + one send per edge per pass => 4E per iteration).

Default workload at every N: BASELINE.json configs[2] (C3), the 4096x4096 Potts grid with 32 labels the north star quotes the
1 -> 8 GPU scaling on.  N > 1 cuts the SAME grid into N row stripes (strong scaling; --scaling weak keeps 4096 rows per GPU).
At N = 1 the line also carries roofline / end-to-end blocks of the other single-GPU configurations (C2, C4, C5h and C3 in f64)
under "workloads", so that one driver run covers them.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (H, W, L, pairwise, p_lo, p_hi, description)
    "C1": (64, 64, 8, "potts", 0.1, 1.0, "64x64 Potts grid, 8 labels (BASELINE config 1)"),
    "C2": (1024, 1024, 16, "trunclin", 0.25, 1.0, "1024x1024 truncated-linear (stereo-like) grid, 16 labels (BASELINE config 2)"),
    "C3": (4096, 4096, 32, "potts", 0.1, 1.0, "4096x4096 Potts grid, 32 labels (BASELINE config 3)"),
    "C5": (2048, 2048, 64, "dense", 0.0, 1.0, "2048x2048 grid, dense 64x64 pairwise tables (BASELINE config 5)"),
    # half of C5's rows: same kernel and per-row behaviour in 72 GB instead of 143 GB of HBM
    "C5h": (1024, 2048, 64, "dense", 0.0, 1.0, "1024x2048 grid, dense 64x64 pairwise tables (upper half of BASELINE config 5)"),
}
GRAPH_WORKLOADS = {
    # name: (nodes, degree, L, description)
    "C4": (1_000_000, 3, 8, "random 3-regular graph, 1M nodes, 8 labels, dense tables, graph-colouring schedule (BASELINE config 4)"),
}
SEED = 0x5EED
METRIC = "pairwise msg updates/s on grid MRF"
UNIT = "msg_updates/s"
TTB_ITERS = {"C1": 1000, "C2": 400, "C3": 300}


def edges_of(H, W):
    return H * (W - 1) + (H - 1) * W


def algorithmic_bytes_per_pass(H, W, L, pairwise, s):
    P = {"dense": L * L, "potts": 1, "trunclin": 2}[pairwise]
    N, E = H * W, edges_of(H, W)
    return s * (2.0 * N * L + E * (P + 5.0 * L))


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def imported_traffic(key):
    """DRAM bytes per launch of the pass kernel from the committed `ncu --set full` capture (profiles/traffic.json): ncu
    cannot run inside a timed bench, so the figure is IMPORTED, not measured in this run -- the line says so."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def stereo_unaries(H, W, L, seed=SEED):
    """stereo-like data term |x - d(r,c)| (slope 0.5, truncated at 2) + noise; the true label d drifts
    smoothly along x (and slowly along y); 10 % of the pixels are outliers with U[0,2) costs.  float32 [H][W][L]."""
    import numpy as np
    rng = np.random.default_rng(seed)
    r, c = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    d = 0.5 * (L - 1) * (1.0 + np.sin(2.0 * np.pi * (3.0 * c / W + 0.5 * r / H)))
    x = np.arange(L, dtype=np.float32)
    theta = np.minimum(0.5 * np.abs(x[None, None, :] - d[..., None]), 2.0).astype(np.float32)
    theta += 0.1 * rng.standard_normal(theta.shape, dtype=np.float32)
    np.abs(theta, out=theta)
    out = rng.random((H, W)) < 0.1
    theta[out] = 2.0 * rng.random((int(out.sum()), L), dtype=np.float32)
    return theta


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled in-process every
    ~2 ms (the timed region of the small workloads lasts tens of ms, shorter than nvidia-smi's start-up); nvidia-smi
    -lms is the fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.nvml, self.running = [], None, index, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else None
            phys = ids[self.index] if ids and self.index < len(ids) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = (pynvml, h)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nvml
        bits = [getattr(pynvml, n, 0) for n in ("nvmlClocksEventReasonHwSlowdown", "nvmlClocksEventReasonHwThermalSlowdown",
                                                "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksEventReasonSwPowerCap")]
        bits = [b or f for b, f in zip(bits, (0x8, 0x40, 0x20, 0x4))]
        while self.running:
            try:
                mhz = int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append([str(mhz), str(self.max_mhz)] + ["Active" if mask & b else "Not Active" for b in bits])
            except Exception:
                pass
            time.sleep(0.002)

    def mark(self):
        """start of the timed region: samples taken so far (warm-up) are dropped"""
        self.rows = []

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.running = False
            self.thread.join(timeout=2)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = sorted({self.NAMES[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---- CPU baselines (the only places that execute oracle/ code; never inside a timed GPU region) -------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_port_baseline(name, budget_s=12.0):
    """The CPU restatement (oracle/, checker code) timed on ONE host core on a bounded sample of the same workload:
    rows are cropped so that ~budget_s of CPU time covers >= 2 iterations.  Reported baseline, not the target."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from checkers import oracle_lp
    H, W, L, pairwise, lo, hi, _ = WORKLOADS[name]
    if pairwise == "dense":
        W = min(W, 256)    # the oracle holds tables in f64: 32 KiB per edge at L = 64

    def build(rows):
        lp = oracle_lp("f64").grid_create(rows, W, L, pairwise)
        lp.grid_fill_synthetic(SEED, lo, hi)
        if name == "C2":
            lp.grid_set_unaries(stereo_unaries(H, W, L)[:rows])
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
        return lp

    probe = build(min(H, 64))
    t0 = time.perf_counter()
    probe.ComputePass(0); probe.LowerBound()
    per_row = (time.perf_counter() - t0) / min(H, 64)
    rows = int(max(8, min(H, budget_s / 3.0 / max(per_row, 1e-9))))
    lp = build(rows)
    lp.ComputePass(0); lp.LowerBound()
    iters = 2
    t0 = time.perf_counter()
    for it in range(iters):
        lp.ComputePass(it + 1); lp.LowerBound()
    dt = time.perf_counter() - t0
    return {"value": 4.0 * edges_of(rows, W) * iters / dt, "unit": UNIT, "cores": 1, "kind": "port", "host_threads_available": host_threads(),
            "sample": f"{rows}x{W} crop (top rows) of the workload, {iters} iterations after 1 warm-up, f64, oracle/lpmp_oracle.cpp (scalar port: 1 thread)"}


def reference_grid_problem(name, crop_rows, crop_cols):
    """host arrays of the reference arm: the top-left crop of the workload (same generators as the GPU arm's host buffers)"""
    import numpy as np
    H, W, L, pairwise, lo, hi, _ = WORKLOADS[name]
    rows, cols = min(H, crop_rows), min(W, crop_cols)
    if name == "C2":
        theta = stereo_unaries(H, W, L)[:rows, :cols].astype(np.float64)
    else:
        theta = np.random.default_rng(SEED).random((rows, cols, L))
    rng = np.random.default_rng(SEED + 1)
    lam_h, lam_v = rng.uniform(lo, hi, (rows, cols)), rng.uniform(lo, hi, (rows, cols))
    return theta.reshape(rows * cols, L), lam_h, lam_v, rows, cols


def reference_arm(args, name):
    """The reference's OWN scheduler/containers (oracle/_ref: /root/reference/include compiled unmodified + restated
    simplex factors as dense tables) on the host.  The reference is single-threaded (its LP_MP_PARALLEL mode does not compile,
    BASELINE.md 1), so 'all the host threads it can use' is 1.  C2 runs the FULL instance (same config as the GPU arm: 1 M
    unaries, 2.1 M dense 16x16 pairwise factors, ~6 GB); the larger grids run a bounded top-left crop."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from checkers import RefLP, grid_edges, ref_library, trunclin_table, potts_table
    H, W, L, pairwise, lo, hi, _ = WORKLOADS[name]
    full = name in ("C1", "C2") and not args.crop
    crop = args.crop or (max(H, W) if full else (96 if pairwise == "dense" else 256))
    theta, lam_h, lam_v, rows, cols = reference_grid_problem(name, crop, crop)
    ei, ej = grid_edges(rows, cols)
    lib = ref_library()
    kind = "reference" if lib is not None else "port"
    t_build = time.perf_counter()
    if pairwise == "trunclin":
        shared = trunclin_table(L, lo, hi)
    if lib is not None:
        lp = RefLP(lib, ["-v", "0"])
        for t in theta:
            lp.AddUnaryFactor(t)
        rng = np.random.default_rng(SEED + 2)
        for e in range(len(ei)):
            i, j = int(ei[e]), int(ej[e])
            if pairwise == "trunclin":
                tab = shared
            elif pairwise == "potts":
                r, c = divmod(i, cols)
                tab = potts_table(L, lam_h[r, c] if j == i + 1 else lam_v[r, c])
            else:
                tab = rng.random((L, L))
            lp.lib.lpmp_ref_add_pairwise(lp.h, i, j, tab.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double)))
        lp.E = len(ei)
    else:  # oracle/_ref did not travel: fall back to the port (the oracle always exists)
        from checkers import oracle_lp
        lp = oracle_lp("f64")
        lp.add_unaries(theta)
        if pairwise == "trunclin":
            lp.add_pairwise_trunclin(ei, ej, lo, hi)
        elif pairwise == "potts":
            lam = np.array([lam_h[divmod(int(i), cols)] if int(j) == int(i) + 1 else lam_v[divmod(int(i), cols)] for i, j in zip(ei, ej)])
            lp.add_pairwise_potts(ei, ej, lam)
        else:
            lp.add_pairwise_dense(ei, ej, np.random.default_rng(SEED + 2).random((len(ei), L, L)))
        lp.finalize()
    t_build = time.perf_counter() - t_build
    lp.Begin(); lp.set_reparametrization("anisotropic")
    for it in range(args.warmup):
        lp.ComputePass(it); lp.LowerBound()
    t0 = time.perf_counter()
    lb = None
    for it in range(args.steps):
        lp.ComputePass(args.warmup + it); lb = lp.LowerBound()
    dt = time.perf_counter() - t0
    value = 4.0 * len(ei) * args.steps / dt
    sample = (f"{'the full' if full else 'top-left crop of the'} workload, {rows}x{cols} nodes, {len(ei)} dense {L}x{L} pairwise factors per step, f64, "
              f"1 thread (the reference is single-threaded; {host_threads()} host threads available)")
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: {WORKLOADS[name][6]}", "sample": sample, "same_instance_as_gpu_arm": bool(full),
                   "build_s": round(t_build, 2), "lower_bound_after_last_step": lb},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def print_and_quiet(line):
    """print the JSON line, then point fd 1 at /dev/null: the reference's static destructors print memory-pool statistics
    to stdout at exit and the process must still exit NORMALLY (exit hooks run)"""
    print(line, flush=True)
    sys.stdout.flush()
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)


# ---- general graphs (BASELINE config 4) ------------------------------------------------------------------------------
def random_regular_graph(n, d, seed):
    """Configuration model with local repair of self loops / multi-edges (numpy; n*d even)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    stubs = np.repeat(np.arange(n, dtype=np.int64), d)
    rng.shuffle(stubs)
    a, b = stubs[0::2].copy(), stubs[1::2].copy()
    for _ in range(200):
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        key = lo * n + hi
        order = np.argsort(key, kind="stable")
        dup = np.zeros(len(key), dtype=bool)
        dup[order[1:]] = key[order[1:]] == key[order[:-1]]
        bad = np.flatnonzero(dup | (a == b))
        if len(bad) == 0:
            return lo.astype(np.uint64), hi.astype(np.uint64)
        # re-pair the stubs of the bad edges with those of as many random good edges
        pick = np.unique(np.concatenate([bad, rng.integers(0, len(a), size=len(bad) + 1)]))
        pool = np.concatenate([a[pick], b[pick]])
        rng.shuffle(pool)
        a[pick], b[pick] = pool[: len(pick)], pool[len(pick):]
    raise RuntimeError("could not repair the configuration model")


def graph_cpu_sample(name, steps, warmup):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from checkers import oracle_lp
    n, deg, L, _ = GRAPH_WORKLOADS[name]
    m = 200_000   # bounded sample: a random 3-regular graph of the same family on m nodes
    si, sj = random_regular_graph(m, deg, SEED)
    rng = np.random.default_rng(SEED)
    lp = oracle_lp("f64")
    lp.add_unaries(rng.random((m, L))); lp.add_pairwise_dense(si, sj, rng.random((len(si), L, L)))
    lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
    for it in range(warmup):
        lp.ComputePass(it); lp.LowerBound()
    t0 = time.perf_counter()
    for it in range(steps):
        lp.ComputePass(warmup + it); lp.LowerBound()
    dt = time.perf_counter() - t0
    sample = f"random {deg}-regular graph on {m} nodes ({len(si)} edges) per step, f64, 1 thread, natural order, oracle/lpmp_oracle.cpp"
    return 4.0 * len(si) * steps / dt, dt, sample


def run_graph(args, name, rank, world, local_rank, dist, with_e2e=True):
    """C4: replicas only (a random 3-regular graph is an expander: no balanced partition with a small cut):
    every rank solves the same problem, value = sum of the ranks' throughputs."""
    import numpy as np
    import torch
    import lp_mp_b200
    n, deg, L, desc = GRAPH_WORKLOADS[name]
    ei, ej = random_regular_graph(n, deg, SEED)
    E = len(ei)
    rng = np.random.default_rng(SEED)
    theta = rng.random((n, L))
    tables = rng.random((E, L, L))
    dtype = args.dtype
    s = 4 if dtype == "f32" else 8
    tune = parse_tune(args.tune) if name == args.workload else {}

    def build():
        lp = lp_mp_b200.LP(dtype=dtype, device=local_rank)
        if tune:
            lp.set_tuning(**tune)
        lp.add_unaries(theta)
        lp.add_pairwise_dense(ei, ej, tables)
        colour, nc = lp.greedy_colouring(ei, ej)
        lp.set_node_ranks(colour.astype(np.uint64) * n + np.arange(n, dtype=np.uint64))
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
        return lp, nc

    lp, nc = build()
    stream = torch.cuda.ExternalStream(lp.stream(), device=local_rank)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()           # polling starts with the warm-up (the first NVML queries of a process are slow) ...
    for it in range(args.warmup):
        lp.ComputePass(it); lp.LowerBound()
    lp.synchronize()
    launches0 = lp.stats()["kernel_launches"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    pass_ev = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.mark()                # ... and only the samples of the timed region are kept
    ev[0].record(stream)
    lb = None
    for it in range(args.steps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream); lp.ComputeForwardPass(); e[1].record(stream); lp.ComputeBackwardPass(); e[2].record(stream)
        pass_ev.append(e)
        lb = lp.LowerBound()
    ev[1].record(stream)
    lp.synchronize(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = max_over_ranks(ev[0].elapsed_time(ev[1]), dist, world)
    clock_info = clocks.stop() if rank == 0 else None
    stats = lp.stats()
    launches = stats["kernel_launches"] - launches0
    value = world * 4.0 * E * args.steps / (ms * 1e-3)
    pass_ms = [e[0].elapsed_time(e[1]) for e in pass_ev] + [e[1].elapsed_time(e[2]) for e in pass_ev]
    mean_pass_s = 1e-3 * sum(pass_ms) / len(pass_ms)
    bytes_pass = s * (2.0 * n * L + E * (L * L + 5.0 * L))
    peak, peak_src = measured_hbm_peak()
    achieved = bytes_pass / mean_pass_s / 1e9
    levels = int(stats["num_levels"])
    per_pass_launches = stats.get("pass_launches", levels)
    roofline = {"bound": "hbm", "kernel": stats.get("pass_kernel", f"graph_node_kernel ({levels} launches = one pass: one per colour class; node-record layout)"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0,
                "traffic": imported_traffic(f"{name}_{dtype}"), "traffic_source": "imported from profiles/traffic.json (ncu --set full capture), not measured in this run",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_pass / max(1, per_pass_launches),
                "launches_per_pass": per_pass_launches, "mean_pass_ms": 1e3 * mean_pass_s}
    del lp
    e2e = None
    if with_e2e:
        def solve():
            q, _ = build()
            bound = 0.0
            for it in range(args.steps):
                q.ComputePass(it)
                bound = q.LowerBound()
            q.ComputeForwardPassAndPrimal(args.steps)
            cost = q.EvaluatePrimal()
            nlab = q.labeling().nbytes
            return bound, cost, nlab

        solve()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = None
        for _ in range(2):     # two timed solves, the faster one is reported (a solve is short; one-off host hiccups are not the metric)
            t0 = time.perf_counter()
            bound, cost, nlab = solve()
            torch.cuda.synchronize()
            d = max_over_ranks(time.perf_counter() - t0, dist, world)
            if dt is None or d < dt:
                dt = d
        h2d = s * (theta.size + tables.size) + 4 * 2 * E + 20 * n
        e2e = {"value": world * 4.0 * E * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d * world / args.steps,
               "d2h_bytes_per_step": (8 * args.steps + 8 + nlab) * world / args.steps,
               "definition": f"one Solve from host arrays through the C ABI (faster of two): add_unaries/add_pairwise_dense + colouring + finalize (upload) + {args.steps} "
                             "iterations (LB read back every iteration) + rounding pass + labeling download; wall clock, max over ranks",
               "lower_bound": bound, "primal": cost}
    return {
        "metric": METRIC.replace("grid MRF", "sparse-graph MRF"), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"{name}: {desc}" + ("; replicas only" if world > 1 else ""), "nodes": n, "edges": E, "colours": int(nc),
                   "step": "one iteration = forward pass + backward pass + lower bound", "schedule": "graph colouring (level sets of the colour-oriented DAG)",
                   "reparametrization": "anisotropic (SRMP)", "l2_policy": f"working set {int(bytes_pass / 2) >> 20} MiB exceeds the 126 MB L2; no flush needed",
                   "lower_bound_after_last_step": lb, "tuning": tune},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clock_info}


# ---- grids -------------------------------------------------------------------------------------------------------------
def parse_tune(text):
    return {k: int(v) for k, v in (kv.split("=") for kv in text.split(",") if kv)}


def max_over_ranks(x, dist, world):
    if world == 1:
        return x
    import torch
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_grid(args, name, rank, world, local_rank, dist, dtype=None, with_e2e=True, with_ttb=True, with_check=True):
    import numpy as np
    import torch
    from lp_mp_b200.stripes import BOTTOM, TOP, StripedLP, StripePlan
    H, W, L, pairwise, lo, hi, desc = WORKLOADS[name]
    dtype = dtype or args.dtype
    s = 4 if dtype == "f32" else 8
    tune = parse_tune(args.tune) if name == args.workload else {}
    H_total = H if (args.scaling == "strong" or world == 1) else H * world
    plan = StripePlan(H_total, world)
    row0, rows_here = plan.rows(rank)
    theta_np = stereo_unaries(H, W, L) if name == "C2" else None      # C2: stereo-like data term (host generated, uploaded before timing)

    def make(coupling=None):
        return StripedLP.make(H_total, W, L, pairwise, rank, world, dist=dist, dtype=dtype, device=local_rank,
                              coupling=coupling or args.coupling, tuning=tune)

    def fill(q):
        q.lp.grid_fill_synthetic(SEED, lo, hi)
        if theta_np is not None:
            q.lp.grid_set_unaries(theta_np[(row0 % H):(row0 % H) + rows_here] if H_total != H else theta_np[row0:row0 + rows_here])
        q.finalize(); q.Begin(); q.set_reparametrization("anisotropic")
        return q

    # ---- device-resident arm: inputs resident in HBM, timed with CUDA events on the library's stream ----
    slp = fill(make())
    lp = slp.lp
    stream = torch.cuda.ExternalStream(lp.stream(), device=local_rank)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()           # polling starts with the warm-up (the first NVML queries of a process are slow) ...
    for it in range(args.warmup):
        slp.ComputePass(it); slp.LowerBound()
    slp.synchronize()
    launches0 = lp.stats()["kernel_launches"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    pass_ev = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.mark()                # ... and only the samples of the timed region are kept
    ev[0].record(stream)
    lb = None
    for it in range(args.steps):
        # per-launch events for the roofline of the dominant kernel (one wavefront pass = one launch)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        slp._deliver(BOTTOM)
        e[0].record(stream); lp.ComputeForwardPass(); e[1].record(stream)
        slp._deliver(TOP)
        e[2].record(stream); lp.ComputeBackwardPass(); e[3].record(stream)
        pass_ev.append(e)
        lb = slp.LowerBound()
    ev[1].record(stream)
    slp.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = max_over_ranks(ev[0].elapsed_time(ev[1]), dist, world)
    clock_info = clocks.stop() if rank == 0 else None
    stats = lp.stats()
    launches = stats["kernel_launches"] - launches0
    E = edges_of(H_total, W)
    value = 4.0 * E * args.steps / (ms * 1e-3)

    pass_ms = [e[0].elapsed_time(e[1]) for e in pass_ev] + [e[2].elapsed_time(e[3]) for e in pass_ev]
    mean_pass_s = 1e-3 * sum(pass_ms) / len(pass_ms)
    bytes_pass = algorithmic_bytes_per_pass(rows_here, W, L, pairwise, s)
    peak, peak_src = measured_hbm_peak()
    achieved = bytes_pass / mean_pass_s / 1e9
    kernel = stats.get("pass_kernel") or {"dense": "dense_row_pass_kernel"}.get(pairwise, "wavefront pass kernel")
    roofline = {"bound": "hbm", "kernel": f"{kernel} (one launch = one pass over this rank's stripe)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0,
                "traffic": imported_traffic(f"{name}_{dtype}"), "traffic_source": "imported from profiles/traffic.json (ncu --set full capture), not measured in this run",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_pass, "mean_launch_ms": 1e3 * mean_pass_s,
                "critical_path_steps_per_launch": rows_here + W - 1,
                "note": "exact reference order: a pass is a dependency chain of H+W-1 node updates, so small grids are bound by the "
                        "per-update latency, not by HBM (DESIGN.md 'Rooflines'; SURVEY.md 7 hard part 1)"}
    slp.close(); del slp, lp   # free the device-resident arm's potentials before the next arm allocates its own

    # ---- N > 1: the striped bound against the unstriped LP of the same problem after the same number of iterations (rank 0's
    # GPU; LP::LowerBound is ONE number, include/LP_MP.h:1507-1518).  Exact for the pipelined coupling; the per-pass halo is a
    # block-Jacobi relaxation across stripes -- its gap is REPORTED, and asserted to stay a small fraction.
    bound_check = None
    if world > 1 and with_check and pairwise != "dense":
        want = None
        if rank == 0:
            q = fill(StripedLP.make(H_total, W, L, pairwise, 0, 1, dist=None, dtype=dtype, device=local_rank, tuning=tune))
            for it in range(args.warmup + args.steps):
                q.ComputePass(it)
            want = q.LowerBound()
            q.close(); del q
        box = [want]
        dist.broadcast_object_list(box, src=0)
        want = box[0]
        gap = (want - lb) / abs(want)
        tol = 1e-6 if args.coupling == "pipelined" else 5e-3
        bound_check = {"unstriped_bound_same_iterations": want, "striped_bound": lb, "rel_gap": gap, "tolerance": tol, "ok": bool(abs(gap) <= tol)}
        assert abs(gap) <= tol, f"striped bound {lb} vs unstriped {want}: relative gap {gap} exceeds {tol}"

    # ---- time to dual bound (BASELINE.json metric iii): fresh LP, maxIter iterations with the bound read back every
    # iteration; LB* = the bound after maxIter (the device path's bound equals the oracle's, tests/test_gpu_parity.py); reported: the
    # first iteration (and wall time up to it) with LB >= LB* - 1e-4 |LB*|, and whether the run had converged by then.
    ttb = None
    if world == 1 and with_ttb and name in TTB_ITERS:
        iters = TTB_ITERS[name]
        q = fill(make())
        q.synchronize()
        t0 = time.perf_counter()
        hist = []
        for it in range(iters):
            q.ComputePass(it)
            hist.append((q.LowerBound(), time.perf_counter() - t0))
        lb_star = hist[-1][0]
        k = next(i for i, (b, _) in enumerate(hist) if b >= lb_star - 1e-4 * abs(lb_star))
        tail = (lb_star - hist[int(0.9 * iters)][0]) / abs(lb_star)
        ttb = {"iterations": k + 1, "ms": 1e3 * hist[k][1], "bound": lb_star, "max_iter": iters, "rel_tol": 1e-4,
               "first_bound": hist[0][0], "ms_all_iterations": 1e3 * hist[-1][1],
               "rel_improvement_over_last_10pct_of_iterations": tail, "converged": bool(tail < 1e-5)}
        q.close(); del q

    # ---- end-to-end arm: the call a user makes, from pinned HOST buffers ----
    # one solve = upload this rank's potentials (H2D) + K iterations each with a lower-bound read-back (D2H) + rounding pass +
    # primal cost + labeling download; reported per iteration (upload amortised over the K iterations of the solve)
    e2e = None
    if with_e2e:
        P = {"dense": L * L, "potts": 1, "trunclin": 2}[pairwise]
        e2e_rows = rows_here
        if world == 1 and 2 * rows_here * W * P * 4 > 4e9:
            # the potentials of the full workload do not fit a host upload (C5: 137 GB of tables): the end-to-end arm solves
            # the top rows that make a 4 GB upload and says so
            e2e_rows = max(8, int(4e9 / (2 * W * P * 4)))
        g = torch.Generator().manual_seed(SEED + rank)
        if theta_np is not None:
            theta_h = torch.from_numpy(np.ascontiguousarray(theta_np[row0 % H:(row0 % H) + e2e_rows])).pin_memory()
        else:
            theta_h = torch.rand((e2e_rows, W, L), dtype=torch.float32, generator=g).pin_memory()

        def edge_params(shape):
            if pairwise == "trunclin":
                return torch.tensor([lo, hi], dtype=torch.float32).repeat(*shape, 1).contiguous().pin_memory()
            return (lo + (hi - lo) * torch.rand((*shape, P), dtype=torch.float32, generator=g)).pin_memory()

        hz_h, vt_h = edge_params((e2e_rows, W - 1)), edge_params((e2e_rows - 1, W))
        top_h = edge_params((W,)) if row0 > 0 else None
        bot_h = edge_params((W,)) if row0 + rows_here < H_total else None
        lab_h = torch.empty(e2e_rows * W, dtype=torch.int32).pin_memory()      # the labeling comes back into pinned memory too
        phase_t = []

        def solve():
            tp = [time.perf_counter()]
            q = make() if e2e_rows == rows_here else StripedLP.make(e2e_rows, W, L, pairwise, 0, 1, dist=None, dtype=dtype, device=local_rank, tuning=tune)
            q.lp._ck(q.lp.api.grid_set_unaries(q.lp.h, theta_h.data_ptr(), 0))
            q.lp._ck(q.lp.api.grid_set_pairwise(q.lp.h, hz_h.data_ptr(), vt_h.data_ptr(), 0))
            if world > 1:
                q.lp._ck(q.lp.api.grid_set_boundary_pairwise(q.lp.h, top_h.data_ptr() if top_h is not None else None,
                                                             bot_h.data_ptr() if bot_h is not None else None, 0))
            tp.append(time.perf_counter())
            q.finalize(); q.Begin(); q.set_reparametrization("anisotropic")
            tp.append(time.perf_counter())
            bound = 0.0
            for it in range(args.steps):
                q.ComputePass(it)
                bound = q.LowerBound()
                tp.append(time.perf_counter())
            q.ComputeForwardPassAndPrimal(args.steps)
            cost = q.EvaluatePrimal()
            nlab = q.lp.labeling(out=lab_h.numpy().view(np.uint32)).nbytes      # every rank downloads its own stripe's labels
            q.close()
            tp.append(time.perf_counter())
            phase_t[:] = [round(1e3 * (y - x), 3) for x, y in zip(tp, tp[1:])]
            return bound, cost, nlab

        solve()  # warm-up (allocator, first touch, IPC mappings)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = None
        for _ in range(2):     # two timed solves, the faster one is reported (a solve is short; one-off host hiccups are not the metric)
            t0 = time.perf_counter()
            bound, cost, nlab = solve()
            torch.cuda.synchronize()
            d = max_over_ranks(time.perf_counter() - t0, dist, world)
            if dt is None or d < dt:
                dt, best_phases = d, list(phase_t)
        h2d = 4 * (theta_h.numel() + hz_h.numel() + vt_h.numel() + (top_h.numel() if top_h is not None else 0) + (bot_h.numel() if bot_h is not None else 0))
        d2h = 8 * args.steps + 8 + nlab
        if world > 1:
            tot = torch.tensor([float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
            dist.all_reduce(tot)
            h2d, d2h = float(tot[0].item()), float(tot[1].item())
        E_e2e = edges_of(e2e_rows, W) if e2e_rows != rows_here else E
        e2e = {"value": 4.0 * E_e2e * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "definition": f"one Solve from pinned host buffers (faster of two): upload + {args.steps} iterations (LB read back every iteration)"
                             " + rounding pass + primal cost + labeling download; wall clock, max over ranks; bytes are per iteration, summed over ranks"
                             + (f"; on the top {e2e_rows} rows of the workload (4 GB host upload)" if e2e_rows != rows_here else ""),
               "lower_bound": bound, "primal": cost,
               "phases_ms": {"upload": best_phases[0], "finalize": best_phases[1], "iterations": best_phases[2:2 + args.steps], "rounding_download_close": best_phases[-1]}}

    coupling_text = {"halo": "per-pass halo through NVLink peer memory (push kernel into the neighbour's self-validating in-buffer; block-Jacobi across stripes)",
                     "halo_nccl": "per-pass halo over NCCL send/recv (block-Jacobi across stripes)",
                     "pipelined": "exact pipelined (NVLink peer stores from the pass kernel; bit-identical to one GPU)"}[args.coupling]
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling if world > 1 else "strong", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"{name}: {desc}" + (f"; {world} row stripes of {rows_here} rows (grid {H_total}x{W})" if world > 1 else ""),
                   "step": "one iteration = forward pass + backward pass + lower bound", "schedule": "wavefront (exact reference order inside a stripe)",
                   "reparametrization": "anisotropic (SRMP)", "stripe_coupling": None if world == 1 else coupling_text,
                   "unaries": "stereo-like data term (SURVEY.md 8d)" if name == "C2" else "U[0,1) (device generator)",
                   "l2_policy": f"working set per GPU (node records{' + tables' if pairwise == 'dense' else ''}, {int(bytes_pass / 2) >> 20} MiB) exceeds the 126 MB L2; no flush needed",
                   "tuning": tune, "lower_bound_after_last_step": lb, "passes_per_s": 2.0 * args.steps / (ms * 1e-3)},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clock_info, "time_to_bound": ttb, "bound_check": bound_check,
    }


def default_solve(name, max_iter, local_rank):
    """A full Solve() with the reference's default control law (MpRoundingSolver + StandardVisitor defaults: rounding iterations with
    damped-uniform weights every 5th iteration, include/visitors/standard_visitor.hxx:32-44; include/solver.hxx:230-257,381-400)
    through the host mirror lp_mp_b200/solver.py: wall clock from Begin() to the end of Solve(), potentials already in HBM."""
    import lp_mp_b200
    from lp_mp_b200 import MpRoundingSolver, StandardVisitor
    H, W, L, pairwise, lo, hi, desc = WORKLOADS[name]
    lp = lp_mp_b200.LP(dtype="f32", device=local_rank).grid_create(H, W, L, pairwise)
    lp.grid_fill_synthetic(SEED, lo, hi)
    lp.finalize()
    s = MpRoundingSolver(lp, StandardVisitor(maxIter=max_iter))
    t0 = time.perf_counter()
    s.Solve()
    lp.synchronize()
    dt = time.perf_counter() - t0
    return {"workload": f"{name}: {desc}", "control_law": "StandardVisitor defaults (primalComputationInterval 5, damped_uniform rounding)", "maxIter": max_iter,
            "seconds": dt, "ms_per_iteration": 1e3 * dt / max_iter, "value": 4.0 * edges_of(H, W) * max_iter / dt, "unit": UNIT,
            "lower_bound": s.lower_bound(), "primal": s.primal_cost(),
            "note": "rounding iterations run the uniform-mode kernels (4 message updates per edge and pass, slower schedule): DESIGN.md 8"}


def default_solve_on_stripes(args, name, rank, world, local_rank, dist, max_iter):
    """The same full Solve() on the row stripes of `name` over all ranks (--default-solve, N > 1): the rounding iterations run the
    uniform-mode passes on stripes (every rank first delivers the vectors its neighbours' last virtual rows receive from).  Wall
    clock from Begin() to the end of Solve(), max over ranks, potentials already in HBM."""
    import torch
    from lp_mp_b200 import MpRoundingSolver, StandardVisitor
    from lp_mp_b200.stripes import StripedLP
    H, W, L, pairwise, lo, hi, desc = WORKLOADS[name]
    out = {}
    for coupling in (args.coupling,) if args.coupling == "pipelined" else (args.coupling, "pipelined"):
        slp = StripedLP.make(H, W, L, pairwise, rank, world, dist=dist, dtype=args.dtype, device=local_rank, coupling=coupling)
        slp.lp.grid_fill_synthetic(SEED, lo, hi)
        slp.finalize()
        s = MpRoundingSolver(slp, StandardVisitor(maxIter=max_iter))
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        s.Solve()
        slp.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0, dist, world)
        out[coupling] = {"seconds": dt, "ms_per_iteration": 1e3 * dt / max_iter, "value": 4.0 * edges_of(H, W) * max_iter / dt, "unit": UNIT,
                         "lower_bound": s.lower_bound(), "primal": s.primal_cost()}
        slp.close()
        dist.barrier()
    if rank == 0:
        # the unstriped run of the same law on rank 0's GPU: what the stripes' bounds are checked against
        one = default_solve(name, max_iter, local_rank)
        out["one_gpu"] = {k: one[k] for k in ("seconds", "ms_per_iteration", "value", "unit", "lower_bound", "primal")}
    dist.barrier()
    out.update({"workload": f"{name}: {desc}; {world} row stripes", "maxIter": max_iter,
                "control_law": "StandardVisitor defaults (primalComputationInterval 5, damped_uniform rounding)"})
    return out


def compact(block):
    """an extra workload's block inside the headline line"""
    keep = ("value", "unit", "ms_per_step", "dtype", "roofline", "e2e", "gpu_launches", "time_to_bound", "cpu_baseline")
    out = {k: block[k] for k in keep if block.get(k) is not None}
    out["workload"] = block["config"]["workload"]
    out["lower_bound_after_last_step"] = block["config"].get("lower_bound_after_last_step")
    if out.get("e2e"):
        out["e2e"] = {k: v for k, v in out["e2e"].items() if k != "phases_ms"}
    return out


def main():
    import faulthandler
    faulthandler.enable()   # a native crash prints the Python stack of every thread
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS) + sorted(GRAPH_WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the workload's grid cut into N row stripes; weak = N stripes of the workload's size")
    ap.add_argument("--coupling", default="halo", choices=["halo", "halo_nccl", "pipelined"],
                    help="N > 1: per-pass halo (peer-memory push / NCCL send-recv; block-Jacobi across stripes) or exact pipelined "
                         "stripes (NVLink peer stores from the pass kernel; bit-identical to one GPU)")
    ap.add_argument("--tune", default="", help="comma separated key=value launch-shape knobs, e.g. wavefront_tile_cols=4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the blocks of the other workloads")
    ap.add_argument("--default-solve", type=int, default=0, metavar="MAXITER",
                    help="N > 1, grid workloads: add a full default-law Solve() of MAXITER iterations on the stripes (and on one GPU) to the line")
    ap.add_argument("--crop", type=int, default=0, help="reference arm: side of the top-left crop (default: full C1/C2, 256 else)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            if args.workload in GRAPH_WORKLOADS:
                value, dt, sample = graph_cpu_sample(args.workload, args.steps, args.warmup)
                line = {"impl": "reference", "metric": METRIC.replace("grid MRF", "sparse-graph MRF"), "value": value, "unit": UNIT, "n_gpus": args.gpus,
                        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
                        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": args.workload, "sample": sample},
                        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
                        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            else:
                line = reference_arm(args, args.workload)
            print_and_quiet(json.dumps(line))
        return

    import torch
    import lp_mp_b200  # noqa: F401  (fails loudly when the CUDA library is missing: there is no fallback)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload in GRAPH_WORKLOADS:
        out = run_graph(args, args.workload, rank, world, local_rank, dist)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            v, _, sample = graph_cpu_sample(args.workload, 3, 1)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}
    else:
        out = run_grid(args, args.workload, rank, world, local_rank, dist)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_port_baseline(args.workload)
    if rank == 0 and "cpu_baseline" not in out:
        out["cpu_baseline"] = None

    # ---- N = 1: the other single-GPU configurations in the same line (their own roofline / end-to-end blocks) ----
    if world == 1 and not args.no_extras and args.workload == "C3" and args.dtype == "f32":
        extras = {}
        saved = args.tune
        args.tune = ""
        for name in ("C2", "C4", "C5h"):
            try:
                blk = run_graph(args, name, 0, 1, local_rank, None) if name in GRAPH_WORKLOADS else run_grid(args, name, 0, 1, local_rank, None)
                if name == "C2" and not args.no_cpu_baseline:
                    blk["cpu_baseline"] = cpu_port_baseline(name, budget_s=6.0)
                extras[name] = compact(blk)
            except Exception as exc:   # an extra block must never cost the headline line
                extras[name] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()
        try:
            extras["C3_f64"] = compact(run_grid(args, "C3", 0, 1, local_rank, None, dtype="f64", with_e2e=False, with_ttb=False))
        except Exception as exc:
            extras["C3_f64"] = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            extras["C2_default_solve"] = default_solve("C2", 100, local_rank)
        except Exception as exc:
            extras["C2_default_solve"] = {"error": f"{type(exc).__name__}: {exc}"}
        args.tune = saved
        out["workloads"] = extras

    if world > 1 and args.default_solve > 0 and args.workload in WORKLOADS and WORKLOADS[args.workload][3] != "dense":
        blk = default_solve_on_stripes(args, args.workload, rank, world, local_rank, dist, args.default_solve)
        if rank == 0:
            out["default_solve_on_stripes"] = blk

    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
