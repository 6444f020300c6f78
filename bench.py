#!/usr/bin/env python
"""Benchmark of the LP_MP message-passing hot path (contract: see the task prompt / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2|C3|C1] [--dtype f32|f64]

A "step" is ONE ITERATION of the hot path = forward pass + backward pass (LP::ComputePass, include/LP_MP.h:869-887)
+ the lower-bound reduction (LP::LowerBound, :1507-1518) over one synthetic MRF.  Metric: pairwise message updates/s
(anisotropic: one receive + one send per edge per pass => 4E per iteration).

N = 1: BASELINE.json configs[1] (C2): 1024x1024 truncated-linear grid, 16 labels, wavefront schedule.
N > 1: the same per-GPU work (weak scaling): a (1024*N)x1024 grid cut into N row stripes, one per rank, with a
       halo exchange of the stripe-boundary message vectors after every pass (NCCL send/recv) and an all-reduce of the
       bound.
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
    "C2": (1024, 1024, 16, "trunclin", 0.25, 1.0, "1024x1024 truncated-linear grid, 16 labels (BASELINE config 2)"),
    "C3": (4096, 4096, 32, "potts", 0.1, 1.0, "4096x4096 Potts grid, 32 labels (BASELINE config 3)"),
}
SEED = 0x5EED
METRIC = "pairwise msg updates/s on grid MRF"
UNIT = "msg_updates/s"


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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---- CPU baselines ---------------------------------------------------------------------------------------------------
def cpu_port_baseline(H, W, L, pairwise, lo, hi, budget_s=12.0):
    """The CPU restatement (oracle/, checker code) timed on ONE host core on a bounded sample of the same workload:
    rows are cropped so that ~budget_s of CPU time covers >= 2 iterations.  Reported baseline, not the target."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from checkers import oracle_lp
    rows = H
    probe = oracle_lp("f64").grid_create(min(H, 64), W, L, pairwise)
    probe.grid_fill_synthetic(SEED, lo, hi)
    probe.finalize(); probe.Begin(); probe.set_reparametrization("anisotropic")
    t0 = time.perf_counter()
    probe.ComputePass(0); probe.LowerBound()
    per_row = (time.perf_counter() - t0) / min(H, 64)
    rows = int(max(8, min(H, budget_s / 3.0 / max(per_row, 1e-9))))
    lp = oracle_lp("f64").grid_create(rows, W, L, pairwise)
    lp.grid_fill_synthetic(SEED, lo, hi)
    lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
    lp.ComputePass(0); lp.LowerBound()
    iters = 2
    t0 = time.perf_counter()
    for it in range(iters):
        lp.ComputePass(it + 1); lp.LowerBound()
    dt = time.perf_counter() - t0
    return {"value": 4.0 * edges_of(rows, W) * iters / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{rows}x{W} crop (top rows) of the workload, {iters} iterations after 1 warm-up, f64, oracle/lpmp_oracle.cpp"}


def reference_arm(args, H, W, L, pairwise, lo, hi):
    """The reference's OWN scheduler/containers (oracle/_ref: /root/reference/include compiled unmodified + restated
    simplex factors as dense tables) on the host, single-threaded as the reference is."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from checkers import RefLP, grid_edges, ref_library, trunclin_table, potts_table
    lib = ref_library()
    crop = min(H, W, 192)
    rng = np.random.default_rng(SEED)
    theta = rng.random((crop * crop, L))
    ei, ej = grid_edges(crop, crop)
    if pairwise == "trunclin":
        tab = trunclin_table(L, lo, hi)
        tables = [tab] * len(ei)
    else:
        lam = rng.uniform(lo, hi, size=len(ei))
        tables = [potts_table(L, x) for x in lam]
    kind = "reference"
    if lib is not None:
        lp = RefLP(lib, ["-v", "0"])
        for t in theta:
            lp.AddUnaryFactor(t)
        for e in range(len(ei)):
            lp.AddPairwiseFactor(int(ei[e]), int(ej[e]), tables[e])
    else:  # oracle/_ref did not travel: fall back to the port (the oracle always exists)
        from checkers import oracle_lp
        kind = "port"
        lp = oracle_lp("f64")
        lp.add_unaries(theta)
        lp.add_pairwise_dense(ei, ej, np.stack(tables))
        lp.finalize()
    lp.Begin(); lp.set_reparametrization("anisotropic")
    for it in range(args.warmup):
        lp.ComputePass(it); lp.LowerBound()
    t0 = time.perf_counter()
    for it in range(args.steps):
        lp.ComputePass(args.warmup + it); lp.LowerBound()
    dt = time.perf_counter() - t0
    value = 4.0 * len(ei) * args.steps / dt
    sample = f"{crop}x{crop} crop of the workload per step, f64, 1 thread (the reference is single-threaded)"
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


# ---- GPU arm ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--coupling", default="halo", choices=["halo", "pipelined"],
                    help="N > 1: per-pass halo exchange (NCCL send/recv, block-Jacobi across stripes) or exact pipelined "
                         "stripes (NVLink peer stores from the pass kernel; bit-identical to one GPU)")
    ap.add_argument("--tune", default="", help="comma separated key=value launch-shape knobs, e.g. wavefront_tile_cols=4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    H, W, L, pairwise, lo, hi, desc = WORKLOADS[args.workload]

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args, H, W, L, pairwise, lo, hi)), flush=True)
        os._exit(0)   # skip the reference's static destructors (its MemoryPool prints pool statistics to stdout at exit)

    import numpy as np
    import torch
    import lp_mp_b200
    from lp_mp_b200.stripes import StripedLP, StripePlan

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    s = 4 if args.dtype == "f32" else 8
    tune = {k: int(v) for k, v in (kv.split("=") for kv in args.tune.split(",") if kv)}
    H_total = H * world      # weak scaling: every rank owns a stripe of the single-GPU workload's size

    def make():
        return StripedLP.make(H_total, W, L, pairwise, rank, world, dist=dist, dtype=args.dtype, device=local_rank,
                              coupling=args.coupling, tuning=tune)

    # ---- device-resident arm: inputs generated in HBM, timed with CUDA events on the library's stream ----
    slp = make()
    slp.lp.grid_fill_synthetic(SEED, lo, hi)
    slp.finalize(); slp.Begin(); slp.set_reparametrization("anisotropic")
    lp = slp.lp
    stream = torch.cuda.ExternalStream(lp.stream(), device=local_rank)

    for it in range(args.warmup):
        slp.ComputePass(it); slp.LowerBound()
    slp.synchronize()
    launches0 = lp.stats()["kernel_launches"]
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    pass_ev = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record(stream)
    lb = None
    for it in range(args.steps):
        # per-launch events for the roofline of the dominant kernel (one wavefront pass = one launch)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        if world > 1 and args.coupling == "halo":
            slp._exchange(1)
        e[0].record(stream); lp.ComputeForwardPass(); e[1].record(stream)
        if world > 1 and args.coupling == "halo":
            slp._exchange(0)
        e[2].record(stream); lp.ComputeBackwardPass(); e[3].record(stream)
        pass_ev.append(e)
        lb = slp.LowerBound()
    ev[1].record(stream)
    slp.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev[0].elapsed_time(ev[1])
    clock_info = clocks.stop() if rank == 0 else None
    launches = lp.stats()["kernel_launches"] - launches0
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    E = edges_of(H_total, W)
    value = 4.0 * E * args.steps / (ms * 1e-3)

    pass_ms = [e[0].elapsed_time(e[1]) for e in pass_ev] + [e[2].elapsed_time(e[3]) for e in pass_ev]
    mean_pass_s = 1e-3 * sum(pass_ms) / len(pass_ms)
    rows_here = StripePlan(H_total, world).rows(rank)[1]
    bytes_pass = algorithmic_bytes_per_pass(rows_here, W, L, pairwise, s)
    peak, peak_src = measured_hbm_peak()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{args.workload}_{args.dtype}")
        except Exception:
            traffic = None
    achieved = bytes_pass / mean_pass_s / 1e9
    roofline = {"bound": "hbm", "kernel": "wavefront_pass_kernel (one launch = one pass over this rank's stripe)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_pass, "mean_launch_ms": 1e3 * mean_pass_s,
                "critical_path_steps_per_launch": rows_here + W - 1,
                "note": "exact reference order: a pass is a dependency chain of H+W-1 node updates, so small grids are bound by the "
                        "per-update latency, not by HBM (DESIGN.md 'Rooflines'; SURVEY.md 7 hard part 1)"}

    # ---- end-to-end arm: the call a user makes, from pinned HOST buffers ----
    # one solve = upload this rank's potentials (H2D) + K iterations each with a lower-bound read-back (D2H) [+ rounding
    # pass + labeling download at N = 1]; reported per iteration (upload amortised over the K iterations of the solve)
    P = {"dense": L * L, "potts": 1, "trunclin": 2}[pairwise]
    row0 = StripePlan(H_total, world).rows(rank)[0]
    g = torch.Generator().manual_seed(SEED + rank)
    theta_h = torch.rand((rows_here, W, L), dtype=torch.float32, generator=g).pin_memory()

    def edge_params(shape):
        if pairwise == "trunclin":
            return torch.tensor([lo, hi], dtype=torch.float32).repeat(*shape, 1).contiguous().pin_memory()
        return (lo + (hi - lo) * torch.rand((*shape, P), dtype=torch.float32, generator=g)).pin_memory()

    hz_h, vt_h = edge_params((rows_here, W - 1)), edge_params((rows_here - 1, W))
    top_h = edge_params((W,)) if row0 > 0 else None
    bot_h = edge_params((W,)) if row0 + rows_here < H_total else None

    def solve():
        q = make()
        q.lp._ck(q.lp.api.grid_set_unaries(q.lp.h, theta_h.data_ptr(), 0))
        q.lp._ck(q.lp.api.grid_set_pairwise(q.lp.h, hz_h.data_ptr(), vt_h.data_ptr(), 0))
        if world > 1:
            q.lp._ck(q.lp.api.grid_set_boundary_pairwise(q.lp.h, top_h.data_ptr() if top_h is not None else None,
                                                         bot_h.data_ptr() if bot_h is not None else None, 0))
        q.finalize(); q.Begin(); q.set_reparametrization("anisotropic")
        bound, cost, nlab = 0.0, None, 0
        for it in range(args.steps):
            q.ComputePass(it)
            bound = q.LowerBound()
        if world == 1:
            q.lp.ComputeForwardPassAndPrimal(args.steps)
            cost = q.lp.EvaluatePrimal()
            nlab = q.lp.labeling().nbytes
        q.close()
        return bound, cost, nlab

    solve()  # warm-up (allocator, first touch, IPC mappings)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    bound, cost, nlab = solve()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    h2d = 4 * (theta_h.numel() + hz_h.numel() + vt_h.numel() + (top_h.numel() if top_h is not None else 0) + (bot_h.numel() if bot_h is not None else 0))
    d2h = 8 * args.steps + (8 + nlab if world == 1 else 0)
    e2e = {"value": 4.0 * E * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d * world / args.steps,
           "d2h_bytes_per_step": d2h * world / args.steps,
           "definition": f"one Solve from pinned host buffers: upload + {args.steps} iterations (LB read back every iteration)"
                         + (" + rounding pass + labeling download" if world == 1 else "") + "; wall clock, max over ranks; bytes are per iteration",
           "lower_bound": bound, "primal": cost}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_baseline(H, W, L, pairwise, lo, hi)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}" + (f"; x{world} row stripes of {H} rows (grid {H_total}x{W})" if world > 1 else ""),
                   "step": "one iteration = forward pass + backward pass + lower bound", "schedule": "wavefront (exact reference order)",
                   "reparametrization": "anisotropic (SRMP)",
                   "stripe_coupling": None if world == 1 else ("per-pass halo (NCCL send/recv; block-Jacobi across stripes)" if args.coupling == "halo"
                                                               else "exact pipelined (NVLink peer stores from the pass kernel)"),
                   "l2_policy": f"working set per GPU (node records, {(rows_here * W * L * 5 * s) >> 20} MiB) exceeds the 126 MB L2; no flush needed",
                   "tuning": tune, "lower_bound_after_last_step": lb},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clock_info,
    }
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
