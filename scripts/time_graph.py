"""Times the phases of a C4-shaped solve (random 3-regular graph, dense tables, colouring schedule) on cuda:0: host construction,
finalize (layout build + upload), forward / backward passes and the bound (CUDA events on the library's stream)."""
import sys, json, time
sys.path.insert(0, '.')
import numpy as np
import torch
import lp_mp_b200
from bench import random_regular_graph, SEED
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dtype = sys.argv[3] if len(sys.argv) > 3 else "f32"
tune = dict(kv.split("=") for kv in sys.argv[4].split(",")) if len(sys.argv) > 4 and sys.argv[4] else {}
ei, ej = random_regular_graph(n, 3, SEED)
rng = np.random.default_rng(SEED)
theta, tables = rng.random((n, L)), rng.random((len(ei), L, L))
t = [time.perf_counter()]
lp = lp_mp_b200.LP(dtype=dtype, device=0)
if tune:
    lp.set_tuning(**{k: int(v) for k, v in tune.items()})
lp.add_unaries(theta); lp.add_pairwise_dense(ei, ej, tables); t.append(time.perf_counter())
colour, nc = lp.greedy_colouring(ei, ej); t.append(time.perf_counter())
lp.set_node_ranks(colour.astype(np.uint64) * n + np.arange(n, dtype=np.uint64))
lp.finalize(); t.append(time.perf_counter())
lp.Begin(); lp.set_reparametrization("anisotropic")
stream = torch.cuda.ExternalStream(lp.stream())
s = 4 if dtype == "f32" else 8
E = len(ei)
bytes_pass = s * (2.0 * n * L + E * (L * L + 5.0 * L))
lbs = []
for it in range(2):
    lp.ComputePass(it); lbs.append(lp.LowerBound())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
fw, bw, lbt = [], [], []
for it in range(5):
    ev[0].record(stream); lp.ComputeForwardPass(); ev[1].record(stream); lp.ComputeBackwardPass(); ev[2].record(stream)
    lbs.append(lp.LowerBound()); ev[3].record(stream)
    lp.synchronize()
    fw.append(ev[0].elapsed_time(ev[1])); bw.append(ev[1].elapsed_time(ev[2])); lbt.append(ev[2].elapsed_time(ev[3]))
print(json.dumps({"graph": [n, E, L, dtype], "colours": int(nc), "tune": tune, "add_s": t[1] - t[0], "colouring_s": t[2] - t[1], "finalize_s": t[3] - t[2],
                  "fwd_ms": min(fw), "bwd_ms": min(bw), "lb_ms": min(lbt), "GBps_fwd": bytes_pass / min(fw) / 1e6, "frac": bytes_pass / min(fw) / 1e6 / 6551.0,
                  "lb": lbs[-3:], "levels": lp.stats()["num_levels"], "device_GB": lp.stats()["device_bytes"] / 1e9}))
