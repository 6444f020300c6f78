"""Times forward/backward passes and the bound of one synthetic grid on cuda:0 (CUDA events on the library's stream)."""
import sys, json
sys.path.insert(0, '.')
import torch
import lp_mp_b200
H, W, L, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
dtype = sys.argv[5] if len(sys.argv) > 5 else "f32"
tune = dict(kv.split("=") for kv in sys.argv[6].split(",")) if len(sys.argv) > 6 and sys.argv[6] else {}
lo, hi = {"potts": (0.1, 1.0), "trunclin": (0.25, 1.0), "dense": (0.0, 1.0)}[kind]
lp = lp_mp_b200.LP(dtype=dtype, device=0).grid_create(H, W, L, kind)
if tune:
    lp.set_tuning(**{k: int(v) for k, v in tune.items()})
lp.grid_fill_synthetic(0x5EED, lo, hi)
lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
stream = torch.cuda.ExternalStream(lp.stream())
s = 4 if dtype == "f32" else 8
P = {"dense": L * L, "potts": 1, "trunclin": 2}[kind]
N, E = H * W, H * (W - 1) + (H - 1) * W
bytes_pass = s * (2.0 * N * L + E * (P + 5.0 * L))
lbs = []
for it in range(2):
    lp.ComputePass(it); lbs.append(lp.LowerBound())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
fw, bw, lbt = [], [], []
for it in range(3):
    ev[0].record(stream); lp.ComputeForwardPass(); ev[1].record(stream); lp.ComputeBackwardPass(); ev[2].record(stream)
    lbs.append(lp.LowerBound()); ev[3].record(stream)
    lp.synchronize()
    fw.append(ev[0].elapsed_time(ev[1])); bw.append(ev[1].elapsed_time(ev[2])); lbt.append(ev[2].elapsed_time(ev[3]))
print(json.dumps({"grid": [H, W, L, kind, dtype], "tune": tune, "fwd_ms": min(fw), "bwd_ms": min(bw), "lb_ms": min(lbt),
                  "GBps_fwd": bytes_pass / min(fw) / 1e6, "GBps_bwd": bytes_pass / min(bw) / 1e6, "lb": lbs,
                  "monotone": all(b >= a - 1e-6 * abs(a) for a, b in zip(lbs, lbs[1:])), "device_GB": lp.stats()["device_bytes"] / 1e9}))
