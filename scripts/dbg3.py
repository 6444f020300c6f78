import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_parity import grid_pair
H, W, L, kind, dtype = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
tune = {k: int(v) for k, v in (kv.split("=") for kv in (sys.argv[6] if len(sys.argv) > 6 else "").split(",") if kv)}
_, lp = grid_pair(dtype, H, W, L, kind, seed=3, **tune)
lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
t0 = time.time()
for name, fn in (("fwd", lp.ComputeForwardPass), ("bwd", lp.ComputeBackwardPass), ("fwd2", lp.ComputeForwardPass)):
    try:
        fn(); lb = lp.LowerBound(); print(sys.argv[1:], name, "LB", lb, "t=%.2f" % (time.time() - t0), flush=True)
    except Exception as e:
        print(sys.argv[1:], name, "FAILED", str(e)[:80], "t=%.2f" % (time.time() - t0), flush=True); break
