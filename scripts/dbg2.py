import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import lp_mp_b200
H, W, L, kind, dtype = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
tune = {k: int(v) for k, v in (kv.split("=") for kv in (sys.argv[6] if len(sys.argv) > 6 else "").split(",") if kv)}
lo, hi = {"potts": (0.1, 1.0), "trunclin": (0.25, 1.0)}[kind]
lp = lp_mp_b200.LP(dtype=dtype, device=0).grid_create(H, W, L, kind)
lp.set_tuning(**tune)
lp.grid_fill_synthetic(0x5EED, lo, hi)
lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
for name, fn in (("fwd", lp.ComputeForwardPass), ("bwd", lp.ComputeBackwardPass), ("fwd2", lp.ComputeForwardPass), ("bwd2", lp.ComputeBackwardPass),
                 ("fwdP", lambda: lp.ComputeForwardPassAndPrimal(3)), ("bwdP", lambda: lp.ComputeBackwardPassAndPrimal(3))):
    try:
        fn(); print(name, "LB", lp.LowerBound(), flush=True)
    except Exception as e:
        print(name, "FAILED", e, flush=True); break
