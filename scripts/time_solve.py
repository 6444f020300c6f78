"""A full default Solve() -- MpRoundingSolver + StandardVisitor with the reference's default options (damped-uniform rounding
iterations every 5th iteration, include/visitors/standard_visitor.hxx:32-44) -- on one synthetic grid: python scripts/time_solve.py C2|C3 [maxIter]"""
import sys, time, json
sys.path.insert(0, '.')
import lp_mp_b200
from lp_mp_b200 import MpRoundingSolver, StandardVisitor
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 100
H, W, L, kind, lo, hi = {"C2": (1024, 1024, 16, "trunclin", 0.25, 1.0), "C3": (4096, 4096, 32, "potts", 0.1, 1.0), "C5q": (512, 512, 64, "dense", 0.0, 1.0)}[wl]
out = {}
for rounding in ("damped_uniform", "anisotropic"):
    lp = lp_mp_b200.LP(dtype="f32", device=0).grid_create(H, W, L, kind)
    lp.grid_fill_synthetic(0x5EED, lo, hi)
    lp.finalize()
    s = MpRoundingSolver(lp, StandardVisitor(maxIter=max_iter, roundingReparametrization=rounding))
    t0 = time.perf_counter(); s.Solve(); lp.synchronize(); dt = time.perf_counter() - t0
    out[rounding] = {"seconds": dt, "ms_per_iteration": 1e3 * dt / max_iter, "lower_bound": s.lower_bound(), "primal": s.primal_cost()}
print(json.dumps({"workload": wl, "maxIter": max_iter, "rounding": out}))
