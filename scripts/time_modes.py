"""Pass times per reparametrization mode (and rounding passes) on one synthetic grid: python scripts/time_modes.py C2|C3"""
import sys, time
sys.path.insert(0, '.')
import torch, lp_mp_b200
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
H, W, L, kind, lo, hi = {"C2": (1024, 1024, 16, "trunclin", 0.25, 1.0), "C3": (4096, 4096, 32, "potts", 0.1, 1.0)}[wl]
lp = lp_mp_b200.LP(dtype="f32", device=0).grid_create(H, W, L, kind)
lp.grid_fill_synthetic(0x5EED, lo, hi)
lp.finalize(); lp.Begin()
stream = torch.cuda.ExternalStream(lp.stream())
def timed(fn, n=3):
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); lp.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for mode in ("anisotropic", "uniform", "damped_uniform"):
    lp.set_reparametrization(mode)
    lp.ComputePass(0)
    print(wl, mode, "fwd %.3f ms" % timed(lp.ComputeForwardPass), "bwd %.3f ms" % timed(lp.ComputeBackwardPass),
          "fwd+primal %.3f ms" % timed(lambda: lp.ComputeForwardPassAndPrimal(1)), "bwd+primal %.3f ms" % timed(lambda: lp.ComputeBackwardPassAndPrimal(1)), "LB", lp.LowerBound())
