import sys, numpy as np
sys.path.insert(0, '.')
import lp_mp_b200
wl = sys.argv[1]; tune = {k: int(v) for k, v in (kv.split("=") for kv in (sys.argv[2] if len(sys.argv) > 2 else "").split(",") if kv)}
H, W, L, kind, lo, hi, T = {"C2": (1024, 1024, 16, "trunclin", 0.25, 1.0, 8), "C3": (4096, 4096, 32, "potts", 0.1, 1.0, 4)}[wl]
lp = lp_mp_b200.LP(dtype="f32", device=0).grid_create(H, W, L, kind)
lp.set_tuning(wavefront_trace=2, **tune)
lp.grid_fill_synthetic(0x5EED, lo, hi)
lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
for it in range(2):
    lp.ComputeForwardPass(); lp.ComputeBackwardPass()
lp.ComputeForwardPass(); lp.synchronize()
nb = H // T
tr = lp.trace(nb).astype(np.int64)
ts = lp.trace_steps(nb).astype(np.int64)
k = nb // 2
a, b = ts[0], ts[1]
print(wl, tune, "producer band", k)
print(" producer end of step T-1 (column 0 of its last row final): 0 ns")
print(" producer's out-mover hands column 0 on:      %+d ns" % (tr[k, 0] - a[T - 1]))
print(" consumer has taken column 0 from its halo:   %+d ns" % (tr[k + 1, 2] - a[T - 1]))
print(" consumer end of step 0:                      %+d ns" % (b[0] - a[T - 1]))
print(" (consumer's own out-mover first hand-out %+d ns after its step T-1)" % (tr[k + 1, 0] - b[T - 1]))
