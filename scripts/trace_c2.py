import sys, numpy as np
sys.path.insert(0, '.')
import lp_mp_b200
H, W, L = 1024, 1024, 16
lp = lp_mp_b200.LP(dtype="f32", device=0).grid_create(H, W, L, "trunclin")
lp.set_tuning(wavefront_trace=1)
lp.grid_fill_synthetic(0x5EED, 0.25, 1.0)
lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
for it in range(3):
    lp.ComputeForwardPass(); lp.ComputeBackwardPass()
lp.ComputeForwardPass(); lp.synchronize()
tr = lp.trace(128).astype(np.int64)
t0 = tr[:, 0].min()
tr -= t0
print("band claimed tile0 firstT done  (us)")
for k in list(range(0, 8)) + list(range(8, 128, 16)) + [127]:
    print(k, *(f"{x/1e3:9.1f}" for x in tr[k]))
d = np.diff(tr[:, 2]) / 1e3
print("lag between consecutive bands (firstT stamp) us: mean %.2f min %.2f max %.2f" % (d.mean(), d.min(), d.max()))
dur = (tr[:, 3] - tr[:, 2]) / 1e3
print("band duration after first T steps us: mean %.1f min %.1f max %.1f" % (dur.mean(), dur.min(), dur.max()))
print("total us", tr[:, 3].max() / 1e3)
