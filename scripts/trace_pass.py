"""Per-band timeline of one wavefront pass (globaltimer stamps written by the kernel, see lpmp_get_trace):
python scripts/trace_pass.py C2|C3 [key=value,...]"""
import sys, numpy as np
sys.path.insert(0, '.')
import lp_mp_b200
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
tune = {k: int(v) for k, v in (kv.split("=") for kv in (sys.argv[2] if len(sys.argv) > 2 else "").split(",") if kv)}
H, W, L, kind, lo, hi, T = {"C2": (1024, 1024, 16, "trunclin", 0.25, 1.0, 8), "C3": (4096, 4096, 32, "potts", 0.1, 1.0, 4)}[wl]
lp = lp_mp_b200.LP(dtype="f32", device=0).grid_create(H, W, L, kind)
lp.set_tuning(wavefront_trace=1, **tune)
lp.grid_fill_synthetic(0x5EED, lo, hi)
lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
for it in range(3):
    lp.ComputeForwardPass(); lp.ComputeBackwardPass()
lp.ComputeForwardPass(); lp.synchronize()
nb = H // T
tr = lp.trace(nb).astype(np.int64)
tr -= tr[:, 0].min()
print(f"{wl} {tune}: band claimed tile0 firstT done  (us)")
for k in list(range(0, 9)) + list(range(16, nb, max(16, nb // 8))) + [nb - 1]:
    print(k, *(f"{x/1e3:9.1f}" for x in tr[k]))
d = np.diff(tr[:, 2]) / 1e3
print("lag between consecutive bands (firstT stamp) us: mean %.2f min %.2f max %.2f" % (d.mean(), d.min(), d.max()))
for m in (2, 4, 8):
    print(f"  mean lag by band index mod {m}:", [round(float(d[(np.arange(len(d)) + 1) % m == r].mean()), 2) for r in range(m)])
dur = (tr[:, 3] - tr[:, 2]) / 1e3
print("band duration after first T steps us: mean %.1f min %.1f max %.1f" % (dur.mean(), dur.min(), dur.max()))
print("total us", tr[:, 3].max() / 1e3)
