"""Per-step timeline of two neighbouring bands (lean variant): python scripts/trace_steps.py C2|C3 [tune]"""
import sys, numpy as np
sys.path.insert(0, '.')
import lp_mp_b200
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
tune = {k: int(v) for k, v in (kv.split("=") for kv in (sys.argv[2] if len(sys.argv) > 2 else "").split(",") if kv)}
H, W, L, kind, lo, hi, T = {"C2": (1024, 1024, 16, "trunclin", 0.25, 1.0, 8), "C3": (4096, 4096, 32, "potts", 0.1, 1.0, 4)}[wl]
lp = lp_mp_b200.LP(dtype="f32", device=0).grid_create(H, W, L, kind)
lp.set_tuning(wavefront_trace=2, **tune)
lp.grid_fill_synthetic(0x5EED, lo, hi)
lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
for it in range(2):
    lp.ComputeForwardPass(); lp.ComputeBackwardPass()
lp.ComputeForwardPass(); lp.synchronize()
nb = H // T
ts = lp.trace_steps(nb).astype(np.int64)
n = W + T - 1
a, b = ts[0, :n], ts[1, :n]
t0 = a[0]
da, db = np.diff(a), np.diff(b)
print(f"{wl} {tune}: producer band {nb//2}, consumer band {nb//2+1}; steps {n}")
print("producer step ns: mean %.0f median %.0f p90 %.0f max %.0f" % (da.mean(), np.median(da), np.percentile(da, 90), da.max()))
print("consumer step ns: mean %.0f median %.0f p90 %.0f max %.0f" % (db.mean(), np.median(db), np.percentile(db, 90), db.max()))
# consumer's step s (group 0 at column s) needs the producer's step s + T - 1
lag = b[:W] - a[T - 1:T - 1 + W]
print("consumer end-of-step(q) minus producer end-of-step(q+T-1) ns: mean %.0f median %.0f min %.0f max %.0f" % (lag.mean(), np.median(lag), lag.min(), lag.max()))
print("same-index lag b[s]-a[s] ns: mean %.0f" % (b[:n] - a[:n]).mean())
for s0 in (64, 512):
    print(" steps", s0, "..", s0 + 24)
    print("  producer d:", da[s0:s0 + 24].tolist())
    print("  consumer d:", db[s0:s0 + 24].tolist())
    print("  lag(q):    ", lag[s0:s0 + 24].tolist())
print(" lag(q) first 48:", lag[:48].tolist())
jumps = [(int(i), int(lag[i]), int(lag[i + 1])) for i in range(len(lag) - 1) if lag[i + 1] - lag[i] > 96]
print(" lag increases > 96 ns (q, before, after):", jumps[:40])
print(" producer first 24 step ns:", da[:24].tolist())
print(" consumer first 24 step ns:", db[:24].tolist())
