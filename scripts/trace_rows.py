"""Row timeline of the split wavefront kernel (wavefront_trace=2): when each grid row hands on its column 0 and when it finishes.
python scripts/trace_rows.py H W L kind"""
import sys, numpy as np
sys.path.insert(0, '.')
import lp_mp_b200
H, W, L, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
lo, hi = {"potts": (0.1, 1.0), "trunclin": (0.25, 1.0)}[kind]
lp = lp_mp_b200.LP(dtype="f32", device=0).grid_create(H, W, L, kind)
lp.set_tuning(wavefront_trace=2)
lp.grid_fill_synthetic(0x5EED, lo, hi)
lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
for it in range(2):
    lp.ComputeForwardPass(); lp.ComputeBackwardPass()
lp.ComputeForwardPass(); lp.synchronize()
nb = (H + 7) // 8
ts = lp.trace_steps(nb).astype(np.int64)
n = min(H, 4096)
st = ts.reshape(4, 4096)[:, :n]          # per row: column 0 handed on, first whole tile reached, last whole tile done, row finished
t0 = st[0].min()
names = ["first step / tile done", "row started", "(unused)", "row finished"]
print(f"{H}x{W} L={L} {kind}: total {(st[3].max() - t0) / 1e3:.1f} us")
inner = np.array([i for i in range(n - 1) if (i + 1) % 8 != 0])
cta = np.array([i for i in range(n - 1) if (i + 1) % 8 == 0 and (i + 1) % 64 != 0])
cl = np.array([i for i in range(n - 1) if (i + 1) % 64 == 0])
for k in (1, 0, 3):
    d = np.diff(st[k])
    print(f" {names[k]:26s} row->row delay ns: inside a CTA {d[inner].mean():.0f}" + (f", CTA->CTA (DSMEM) {d[cta].mean():.0f}" if len(cta) else "") + (f", cluster->cluster (L2) {d[cl].mean():.0f}" if len(cl) else ""))
seg = st[3] - st[1]
print(f" row duration: row 0 {seg[0]} ns, row 1 {seg[1]}, row 7 {seg[7]}, mean {seg.mean():.0f}, last row {seg[-1]}  ({seg[0] / W:.1f} ns / step in row 0)")
m = n // 2
for k in (0, 3):
    print(f" rows {m}..{m+16} {names[k]} diffs:", np.diff(st[k])[m:m + 16].tolist())
