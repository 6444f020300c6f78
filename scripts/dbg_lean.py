"""Where does the lean variant differ from the base wavefront kernel?  python scripts/dbg_lean.py H W L kind [fwd|bwd] [tune]"""
import sys, numpy as np
sys.path.insert(0, '.')
import lp_mp_b200
H, W, L, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
which = sys.argv[5] if len(sys.argv) > 5 else "fwd"
tune = {k: int(v) for k, v in (kv.split("=") for kv in (sys.argv[6] if len(sys.argv) > 6 else "").split(",") if kv)}
lo, hi = {"potts": (0.1, 1.0), "trunclin": (0.25, 1.0)}[kind]
def make(**t):
    lp = lp_mp_b200.LP(dtype="f32", device=0).grid_create(H, W, L, kind)
    lp.set_tuning(**t)
    lp.grid_fill_synthetic(0x5EED, lo, hi)
    lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
    return lp
a, b = make(), make(**{"wavefront_lean": 1, **tune})
for lp in (a, b):
    lp.ComputeForwardPass()
    if which == "bwd":
        lp.ComputeBackwardPass()
    lp.synchronize()
ua, ub = a.unaries().reshape(H, W, L), b.unaries().reshape(H, W, L)
bad = np.argwhere((ua != ub).any(axis=2))
print(f"{H}x{W} L={L} {kind} {which} {tune}: theta differs at {len(bad)} nodes")
if len(bad):
    rows, cnt = np.unique(bad[:, 0], return_counts=True)
    print(" rows:", dict(zip(rows.tolist()[:40], cnt.tolist()[:40])))
    r0 = rows[0]
    print(" first bad row", r0, "cols", bad[bad[:, 0] == r0][:40, 1].tolist())
(xa, xb), (ya, yb) = a.messages(), b.messages()
print(" a differs:", int((xa != ya).any(axis=-1).sum()) if xa.ndim > 1 else int((xa != ya).sum()), " b differs:", int((xb != yb).any(axis=-1).sum()) if xb.ndim > 1 else int((xb != yb).sum()))
if len(bad):
    r0, c0 = bad[0]
    print(" base ", ua[r0, c0, :6], "\n lean ", ub[r0, c0, :6])
    if c0 + 1 < W: print(" right neighbour equal:", bool((ua[r0, c0 + 1] == ub[r0, c0 + 1]).all()))
# per-edge detail (edge order: per node, right edge then down edge)
ea = []
for r in range(H):
    for c in range(W):
        if c + 1 < W: ea.append(("R", r, c))
        if r + 1 < H: ea.append(("D", r, c))
da = np.argwhere((xa != ya).any(axis=1))[:, 0]
db = np.argwhere((xb != yb).any(axis=1))[:, 0]
from collections import Counter
print(" a-side diffs by (kind,row):", sorted(Counter((ea[e][0], ea[e][1]) for e in da).items())[:30])
print(" b-side diffs by (kind,row of first endpoint):", sorted(Counter((ea[e][0], ea[e][1]) for e in db).items())[:30])
print(" first a diffs:", [ea[e] for e in da[:12]], " first b diffs:", [ea[e] for e in db[:12]])
