import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_parity import grid_pair
H, W, L, kind, dtype = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], "f32"
tuning = {k: int(v) for k, v in (kv.split("=") for kv in sys.argv[5].split(","))} if len(sys.argv) > 5 else {}
lps = grid_pair(dtype, H, W, L, kind, seed=H + W + L, **tuning)
for lp in lps:
    lp.finalize(); lp.Begin(); lp.set_reparametrization("uniform"); lp.ComputeForwardPass()
t0, t1 = lps[0].unaries().reshape(H, W, L), lps[1].unaries().reshape(H, W, L)
a0, b0 = lps[0].messages(); a1, b1 = lps[1].messages()
eh, ev, e = {}, {}, 0          # edges in insertion order: per node, horizontal then vertical
for r in range(H):
    for c in range(W):
        if c + 1 < W: eh[(r, c)] = e; e += 1
        if r + 1 < H: ev[(r, c)] = e; e += 1
def field(a, b, r, c):
    out = {}
    if c + 1 < W: out["AR"] = a[eh[(r, c)]]
    if c > 0: out["BL"] = b[eh[(r, c - 1)]]
    if r + 1 < H: out["AD"] = a[ev[(r, c)]]
    if r > 0: out["BU"] = b[ev[(r - 1, c)]]
    return out
cnt = 0
for r in range(0, H):
    for c in range(0, W):
        f0, f1 = field(a0, b0, r, c), field(a1, b1, r, c)
        bad = [k for k in f0 if not np.array_equal(f0[k], f1[k])]
        tb = not np.array_equal(t0[r, c], t1[r, c])
        if bad or tb:
            k = bad[0] if bad else None
            msg = ""
            if k:
                i = int(np.nonzero(f0[k] != f1[k])[0][0])
                msg = f" {k}[{i}] oracle {f0[k][i]!r} gpu {f1[k][i]!r} (labels differing: {int((f0[k] != f1[k]).sum())})"
            cnt += 1
            if cnt <= 10: print(f"node ({r},{c}): theta {'BAD' if tb else 'ok'} fields bad {bad}{msg}")

print("total bad nodes", cnt)
