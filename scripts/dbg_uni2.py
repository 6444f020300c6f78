import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_parity import grid_pair
kind, L, dtype = sys.argv[1], int(sys.argv[2]), sys.argv[3]
for (H, W) in ((16, 32), (16, 16)):
  for tuning in ({}, dict(wavefront_cluster=1), dict(wavefront_max_ctas=1), dict(wavefront_cluster=1, wavefront_max_ctas=1)):
    for eo in ("lifo",):
        lps = grid_pair(dtype, H, W, L, kind, seed=H + W + L, edge_order=eo, **tuning)
        for lp in lps:
            lp.finalize(); lp.Begin(); lp.set_reparametrization("uniform")
            lp.ComputeForwardPass()
        t0, t1 = lps[0].unaries().reshape(H, W, L), lps[1].unaries().reshape(H, W, L)
        bad = (t0 != t1).any(axis=2)
        rows = {int(r): int(bad[r].sum()) for r in range(H) if bad[r].any()}
        print(f"{kind} L={L} {H}x{W} {tuning}: bad nodes per row {rows}")
        if bad.any():
            r, c = np.argwhere(bad)[0]
            print("   first bad node", r, c, "labels differing", np.nonzero(t0[r, c] != t1[r, c])[0].tolist(), "oracle", t0[r, c][:4], "gpu", t1[r, c][:4])
