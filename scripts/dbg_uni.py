"""debug: where does the split kernel's uniform-mode state differ from the oracle (python scripts/dbg_uni.py kind L dtype)"""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_parity import grid_pair
kind, L, dtype = sys.argv[1], int(sys.argv[2]), sys.argv[3]
for (H, W), tuning in (((8, 16), {}), ((1, 24), {}), ((37, 32), {}), ((9, 40), dict(wavefront_cluster=1)), ((70, 48), dict(wavefront_max_ctas=2))):
    for mode in ("uniform",):
        for fwd_only in (True, False):
            lps = grid_pair(dtype, H, W, L, kind, seed=H + W + L, **tuning)
            for lp in lps:
                lp.finalize(); lp.Begin(); lp.set_reparametrization(mode)
                if fwd_only: lp.ComputeForwardPass()
                else: lp.ComputeBackwardPass()
            t0, t1 = lps[0].unaries().reshape(H, W, L), lps[1].unaries().reshape(H, W, L)
            bad = np.argwhere((t0 != t1).any(axis=2))
            a0, b0 = lps[0].messages(); a1, b1 = lps[1].messages()
            print(f"{kind} L={L} {dtype} {H}x{W} {'fwd' if fwd_only else 'bwd'}: theta differs at {len(bad)} nodes; first {bad[:6].tolist()}; a differs {int((a0 != a1).any(axis=-1).sum())}, b differs {int((b0 != b1).any(axis=-1).sum())}"
                  + (f"; max abs diff {np.abs(t0 - t1).max():.3g}" if len(bad) else ""))
