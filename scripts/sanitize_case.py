"""One small grid through anisotropic and uniform passes + rounding, for compute-sanitizer: python scripts/sanitize_case.py H W L kind [tune]"""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_parity import grid_pair
from common import run_lockstep
H, W, L, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
tuning = {k: int(v) for k, v in (kv.split("=") for kv in sys.argv[5].split(","))} if len(sys.argv) > 5 else {}
for mode in ("anisotropic", "damped_uniform"):
    lps = grid_pair("f32", H, W, L, kind, seed=7, **tuning)
    run_lockstep(lps, mode, iters=2, lb_rtol=1e-6)
print("sanitize case ok", H, W, L, kind, tuning)
