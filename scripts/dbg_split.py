"""A/B of the split / rows variants against the oracle on one grid: which configurations differ, and how long a pass takes."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import lp_mp_b200
from checkers import oracle_lp

def build(make, dtype, H, W, L, kind, seed, tuning=None):
    rng = np.random.default_rng(seed)
    theta = rng.random((H, W, L))
    if kind == "potts":
        hz, vt = rng.uniform(-0.3, 1.0, size=(H, W - 1, 1)), rng.uniform(-0.3, 1.0, size=(H - 1, W, 1))
    else:
        hz = np.stack([rng.uniform(0.05, 0.5, size=(H, W - 1)), rng.uniform(0.2, 1.5, size=(H, W - 1))], axis=-1)
        vt = np.stack([rng.uniform(0.05, 0.5, size=(H - 1, W)), rng.uniform(0.2, 1.5, size=(H - 1, W))], axis=-1)
    lp = make(dtype).grid_create(H, W, L, kind)
    lp.grid_set_unaries(theta); lp.grid_set_pairwise(hz, vt)
    if tuning: lp.set_tuning(**tuning)
    lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
    return lp

cases = [(k, L, d, m, v) for v in ("wavefront_split",) for k, L in (("potts", 16), ("trunclin", 16), ("potts", 32)) for d in ("f32", "f64") for m in (0, 1, 3)]
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (203, 157)
for kind, L, dtype, mc, var in cases:
    if L == 32 and dtype == "f64": continue
    ref = build(oracle_lp, dtype, H, W, L, kind, 23)
    tun = {var: 1}
    if mc: tun["wavefront_max_ctas"] = mc
    lp = build(lambda d: lp_mp_b200.LP(dtype=d, device=0), dtype, H, W, L, kind, 23, tun)
    res = []
    for fwd in (True, False):
        t0 = time.time()
        try:
            (lp.ComputeForwardPass if fwd else lp.ComputeBackwardPass)(); lp.synchronize()
            err = ""
        except Exception as e:
            err = str(e)[:60]
        dt = time.time() - t0
        (ref.ComputeForwardPass if fwd else ref.ComputeBackwardPass)()
        a, b = lp.unaries(), ref.unaries()
        bad = np.argwhere((a != b).any(axis=1)).ravel()
        res.append(f"{'fwd' if fwd else 'bwd'} {dt*1e3:.1f}ms ndiff={len(bad)} first={(divmod(int(bad[0]), W) if len(bad) else None)} {err}")
    print(kind, L, dtype, "max_ctas", mc, "|", " | ".join(res), flush=True)
