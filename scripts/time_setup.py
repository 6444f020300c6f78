import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import lp_mp_b200
H, W, L = 1024, 1024, 16
theta = torch.rand((H, W, L), dtype=torch.float32).pin_memory()
hz = torch.tensor([0.25, 1.0], dtype=torch.float32).repeat(H, W - 1, 1).contiguous().pin_memory()
vt = torch.tensor([0.25, 1.0], dtype=torch.float32).repeat(H - 1, W, 1).contiguous().pin_memory()
for rep in range(4):
    t = [time.perf_counter()]
    lp = lp_mp_b200.LP(dtype="f32", device=0); t.append(time.perf_counter())
    lp.grid_create(H, W, L, "trunclin"); t.append(time.perf_counter())
    lp._ck(lp.api.grid_set_unaries(lp.h, theta.data_ptr(), 0)); t.append(time.perf_counter())
    lp._ck(lp.api.grid_set_pairwise(lp.h, hz.data_ptr(), vt.data_ptr(), 0)); t.append(time.perf_counter())
    lp.finalize(); t.append(time.perf_counter())
    lp.Begin(); lp.set_reparametrization("anisotropic"); lp.ComputePass(0); lb = lp.LowerBound(); t.append(time.perf_counter())
    del lp; torch.cuda.synchronize(); t.append(time.perf_counter())
    print("create %.2f grid_create %.2f set_unaries %.2f set_pairwise %.2f finalize %.2f first_iter %.2f destroy %.2f ms" % tuple(1e3 * (b - a) for a, b in zip(t, t[1:])))
