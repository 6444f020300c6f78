"""Pass times of ONE row stripe (one GPU's share of C3 at 8 GPUs: 512 x 4096, Potts, L = 32; or of C2) per reparametrization mode.
The stripe has neighbours on both sides; what they would deliver is pushed from the stripe's own boundary rows (timing only):
python scripts/time_stripe_modes.py [C3|C2] [rows]"""
import sys
sys.path.insert(0, '.')
import torch, lp_mp_b200
wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
H, W, L, kind, lo, hi = {"C2": (1024, 1024, 16, "trunclin", 0.25, 1.0), "C3": (4096, 4096, 32, "potts", 0.1, 1.0)}[wl]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else H // 8
lp = lp_mp_b200.LP(dtype="f32", device=0).grid_create(rows, W, L, kind)
lp.grid_set_stripe(rows, H)
lp.grid_fill_synthetic(0x5EED, lo, hi)
lp.finalize(); lp.Begin()
in_top, in_bot = lp.grid_boundary_ptrs(0)[0], lp.grid_boundary_ptrs(1)[0]
stream = torch.cuda.ExternalStream(lp.stream())
def run(fwd, uni, primal):
    # carried side: what the neighbour's last virtual row would hand over; later side (uniform modes): its first virtual row's vectors
    lp.grid_push_boundary(1 if fwd else 0, in_top if fwd else in_bot, not fwd)
    if uni:
        lp.grid_push_boundary(0 if fwd else 1, in_bot if fwd else in_top, not fwd)
    name = "Compute" + ("Forward" if fwd else "Backward") + "Pass" + ("AndPrimal" if primal else "")
    getattr(lp, name)(*((1,) if primal else ()))
def timed(fn, n=3):
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); lp.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for mode in ("anisotropic", "uniform", "damped_uniform"):
    lp.set_reparametrization(mode)
    uni = mode != "anisotropic"
    run(True, uni, False); run(False, uni, False)
    print(wl, f"stripe {rows}x{W}", mode, "fwd %.3f ms" % timed(lambda: run(True, uni, False)), "bwd %.3f ms" % timed(lambda: run(False, uni, False)),
          "fwd+primal %.3f ms" % timed(lambda: run(True, uni, True)), "bwd+primal %.3f ms" % timed(lambda: run(False, uni, True)), flush=True)
