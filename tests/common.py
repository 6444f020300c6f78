"""Test bodies shared by the CPU suite (oracle) and the GPU suite (CUDA library through the C ABI)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["grid_dense_anisotropic", "grid_dense_uniform", "grid_dense_damped_uniform", "grid_potts", "grid_trunclin",
                "graph_dense"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def build_from_golden(lp, g, native=True, use_grid_api=False):
    """native=True feeds Potts / truncated-linear parameters (the O(L) paths); False feeds the dense tables."""
    kind = str(g["kind"])
    H, W, L = int(g["H"]), int(g["W"]), int(g["L"])
    ei, ej = g["ei"], g["ej"]
    if use_grid_api and H > 0 and native and kind in ("potts", "trunclin"):
        lp.grid_create(H, W, L, kind)
        lp.grid_set_unaries(g["theta"].reshape(H, W, L))
        hz, vt = np.zeros((H, W - 1, 1 if kind == "potts" else 2)), np.zeros((H - 1, W, 1 if kind == "potts" else 2))
        for e, (i, j) in enumerate(zip(ei, ej)):
            r, c = divmod(int(i), W)
            p = [g["lam"][e]] if kind == "potts" else [float(g["slope"]), float(g["trunc"])]
            if int(j) == int(i) + 1:
                hz[r, c] = p
            else:
                vt[r, c] = p
        lp.grid_set_pairwise(hz, vt)
        return lp
    lp.add_unaries(g["theta"])
    if not native or kind == "dense":
        lp.add_pairwise_dense(ei, ej, g["tables"])
    elif kind == "potts":
        lp.add_pairwise_potts(ei, ej, g["lam"])
    else:
        lp.add_pairwise_trunclin(ei, ej, float(g["slope"]), float(g["trunc"]))
    return lp


def check_against_golden(lp, g, exact=True, rtol=1e-12):
    """Replays the recorded reference run (6 passes + forward/backward rounding) on `lp`."""
    lp.finalize()
    lp.Begin()
    lp.set_reparametrization(str(g["mode"]))
    assert np.isclose(lp.LowerBound(), float(g["ref_lb0"]), rtol=rtol, atol=1e-12)
    for it, want in enumerate(g["ref_lb"]):
        lp.ComputePass(it)
        assert np.isclose(lp.LowerBound(), want, rtol=rtol, atol=1e-12), (it, lp.LowerBound(), want)
    a, b = lp.messages()
    if exact:
        assert np.array_equal(lp.unaries(), g["ref_theta"])
        assert np.array_equal(a, g["ref_a"]) and np.array_equal(b, g["ref_b"])
    else:
        scale = max(1.0, np.abs(g["ref_theta"]).max())
        assert np.allclose(lp.unaries(), g["ref_theta"], rtol=0, atol=1e-5 * scale)
    n = len(g["ref_lb"])
    lp.ComputeForwardPassAndPrimal(n)
    if exact:
        assert np.array_equal(lp.labeling(), g["ref_labels_fwd"])
        assert np.isclose(lp.EvaluatePrimal(), float(g["ref_primal_fwd"]), rtol=rtol)
    lp.ComputeBackwardPassAndPrimal(n)
    if exact:
        assert np.array_equal(lp.labeling(), g["ref_labels_bwd"])
        assert np.isclose(lp.EvaluatePrimal(), float(g["ref_primal_bwd"]), rtol=rtol)
        assert np.array_equal(lp.unaries(), g["ref_theta_end"])
    else:
        assert lp.EvaluatePrimal() >= lp.LowerBound() - 1e-4


def same_state(x, y):
    """bit-level equality of (theta, a, b) of two LPs"""
    if not np.array_equal(x.unaries(), y.unaries()):
        return False
    xa, xb = x.messages()
    ya, yb = y.messages()
    return np.array_equal(xa, ya) and np.array_equal(xb, yb)


def run_lockstep(lps, mode, iters=4, rounding=True, lb_rtol=1e-12):
    """Drives several LPs through the same calls and asserts bit-identical state after every pass."""
    ref = lps[0]
    for lp in lps:
        if not lp.finalized:
            lp.finalize()
        lp.Begin()
        lp.set_reparametrization(mode)
    for it in range(iters):
        for lp in lps:
            lp.ComputePass(it)
        lb0 = ref.LowerBound()
        for lp in lps[1:]:
            assert same_state(ref, lp), f"state differs after pass {it}"
            assert np.isclose(lp.LowerBound(), lb0, rtol=lb_rtol, atol=1e-9)
    if rounding:
        for fwd in (True, False):
            for lp in lps:
                (lp.ComputeForwardPassAndPrimal if fwd else lp.ComputeBackwardPassAndPrimal)(iters)
            for lp in lps[1:]:
                assert np.array_equal(ref.labeling(), lp.labeling()), "labeling differs"
                assert np.isclose(lp.EvaluatePrimal(), ref.EvaluatePrimal(), rtol=lb_rtol)
                assert same_state(ref, lp)
