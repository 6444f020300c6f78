"""Anisotropic2 weights (include/LP_MP.h:1084-1154) and residual sends (include/factors_messages.hxx:2270-2279,
2957-3007) -- the standalone oracle against the reference core on the CPU (the GPU versions are in tests/test_gpu_graphs.py)."""
import numpy as np
import pytest

from checkers import RefLP, grid_edges, oracle_lp
from test_dual_archive import build, mixed_instance, same_state

ALL_MODES = ["anisotropic", "anisotropic2", "uniform", "damped_uniform"]


def lockstep(ref, others, mode, iters=4):
    for lp in [ref] + others:
        lp.set_reparametrization(mode)
    for it in range(iters):
        ref.ComputePass(it)
        for lp in others:
            lp.ComputePass(it)
            assert same_state(ref, lp), (mode, it)
            assert np.isclose(lp.LowerBound(), ref.LowerBound(), rtol=1e-12, atol=1e-12)
    # rounding passes always send in shared mode (UpdateFactorPrimal, include/LP_MP.h:1592-1602)
    ref.ComputeForwardPassAndPrimal(iters)
    for lp in others:
        lp.ComputeForwardPassAndPrimal(iters)
        assert np.array_equal(ref.labeling(), lp.labeling()) and same_state(ref, lp)
    ref.ComputeBackwardPassAndPrimal(iters)
    for lp in others:
        lp.ComputeBackwardPassAndPrimal(iters)
        assert np.array_equal(ref.labeling(), lp.labeling()) and same_state(ref, lp)


@pytest.mark.parametrize("send", ["shared", "residual"])
@pytest.mark.parametrize("mode", ALL_MODES)
@pytest.mark.parametrize("shape", ["mixed", "grid"])
def test_send_variants_bit_parity_with_the_reference_core(ref_lib, send, mode, shape):
    if shape == "mixed":
        theta, edges, tables = mixed_instance(3, n=22, m=50)
    else:
        rng = np.random.default_rng(8)
        ei, ej = grid_edges(5, 6)
        edges = list(zip(ei.tolist(), ej.tolist()))
        theta, tables = list(rng.random((30, 4))), list(rng.random((len(edges), 4, 4)) - 0.3)
    ref, o = RefLP(ref_lib, ["-v", "0", "--reparametrizationType", send]), oracle_lp("f64")
    build(ref, theta, edges, tables); build(o, theta, edges, tables)
    ref.Begin(); o.Begin(); o.set_send_mode(send)
    lockstep(ref, [o], mode)


def test_adaptive_sends_are_refused():
    """the reference's adaptive mode needs send_message_to_right_improvement of the (absent) marginalisation message: with the
    restated message the reference core makes no progress at all, so the mode is refused rather than guessed"""
    o = oracle_lp("f64")
    o.add_unaries(np.zeros((2, 2))); o.add_pairwise_dense([0], [1], np.zeros((1, 2, 2)))
    o.finalize(); o.Begin()
    with pytest.raises(Exception):
        o.set_send_mode("adaptive")
