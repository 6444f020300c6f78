"""SURVEY.md 8f-2 / 8f-3 on the CPU: mixed label counts and the dual archive of the standalone oracle against the reference core
(oracle/_ref: the reference's own containers and archive classes, include/factor_archive.hxx:34-48,
include/serialization.hxx:226-425).  The GPU versions of the same checks live in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from checkers import RefLP, grid_edges, oracle_lp, ref_library

MODES = ["anisotropic", "uniform", "damped_uniform"]


def mixed_instance(seed, n=14, m=26, counts=(2, 3, 4, 5)):
    """variables with 2..5 labels as in test/graphical_model.cpp:74-96 (negPotts23, posPotts24, posPotts34)"""
    rng = np.random.default_rng(seed)
    Ls = rng.choice(counts, size=n)
    edges = set()
    while len(edges) < m:
        i, j = rng.integers(0, n, size=2)
        if i != j:
            edges.add((int(min(i, j)), int(max(i, j))))
    edges = sorted(edges)
    rng.shuffle(edges)
    theta = [rng.random(L) for L in Ls]
    tables = [rng.random((Ls[i], Ls[j])) - 0.3 for i, j in edges]
    return theta, edges, tables


def build(lp, theta, edges, tables, interleaved=False):
    if interleaved:   # factors in mixed insertion order: the archive follows the order of the add_* calls
        half = len(theta) // 2
        lp.add_unaries_var(theta[:half])
        first = [(k, e) for k, e in enumerate(edges) if e[1] < half]
        rest = [(k, e) for k, e in enumerate(edges) if e[1] >= half]
        if first:
            lp.add_pairwise_dense([e[0] for _, e in first], [e[1] for _, e in first], [tables[k] for k, _ in first])
        lp.add_unaries_var(theta[half:])
        lp.add_pairwise_dense([e[0] for _, e in rest], [e[1] for _, e in rest], [tables[k] for k, _ in rest])
        return [k for k, _ in first] + [k for k, _ in rest]
    for t in theta:
        lp.AddUnaryFactor(t)
    for (i, j), t in zip(edges, tables):
        lp.AddPairwiseFactor(i, j, t)
    return list(range(len(edges)))


def same_state(x, y):
    xa, xb = x.messages()
    ya, yb = y.messages()
    return np.array_equal(x.unaries(), y.unaries()) and np.array_equal(xa, ya) and np.array_equal(xb, yb)


@pytest.mark.parametrize("mode", MODES)
def test_mixed_label_counts_bit_parity_with_the_reference_core(ref_lib, mode):
    theta, edges, tables = mixed_instance(11)
    ref, o = RefLP(ref_lib, ["-v", "0"]), oracle_lp("f64")
    build(ref, theta, edges, tables); build(o, theta, edges, tables)
    ref.Begin(); o.Begin()
    ref.set_reparametrization(mode); o.set_reparametrization(mode)
    assert o.LowerBound() == pytest.approx(ref.LowerBound(), rel=1e-13)
    for it in range(5):
        ref.ComputePass(it); o.ComputePass(it)
        assert same_state(ref, o)
        assert o.LowerBound() == pytest.approx(ref.LowerBound(), rel=1e-13, abs=1e-13)
    for it in (5, 6):
        ref.ComputeForwardPassAndPrimal(it); o.ComputeForwardPassAndPrimal(it)
        assert np.array_equal(ref.labeling(), o.labeling())
        ref.ComputeBackwardPassAndPrimal(it); o.ComputeBackwardPassAndPrimal(it)
        assert np.array_equal(ref.labeling(), o.labeling())
        assert o.EvaluatePrimal() == pytest.approx(ref.EvaluatePrimal(), rel=1e-13)
        assert same_state(ref, o)


def test_reference_test_tables_with_mixed_label_counts(ref_lib):
    """the 2x3 / 2x4 / 3x4 tables of test/graphical_model.cpp:74-96 on a chain of a 2-, 3-, 4- and 2-label variable"""
    negPotts23 = np.array([[1.0, 0.0, 2.0], [0.0, 1.0, 2.0]])
    posPotts34 = np.array([[0.0, 1.0, 2.0, 2.0], [1.0, 0.0, 2.0, 2.0], [2.0, 2.0, 2.0, 2.0]])
    posPotts24 = np.array([[0.0, 1.0, 2.0, 2.0], [1.0, 0.0, 2.0, 2.0]])
    theta = [np.zeros(2), np.zeros(3), np.zeros(4), np.zeros(2)]
    edges, tables = [(0, 1), (1, 2), (2, 3)], [negPotts23, posPotts34, np.ascontiguousarray(posPotts24.T)]
    ref, o = RefLP(ref_lib, ["-v", "0"]), oracle_lp("f64")
    build(ref, theta, edges, tables); build(o, theta, edges, tables)
    for lp in (ref, o):
        lp.Begin(); lp.set_reparametrization("anisotropic")
        for it in range(10):
            lp.ComputePass(it)
        lp.ComputeForwardPassAndPrimal(10)
    assert same_state(ref, o)
    assert o.LowerBound() == pytest.approx(0.0, abs=1e-12) and ref.LowerBound() == pytest.approx(0.0, abs=1e-12)
    assert o.EvaluatePrimal() == pytest.approx(ref.EvaluatePrimal(), abs=1e-12)
    assert np.array_equal(ref.labeling(), o.labeling())


@pytest.mark.parametrize("shape", ["uniform", "mixed"])
def test_dual_archive_has_the_reference_layout(ref_lib, shape):
    """save_dual == the bytes the reference's own save_archive writes over all factors in insertion order"""
    if shape == "mixed":
        theta, edges, tables = mixed_instance(5)
    else:
        rng = np.random.default_rng(2)
        ei, ej = grid_edges(4, 5)
        edges = list(zip(ei.tolist(), ej.tolist()))
        theta, tables = list(rng.random((20, 3))), list(rng.random((len(edges), 3, 3)))
    ref, o = RefLP(ref_lib, ["-v", "0"]), oracle_lp("f64")
    build(ref, theta, edges, tables); build(o, theta, edges, tables)
    ref.Begin(); o.Begin()
    ref.set_reparametrization("damped_uniform"); o.set_reparametrization("damped_uniform")
    for it in range(3):
        ref.ComputePass(it); o.ComputePass(it)
    blob = o.save_dual()
    assert len(blob) == 8 * (sum(len(t) for t in theta) + sum(t.shape[0] + t.shape[1] for t in tables))
    assert blob == ref.save_dual()
    # ... and the reference reads it back: load into a fresh reference LP, continue, same state as the uninterrupted run
    ref2 = RefLP(ref_lib, ["-v", "0"])
    build(ref2, theta, edges, tables)
    ref2.Begin(); ref2.set_reparametrization("damped_uniform")
    ref2.load_dual(blob)
    for it in range(3, 6):
        ref.ComputePass(it); ref2.ComputePass(it); o.ComputePass(it)
    assert same_state(ref, ref2) and same_state(ref, o)


def test_dual_archive_follows_the_insertion_order_and_resumes():
    theta, edges, tables = mixed_instance(9)
    a, b = oracle_lp("f64"), oracle_lp("f64")
    build(a, theta, edges, tables)
    order = build(b, theta, edges, tables, interleaved=True)
    for lp in (a, b):
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
    a.ComputePass(0)
    blob = a.save_dual()
    # b holds the same factors in another insertion order: its archive is a permutation of a's, not the same bytes
    b.ComputePass(0)
    assert sorted(np.frombuffer(b.save_dual())) == sorted(np.frombuffer(blob)) and b.save_dual() != blob
    # checkpoint / resume: new handle, same construction calls, load, continue == uninterrupted run
    c = oracle_lp("f64")
    build(c, theta, edges, tables)
    c.finalize(); c.Begin(); c.set_reparametrization("anisotropic")
    c.load_dual(blob)
    for it in range(1, 4):
        a.ComputePass(it); c.ComputePass(it)
        assert same_state(a, c) and a.LowerBound() == c.LowerBound()
    with pytest.raises(Exception):
        c.load_dual(blob[:-8])
    assert order   # (the interleaved build really did split the edges)
