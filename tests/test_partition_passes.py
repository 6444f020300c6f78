"""SURVEY.md 8f-4 on the CPU: LP::put_in_same_partition / compute_partition_pass / compute_overlapping_partition_pass
(include/LP_MP.h:463-465,1716-2070) of the standalone oracle against the reference core.  Partitions hold at most 16 variables: the
reference sorts a partition's factors with a comparator that compares an element with itself (include/LP_MP.h:1776), so only such
partitions keep a well-defined order (std::sort then is a plain insertion sort)."""
import numpy as np
import pytest

from checkers import RefLP, grid_edges, oracle_lp
from test_dual_archive import build, mixed_instance, same_state


def grid_problem(H, W, L, seed):
    rng = np.random.default_rng(seed)
    ei, ej = grid_edges(H, W)
    edges = list(zip(ei.tolist(), ej.tolist()))
    return list(rng.random((H * W, L))), edges, list(rng.random((len(edges), L, L)) - 0.3)


def row_partitions(H, W, rows):
    pi, pj = [], []
    for r in range(H):
        for c in range(W):
            n = r * W + c
            if c + 1 < W:
                pi.append(n); pj.append(n + 1)
            if r % rows != rows - 1 and r + 1 < H:
                pi.append(n); pj.append(n + W)
    return pi, pj


@pytest.mark.parametrize("kind", ["partition", "overlapping_partition"])
@pytest.mark.parametrize("shape", ["grid_rows", "grid_row_pairs", "mixed_chunks", "singletons"])
def test_partition_passes_bit_parity_with_the_reference_core(ref_lib, kind, shape):
    if shape == "mixed_chunks":
        theta, edges, tables = mixed_instance(4, n=30, m=64)
        pi = [k for k in range(29) if k % 5 != 4]
        pj = [k + 1 for k in pi]
    else:
        H, W = 6, 5
        theta, edges, tables = grid_problem(H, W, 3, 21)
        pi, pj = ([], []) if shape == "singletons" else row_partitions(H, W, 1 if shape == "grid_rows" else 2)
    inner = 2
    ref = RefLP(ref_lib, ["-v", "0", "--reparametrizationType", kind, "--innerIteration", str(inner)])
    o = oracle_lp("f64")
    build(ref, theta, edges, tables); build(o, theta, edges, tables)
    if pi:
        ref.put_in_same_partition(pi, pj); o.put_in_same_partition(pi, pj)
    ref.Begin(); o.Begin()
    ref.set_reparametrization("anisotropic"); o.set_reparametrization("anisotropic")
    prev = o.LowerBound()
    for it in range(3):
        ref.ComputePass(it)     # LP::ComputePass in partition mode == the staged pass (include/LP_MP.h:869-887)
        o.compute_partition_pass(inner, kind == "overlapping_partition")
        if kind == "overlapping_partition":
            o.ComputeForwardPass(); o.ComputeBackwardPass()
        assert same_state(ref, o), it
        lb = o.LowerBound()
        assert lb == pytest.approx(ref.LowerBound(), rel=1e-13) and lb >= prev - 1e-9
        prev = lb
