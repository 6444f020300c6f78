"""Standalone oracle (oracle/lpmp_oracle.cpp) against the reference's own scheduler/containers/solver/visitor
(oracle/_ref, built from /root/reference/include unmodified) -- bit-level state parity on random instances, the
StandardVisitor control law, and the order-invariance claim of SURVEY.md App. B.  CPU only.
"""
import numpy as np
import pytest

from checkers import RefLP, grid_edges, oracle_lp, potts_table, ref_library, trunclin_table
from lp_mp_b200 import MpRoundingSolver, StandardVisitor

MODES = ["anisotropic", "uniform", "damped_uniform"]


def random_graph(rng, n, m):
    edges = set()
    while len(edges) < m:
        i, j = rng.integers(0, n, size=2)
        if i != j:
            edges.add((min(i, j), max(i, j)))
    edges = sorted(edges)
    rng.shuffle(edges)
    return np.array([e[0] for e in edges], dtype=np.uint64), np.array([e[1] for e in edges], dtype=np.uint64)


def build_pair(lib, theta, ei, ej, tables, options=("-v", "0"), edge_order="lifo"):
    ref = RefLP(lib, options)
    for t in theta:
        ref.AddUnaryFactor(t)
    for e in range(len(ei)):
        ref.AddPairwiseFactor(int(ei[e]), int(ej[e]), tables[e])
    o = oracle_lp("f64")
    o.add_unaries(theta)
    o.add_pairwise_dense(ei, ej, tables)
    o.set_edge_order(edge_order)
    o.finalize()
    return ref, o


def assert_same_state(ref, o):
    assert np.array_equal(ref.unaries(), o.unaries())
    ra, rb = ref.messages()
    oa, ob = o.messages()
    assert np.array_equal(ra, oa) and np.array_equal(rb, ob)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape", ["grid", "random", "star"])
def test_state_bit_parity(ref_lib, mode, shape):
    rng = np.random.default_rng(hash((mode, shape)) % 2 ** 32)
    L = 4
    if shape == "grid":
        H, W = 5, 6
        n = H * W
        ei, ej = grid_edges(H, W)
    elif shape == "random":
        n = 24
        ei, ej = random_graph(rng, n, 50)
    else:
        n = 9
        ei = np.array([0, 0, 0, 0, 1, 2, 3, 4], dtype=np.uint64)
        ej = np.array([5, 6, 7, 8, 8, 8, 8, 8], dtype=np.uint64)
    theta = rng.random((n, L))
    tables = rng.random((len(ei), L, L)) - 0.3
    ref, o = build_pair(ref_lib, theta, ei, ej, tables)
    ref.Begin(); o.Begin()
    ref.set_reparametrization(mode); o.set_reparametrization(mode)
    assert ref.LowerBound() == pytest.approx(o.LowerBound(), rel=1e-13)
    prev = -np.inf
    for it in range(6):
        ref.ComputePass(it); o.ComputePass(it)
        assert_same_state(ref, o)
        lb = o.LowerBound()
        assert ref.LowerBound() == pytest.approx(lb, rel=1e-13, abs=1e-13)
        assert lb >= prev - 1e-9  # monotone dual ascent (debug asserts of include/factors_messages.hxx:2457-2464)
        prev = lb
    for it in (6, 7):
        ref.ComputeForwardPassAndPrimal(it); o.ComputeForwardPassAndPrimal(it)
        assert np.array_equal(ref.labeling(), o.labeling())
        assert ref.EvaluatePrimal() == pytest.approx(o.EvaluatePrimal(), rel=1e-13)
        ref.ComputeBackwardPassAndPrimal(it); o.ComputeBackwardPassAndPrimal(it)
        assert np.array_equal(ref.labeling(), o.labeling())
        assert ref.EvaluatePrimal() == pytest.approx(o.EvaluatePrimal(), rel=1e-13)
        assert_same_state(ref, o)


def test_fifo_storage_order():
    """both message counts variable => chunked FIFO storage (include/factors_messages.hxx:1772-1930)"""
    lib = ref_library(fifo=True)
    if lib is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    H, W, L = 4, 5, 3
    ei, ej = grid_edges(H, W)
    theta, tables = rng.random((H * W, L)), rng.random((len(ei), L, L))
    ref, o = build_pair(lib, theta, ei, ej, tables, edge_order="fifo")
    ref.Begin(); o.Begin()
    ref.set_reparametrization("damped_uniform"); o.set_reparametrization("damped_uniform")
    for it in range(4):
        ref.ComputePass(it); o.ComputePass(it)
    assert_same_state(ref, o)


@pytest.mark.parametrize("options", [
    dict(maxIter=13),
    dict(maxIter=12, roundingReparametrization="anisotropic"),
    dict(maxIter=9, primalComputationInterval=3, primalComputationStart=2, lowerBoundComputationInterval=2),
    dict(maxIter=40, minDualImprovement=1e-2, minDualImprovementInterval=4),
])
def test_solve_control_law(ref_lib, options):
    """Python Solver/StandardVisitor mirror over the standalone oracle == the reference's own Solve() loop
    (include/solver.hxx:230-257, include/visitors/standard_visitor.hxx:110-190), visit by visit."""
    rng = np.random.default_rng(8)
    H, W, L = 8, 8, 4
    ei, ej = grid_edges(H, W)
    theta, tables = rng.random((H * W, L)), rng.random((len(ei), L, L))
    cli = ["-v", "0"]
    for k, v in options.items():
        cli += ["--" + k, str(v)]
    ref, o = build_pair(ref_lib, theta, ei, ej, tables, options=cli)
    r = ref.solve()
    s = MpRoundingSolver(o, StandardVisitor(**options))
    s.Solve()
    assert len(s.history) == len(r["lb"])
    for k, (lb, ub, c) in enumerate(s.history):
        assert lb == pytest.approx(r["lb"][k], rel=1e-12)
        assert ub == r["ub"][k] or ub == pytest.approx(r["ub"][k], rel=1e-12)
        assert (r["flags"][k] & 1) == int(c.computeLowerBound) and ((r["flags"][k] >> 1) & 1) == int(c.computePrimal)
    assert s.lower_bound() == pytest.approx(r["final_lb"], rel=1e-12)
    assert s.primal_cost() == pytest.approx(r["final_ub"], rel=1e-12)


def test_solver_keeps_the_best_labeling():
    """Solver::RegisterPrimal keeps the solution of the best primal cost seen (include/solver.hxx:321-340); the mirror keeps it
    inside the LP (lpmp_keep_labeling: a device-side copy in the product) and fetches it once at the end of Solve()"""
    rng = np.random.default_rng(3)
    H, W, L = 7, 9, 5
    ei, ej = grid_edges(H, W)
    theta, tables = rng.random((H * W, L)), rng.random((len(ei), L, L))
    seen = []

    class Recording(MpRoundingSolver):
        def RegisterPrimal(self):
            before = self.bestPrimalCost_
            super().RegisterPrimal()
            if self.bestPrimalCost_ < before:
                seen.append((self.bestPrimalCost_, self.lp_.labeling().copy()))

    o = oracle_lp("f64")
    o.add_unaries(theta); o.add_pairwise_dense(ei, ej, tables)
    s = Recording(o, StandardVisitor(maxIter=23, primalComputationInterval=2))
    assert s.Solve()
    assert len(seen) >= 2, "the instance should improve its primal more than once"
    best_cost, best_labels = seen[-1]
    assert s.primal_cost() == best_cost
    assert np.array_equal(s.solution_, best_labels)
    # the kept labeling really is the best one, not the last one rounded
    lab = s.solution_
    cost = theta[np.arange(H * W), lab].sum() + sum(tables[e][lab[i], lab[j]] for e, (i, j) in enumerate(zip(ei, ej)))
    assert cost == pytest.approx(best_cost, rel=1e-12)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("kind", ["potts", "potts_neg", "trunclin"])
def test_fast_paths_equal_dense(dtype, kind):
    """O(L) Potts / windowed truncated-linear min-marginals are bit-identical to the materialised table"""
    rng = np.random.default_rng(4)
    H, W, L = 5, 5, 7
    ei, ej = grid_edges(H, W)
    theta = rng.random((H * W, L))
    f = np.float32 if dtype == "f32" else np.float64
    a = oracle_lp(dtype); b = oracle_lp(dtype)
    a.add_unaries(theta); b.add_unaries(theta)
    if kind.startswith("potts"):
        lam = rng.uniform(-0.5 if kind == "potts_neg" else 0.1, 1.0, size=len(ei))
        a.add_pairwise_potts(ei, ej, lam)
        b.add_pairwise_dense(ei, ej, np.stack([potts_table(L, f(x)) for x in lam]))
    else:
        slope, trunc = rng.uniform(0.1, 0.5, size=len(ei)), rng.uniform(0.3, 1.5, size=len(ei))
        a.add_pairwise_trunclin(ei, ej, slope, trunc)
        b.add_pairwise_dense(ei, ej, np.stack([trunclin_table(L, s, t, f) for s, t in zip(slope, trunc)]))
    for lp in (a, b):
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
    for it in range(5):
        a.ComputePass(it); b.ComputePass(it)
        assert np.array_equal(a.unaries(), b.unaries())
        assert a.LowerBound() == pytest.approx(b.LowerBound(), rel=1e-6 if dtype == "f32" else 1e-13)
    a.ComputeForwardPassAndPrimal(5); b.ComputeForwardPassAndPrimal(5)
    assert np.array_equal(a.labeling(), b.labeling())


def test_order_invariance():
    """every linear extension of the relation DAG gives bit-identical state.  Ranks that keep all
    edge orientations (anti-diagonal order, column-major order) must reproduce row-major results exactly."""
    rng = np.random.default_rng(12)
    H, W, L = 6, 5, 4
    ei, ej = grid_edges(H, W)
    theta, tables = rng.random((H * W, L)), rng.random((len(ei), L, L))
    r, c = np.divmod(np.arange(H * W), W)
    ranks = {"row_major": np.arange(H * W), "antidiagonal": np.argsort(np.argsort((r + c) * H * W + r, kind="stable")),
             "column_major": c * H + r}
    states = []
    for name, rank in ranks.items():
        lp = oracle_lp("f64")
        lp.add_unaries(theta); lp.add_pairwise_dense(ei, ej, tables)
        lp.set_node_ranks(rank.astype(np.uint64))
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
        for it in range(5):
            lp.ComputePass(it)
        lp.ComputeForwardPassAndPrimal(5)
        states.append((lp.unaries(), *lp.messages(), lp.labeling()))
    for s in states[1:]:
        for x, y in zip(states[0], s):
            assert np.array_equal(x, y)


def test_grid_constructor_matches_edge_list():
    rng = np.random.default_rng(21)
    H, W, L = 4, 6, 5
    ei, ej = grid_edges(H, W)
    theta = rng.random((H, W, L))
    hz, vt = rng.uniform(0.1, 1, size=(H, W - 1)), rng.uniform(0.1, 1, size=(H - 1, W))
    lam = np.zeros(len(ei))
    for e, (i, j) in enumerate(zip(ei, ej)):
        r, c = divmod(int(i), W)
        lam[e] = hz[r, c] if int(j) == int(i) + 1 else vt[r, c]
    a = oracle_lp("f64").grid_create(H, W, L, "potts")
    a.grid_set_unaries(theta); a.grid_set_pairwise(hz, vt); a.finalize()
    b = oracle_lp("f64")
    b.add_unaries(theta.reshape(-1, L)); b.add_pairwise_potts(ei, ej, lam); b.finalize()
    for lp in (a, b):
        lp.Begin(); lp.set_reparametrization("anisotropic")
        for it in range(3):
            lp.ComputePass(it)
    assert np.array_equal(a.unaries(), b.unaries())
    for x, y in zip(a.messages(), b.messages()):
        assert np.array_equal(x, y)


def test_error_behaviour():
    from lp_mp_b200 import LpmpError
    lp = oracle_lp("f64")
    lp.add_unaries(np.zeros((3, 2)))
    with pytest.raises(LpmpError):
        lp.add_pairwise_potts([0], [7], [1.0])     # endpoint out of range
    lp.add_pairwise_potts([0, 1], [1, 2], [1.0, 1.0])
    with pytest.raises(LpmpError):
        lp.ComputePass(0)                          # not finalized
    lp.finalize()
    lp.Begin()
    with pytest.raises(LpmpError, match="no reparametrization mode set"):  # include/LP_MP.h:458
        lp.ComputePass(0)
    with pytest.raises(LpmpError):
        lp.add_unaries(np.zeros((1, 2)))           # mutation after finalize
    assert lp.EvaluatePrimal() == np.inf           # no rounding pass yet
