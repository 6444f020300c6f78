"""GPU parity of the general-graph path (node-record layout, kernels_graph.cuh) beyond what tests/test_gpu_parity.py covers:
mixed label counts, the dual archive (8f-3), and the reference's UNMODIFIED Solver / MpRoundingSolver /
StandardVisitor driving the product's C++ binding include/LP_cuda.hxx (8b).  All calls go through the C ABI."""
import numpy as np
import pytest

import lp_mp_b200
from checkers import RefLP, grid_edges, oracle_lp, ref_library
from test_dual_archive import build, mixed_instance, same_state

pytestmark = pytest.mark.gpu
MODES = ["anisotropic", "uniform", "damped_uniform"]


def cuda_lp(dtype="f64"):
    return lp_mp_b200.LP(dtype=dtype, device=0)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("edge_order", ["lifo", "fifo"])
def test_mixed_label_counts_bit_parity(dtype, mode, edge_order):
    """variables with 2..5 labels in one LP (test/graphical_model.cpp:74-96): theta / messages / labels bit-identical to the oracle"""
    theta, edges, tables = mixed_instance(23, n=60, m=130)
    lps = [oracle_lp(dtype), cuda_lp(dtype)]
    for lp in lps:
        build(lp, theta, edges, tables)
        lp.set_edge_order(edge_order)
        lp.finalize(); lp.Begin(); lp.set_reparametrization(mode)
    o, c = lps
    assert np.isclose(c.LowerBound(), o.LowerBound(), rtol=1e-12 if dtype == "f64" else 1e-6)
    for it in range(4):
        o.ComputePass(it); c.ComputePass(it)
        assert same_state(o, c), it
        assert np.isclose(c.LowerBound(), o.LowerBound(), rtol=1e-12 if dtype == "f64" else 1e-6)
    for it in (4, 5):
        o.ComputeForwardPassAndPrimal(it); c.ComputeForwardPassAndPrimal(it)
        assert np.array_equal(o.labeling(), c.labeling())
        o.ComputeBackwardPassAndPrimal(it); c.ComputeBackwardPassAndPrimal(it)
        assert np.array_equal(o.labeling(), c.labeling())
        assert np.isclose(c.EvaluatePrimal(), o.EvaluatePrimal(), rtol=1e-12 if dtype == "f64" else 1e-6)
        assert c.CheckPrimalConsistency()
        assert same_state(o, c)


def test_mixed_label_counts_against_the_reference_core(ref_lib):
    theta, edges, tables = mixed_instance(31, n=30, m=64)
    ref, c = RefLP(ref_lib, ["-v", "0"]), cuda_lp("f64")
    build(ref, theta, edges, tables); build(c, theta, edges, tables)
    ref.Begin(); c.Begin()
    ref.set_reparametrization("anisotropic"); c.set_reparametrization("anisotropic")
    for it in range(5):
        ref.ComputePass(it); c.ComputePass(it)
        assert same_state(ref, c)
        assert np.isclose(c.LowerBound(), ref.LowerBound(), rtol=1e-12)
    ref.ComputeForwardPassAndPrimal(5); c.ComputeForwardPassAndPrimal(5)
    assert np.array_equal(ref.labeling(), c.labeling())


@pytest.mark.parametrize("L,deg_heavy", [(4, False), (8, False), (8, True), (16, False), (5, False), (32, False)])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_fast_and_generic_node_kernels(dtype, L, deg_heavy):
    """L in {4, 8, 16} with at most 4 incident edges runs the register kernel, everything else (other label counts, hubs of
    higher degree inside the same level) the generic one: both bit-identical to the oracle, all three reparametrisations,
    natural order and colouring order"""
    rng = np.random.default_rng(L * 7 + deg_heavy)
    n = 300
    perm = rng.permutation(n)
    edges = {(min(i, (i + 1) % n), max(i, (i + 1) % n)) for i in range(n)}
    for k in range(0, n, 2):
        i, j = int(perm[k]), int(perm[k + 1])
        edges.add((min(i, j), max(i, j)))
    if deg_heavy:   # a few hubs with more than 4 incident edges
        for hub in (3, 100, 250):
            for t in rng.choice(n, size=7, replace=False):
                if int(t) != hub:
                    edges.add((min(hub, int(t)), max(hub, int(t))))
    edges = sorted(edges)
    ei = np.array([e[0] for e in edges], dtype=np.uint64)
    ej = np.array([e[1] for e in edges], dtype=np.uint64)
    theta, tables = rng.random((n, L)), rng.random((len(edges), L, L)) - 0.2
    for use_colouring in (False, True):
        for mode in MODES:
            lps = [oracle_lp(dtype), cuda_lp(dtype)]
            for lp in lps:
                lp.add_unaries(theta); lp.add_pairwise_dense(ei, ej, tables)
            if use_colouring:
                colour, nc = lps[1].greedy_colouring(ei, ej)
                rank = colour.astype(np.uint64) * n + np.arange(n, dtype=np.uint64)
                for lp in lps:
                    lp.set_node_ranks(rank)
            for lp in lps:
                lp.finalize(); lp.Begin(); lp.set_reparametrization(mode)
            o, c = lps
            for it in range(3):
                o.ComputePass(it); c.ComputePass(it)
                assert same_state(o, c), (use_colouring, mode, it)
                assert np.isclose(c.LowerBound(), o.LowerBound(), rtol=1e-12 if dtype == "f64" else 1e-6)
            o.ComputeBackwardPassAndPrimal(3); c.ComputeBackwardPassAndPrimal(3)
            assert np.array_equal(o.labeling(), c.labeling())
            assert np.isclose(c.EvaluatePrimal(), o.EvaluatePrimal(), rtol=1e-12 if dtype == "f64" else 1e-6)


@pytest.mark.parametrize("shape", ["graph", "mixed", "grid_wavefront", "grid_dense"])
def test_dual_archive_on_device(shape):
    """lpmp_save_dual writes the reference's archive layout (== the oracle's bytes, which tests/test_dual_archive.py pins against
    the reference's own save_archive); save -> new handle -> load resumes with a bit-identical trajectory"""
    def make(lp):
        if shape == "mixed":
            build(lp, *mixed_instance(5, n=40, m=90))
        elif shape == "graph":
            rng = np.random.default_rng(4)
            ei, ej = grid_edges(6, 7)
            lp.add_unaries(rng.random((42, 8))); lp.add_pairwise_dense(ei, ej, rng.random((len(ei), 8, 8)))
        elif shape == "grid_wavefront":
            lp.grid_create(19, 23, 16, "potts"); lp.grid_fill_synthetic(7, -0.2, 1.0)
        else:
            lp.grid_create(9, 11, 16, "dense"); lp.grid_fill_synthetic(7, 0.0, 1.0)
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
        return lp
    o, c = make(oracle_lp("f64")), make(cuda_lp("f64"))
    for it in range(2):
        o.ComputePass(it); c.ComputePass(it)
    blob = c.save_dual()
    assert blob == o.save_dual()
    resumed = make(cuda_lp("f64"))
    resumed.load_dual(blob)
    assert np.isclose(resumed.LowerBound(), c.LowerBound(), rtol=1e-14)   # (a fused bound sums in another order than the sweep)
    for it in range(2, 5):
        c.ComputePass(it); resumed.ComputePass(it); o.ComputePass(it)
        assert same_state(c, resumed) and same_state(o, c)
        assert resumed.LowerBound() == c.LowerBound()   # from here on both took the same route to the bound
    with pytest.raises(lp_mp_b200.LpmpError):
        resumed.load_dual(blob[8:])


# ---- include/LP_cuda.hxx under the reference's own solver -----------------------------------------------------------------------
def adapter_pair(theta, edges, tables, options):
    lib_ref, lib_cuda = ref_library(), ref_library(cuda_adapter=True)
    if lib_ref is None or lib_cuda is None:
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    ref, cu = RefLP(lib_ref, options), RefLP(lib_cuda, options)
    build(ref, theta, edges, tables); build(cu, theta, edges, tables)
    return ref, cu


@pytest.mark.parametrize("options", [
    ["--maxIter", "23", "-v", "0"],                                                    # StandardVisitor defaults: damped-uniform rounding every 5th iteration
    ["--maxIter", "20", "--roundingReparametrization", "anisotropic", "--primalComputationInterval", "3", "-v", "0"],
    ["--maxIter", "60", "--minDualImprovement", "1e-3", "--minDualImprovementInterval", "4", "--lowerBoundComputationInterval", "2", "-v", "0"],
])
@pytest.mark.parametrize("shape", ["uniform", "mixed"])
def test_reference_solver_drives_lp_cuda(options, shape):
    """MpRoundingSolver<Solver<LP_cuda, StandardVisitor>>::Solve() (the reference's unmodified include/solver.hxx:230-257,381-400 and
    include/visitors/standard_visitor.hxx:48-199 compiled against include/LP_cuda.hxx) == the same Solve() over the reference's own
    LP<FMC>: every visit reports the same lower bound, primal bound and control flags; same final labeling and state"""
    if shape == "mixed":
        theta, edges, tables = mixed_instance(77, n=36, m=80)
    else:
        rng = np.random.default_rng(12)
        ei, ej = grid_edges(7, 8)
        edges = list(zip(ei.tolist(), ej.tolist()))
        theta, tables = list(rng.random((56, 6))), list(rng.random((len(edges), 6, 6)) - 0.25)
    ref, cu = adapter_pair(theta, edges, tables, options)
    r, c = ref.solve(), cu.solve()
    assert len(r["lb"]) == len(c["lb"]) and np.array_equal(r["flags"], c["flags"])
    assert np.allclose(c["lb"], r["lb"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(np.isinf(r["ub"]), np.isinf(c["ub"]))
    fin = ~np.isinf(r["ub"])
    assert np.allclose(c["ub"][fin], r["ub"][fin], rtol=1e-12)
    assert np.isclose(c["final_lb"], r["final_lb"], rtol=1e-12) and np.isclose(c["final_ub"], r["final_ub"], rtol=1e-12)
    assert np.array_equal(ref.labeling(), cu.labeling())
    assert same_state(ref, cu)
    # the archive of the binding is the reference's archive
    assert cu.save_dual() == ref.save_dual()


# ---- Anisotropic2 weights and residual sends ------------------------------------------------------------------------
ALL_MODES = ["anisotropic", "anisotropic2", "uniform", "damped_uniform"]


@pytest.mark.parametrize("send", ["shared", "residual"])
@pytest.mark.parametrize("mode", ALL_MODES)
@pytest.mark.parametrize("shape", ["mixed", "graph8", "graph5_potts", "grid_levels", "grid_wavefront", "grid_dense"])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_send_variants_on_device(dtype, send, mode, shape):
    """every reparametrisation mode x {shared, residual} on the register kernel (graph8), the generic kernel (mixed counts, Potts),
    and grids (level schedule; grids built for the wavefront kernels switch to it for Anisotropic2 / residual): bit-identical to
    the oracle, which tests/test_send_modes.py pins against the reference core"""
    from test_send_modes import lockstep
    def make(lp):
        rng = np.random.default_rng(41)
        if shape == "mixed":
            build(lp, *mixed_instance(3, n=50, m=110))
        elif shape == "graph8":
            ei, ej = grid_edges(7, 9)
            lp.add_unaries(rng.random((63, 8))); lp.add_pairwise_dense(ei, ej, rng.random((len(ei), 8, 8)) - 0.2)
        elif shape == "graph5_potts":
            ei, ej = grid_edges(6, 5)
            lp.add_unaries(rng.random((30, 5))); lp.add_pairwise_potts(ei, ej, rng.uniform(-0.4, 1.0, len(ei)))
        elif shape == "grid_levels":
            lp.grid_create(7, 6, 5, "potts"); lp.grid_fill_synthetic(3, -0.2, 1.0)
        elif shape == "grid_wavefront":
            lp.grid_create(19, 21, 16, "trunclin"); lp.grid_fill_synthetic(3, 0.25, 1.0)
        else:
            lp.grid_create(6, 9, 16, "dense"); lp.grid_fill_synthetic(3, 0.0, 1.0)
        lp.finalize(); lp.Begin(); lp.set_send_mode(send)
        return lp
    lockstep(make(oracle_lp(dtype)), [make(cuda_lp(dtype))], mode, iters=3)


def test_reference_solver_drives_lp_cuda_with_residual_sends():
    theta, edges, tables = mixed_instance(19, n=30, m=66)
    options = ["--maxIter", "17", "--reparametrizationType", "residual", "--standardReparametrization", "anisotropic2", "-v", "0"]
    ref, cu = adapter_pair(theta, edges, tables, options)
    r, c = ref.solve(), cu.solve()
    assert np.array_equal(r["flags"], c["flags"]) and np.allclose(c["lb"], r["lb"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(ref.labeling(), cu.labeling()) and same_state(ref, cu)
    with pytest.raises(lp_mp_b200.LpmpError):
        cuda_lp("f64").set_send_mode("adaptive")


# ---- staged optimisation: partition / overlapping-partition passes --------------------------------------------------
def _row_pair_partitions(H, W):
    """put_in_same_partition calls that make every pair of grid rows one partition"""
    pi, pj = [], []
    for r in range(H):
        for c in range(W):
            n = r * W + c
            if c + 1 < W:
                pi.append(n); pj.append(n + 1)
            if r % 2 == 0 and r + 1 < H:
                pi.append(n); pj.append(n + W)
    return pi, pj


@pytest.mark.parametrize("overlapping", [False, True])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_partition_passes_on_device(dtype, overlapping):
    """LP::compute_partition_pass / compute_overlapping_partition_pass (include/LP_MP.h:1961-2070): bit-identical to the oracle, which
    tests/test_partition_passes.py pins against the reference core"""
    rng = np.random.default_rng(5)
    H, W, L = 8, 7, 8
    ei, ej = grid_edges(H, W)
    theta, tables = rng.random((H * W, L)), rng.random((len(ei), L, L)) - 0.3
    pi, pj = _row_pair_partitions(H, W)
    lps = [oracle_lp(dtype), cuda_lp(dtype)]
    for lp in lps:
        lp.add_unaries(theta); lp.add_pairwise_dense(ei, ej, tables)
        lp.put_in_same_partition(pi, pj)
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
    o, c = lps
    prev = c.LowerBound()
    for it in range(3):
        for lp in lps:
            lp.compute_partition_pass(2, overlapping)
            if overlapping:
                lp.ComputeForwardPass(); lp.ComputeBackwardPass()
        assert same_state(o, c), it
        lb = c.LowerBound()
        assert np.isclose(lb, o.LowerBound(), rtol=1e-12 if dtype == "f64" else 1e-6) and lb >= prev - 1e-6 * abs(prev)
        prev = lb
    # changing the partition afterwards rebuilds the sweeps
    for lp in lps:
        lp.put_in_same_partition([0], [H * W - 1])
        lp.compute_partition_pass(1, overlapping)
    assert same_state(o, c)


def test_partition_passes_mixed_labels_against_the_reference_core(ref_lib):
    theta, edges, tables = mixed_instance(13, n=36, m=80)
    pi = [k for k in range(35) if k % 6 != 5]
    pj = [k + 1 for k in pi]   # six partitions of six consecutive variables
    ref = RefLP(ref_lib, ["-v", "0", "--reparametrizationType", "partition", "--innerIteration", "3"])
    c = cuda_lp("f64")
    build(ref, theta, edges, tables); build(c, theta, edges, tables)
    ref.put_in_same_partition(pi, pj); c.put_in_same_partition(pi, pj)
    ref.Begin(); c.Begin(); ref.set_reparametrization("anisotropic"); c.set_reparametrization("anisotropic")
    for it in range(3):
        ref.ComputePass(it); c.compute_partition_pass(3, False)
        assert same_state(ref, c) and np.isclose(c.LowerBound(), ref.LowerBound(), rtol=1e-12)


@pytest.mark.parametrize("kind", ["partition", "overlapping_partition"])
def test_reference_solver_drives_lp_cuda_in_partition_mode(kind):
    """--reparametrizationType partition / overlapping_partition: Solver::Solve() over LP_cuda == over the reference's LP<FMC>"""
    theta, edges, tables = mixed_instance(29, n=30, m=64)
    options = ["--maxIter", "9", "--reparametrizationType", kind, "--innerIteration", "2", "--roundingReparametrization", "anisotropic", "-v", "0"]
    ref, cu = adapter_pair(theta, edges, tables, options)
    pi = [k for k in range(29) if k % 5 != 4]
    pj = [k + 1 for k in pi]
    ref.put_in_same_partition(pi, pj); cu.put_in_same_partition(pi, pj)
    r, c = ref.solve(), cu.solve()
    assert np.array_equal(r["flags"], c["flags"]) and np.allclose(c["lb"], r["lb"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(ref.labeling(), cu.labeling()) and same_state(ref, cu)


def test_lp_cuda_factor_relations_reorient_edges():
    """LP_cuda::AddFactorRelation(u_j, p), (p, u_i) turns edge (i, j) around: same dual trajectory as the reference core built with
    the edge given as (j, i) and the transposed table (the per-node message order differs, so bounds agree to round-off, not bits)"""
    import ctypes as C
    lib_ref, lib_cuda = ref_library(), ref_library(cuda_adapter=True)
    if lib_ref is None or lib_cuda is None:
        pytest.skip("oracle/_ref not built")
    lib_cuda.lpmp_ref_reverse_edge.argtypes = [C.c_void_p, C.c_uint64]
    lib_cuda.lpmp_ref_reverse_edge.restype = None
    rng = np.random.default_rng(6)
    ei, ej = grid_edges(5, 6)
    edges = list(zip(ei.tolist(), ej.tolist()))
    theta, tables = list(rng.random((30, 4))), list(rng.random((len(edges), 4, 4)) - 0.3)
    flip = [e for e, (i, j) in enumerate(edges) if j == i + 1]   # all horizontal edges point left: still a DAG
    ref, cu = RefLP(lib_ref, ["-v", "0"]), RefLP(lib_cuda, ["-v", "0"])
    for t in theta:
        ref.AddUnaryFactor(t); cu.AddUnaryFactor(t)
    for e, ((i, j), t) in enumerate(zip(edges, tables)):
        cu.AddPairwiseFactor(i, j, t)
        if e in flip:
            ref.AddPairwiseFactor(j, i, np.ascontiguousarray(t.T))
            lib_cuda.lpmp_ref_reverse_edge(cu.h, e)
        else:
            ref.AddPairwiseFactor(i, j, t)
    ref.Begin(); cu.Begin()
    ref.set_reparametrization("anisotropic"); cu.set_reparametrization("anisotropic")
    for it in range(6):
        ref.ComputePass(it); cu.ComputePass(it)
        assert np.isclose(cu.LowerBound(), ref.LowerBound(), rtol=1e-11), it
    assert np.allclose(ref.unaries(), cu.unaries(), atol=1e-10)


# ---- single-factor entry point on the HOT kernels (lpmp_node_update) ------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("kind,L,deg", [("dense", 8, 3), ("dense", 4, 4), ("dense", 16, 2), ("dense", 8, 6), ("dense", 5, 3), ("potts", 12, 4), ("trunclin", 16, 3)])
def test_node_update_entry_point(cuda_api, oracle_api, dtype, kind, L, deg):
    """one UpdateFactor call per test node, run by the kernels the general-graph passes launch (the register kernel for dense L in
    {4, 8, 16} with deg <= 4, the generic kernel otherwise): theta / own-side vectors / label bit-identical to the oracle for every
    weight mode and both directions; the anisotropic receive agrees with the reference's min-marginals within 1e-5 (north_star (i))"""
    import ctypes as C
    from lp_mp_b200 import _capi as K
    rng = np.random.default_rng(L * 31 + deg)
    n = 257
    P = {"dense": L * L, "potts": 1, "trunclin": 2}[kind]
    theta, own, other = rng.random((n, L)), rng.random((n, deg, L)) - 0.5, rng.random((n, deg, L))
    if kind == "dense":
        params = rng.random((n, deg, P)) - 0.2
    elif kind == "potts":
        params = rng.uniform(-0.4, 1.0, (n, deg, 1))
    else:
        params = np.stack([rng.uniform(0.05, 0.5, (n, deg)), rng.uniform(0.2, 1.5, (n, deg))], axis=-1)
    later = (rng.random((n, deg)) < 0.5).astype(np.uint8)
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    def run(api, repam, forward):
        th, ow, lab = np.zeros((n, L)), np.zeros((n, deg, L)), np.zeros(n, dtype=np.uint32)
        st = api.node_update(K.F64 if dtype == "f64" else K.F32, n, L, deg, K.PAIRWISE[kind], repam, forward, later.ctypes.data_as(C.POINTER(C.c_uint8)),
                             pd(theta), pd(own), pd(other), pd(params), pd(th), pd(ow), lab.ctypes.data_as(C.POINTER(C.c_uint32)))
        assert st == 0
        return th, ow, lab
    for repam in range(4):
        for forward in (1, 0):
            got, want = run(cuda_api, repam, forward), run(oracle_api, repam, forward)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), (repam, forward)
    # the receive of one message == own + min-marginal of the reference's arithmetic (lpmp_minmarginal_*: matrix::min1)
    if kind == "dense":
        th, ow, _ = run(cuda_api, 1, 1)     # uniform: every message is received
        h = (params.reshape(n, deg, L, L) + other[:, :, None, :]).min(axis=3)
        m = h + own
        sent = (theta + m.sum(axis=1)) / deg
        assert np.allclose(ow, own - m + sent[:, None, :], rtol=1e-5 if dtype == "f32" else 1e-12, atol=1e-5 if dtype == "f32" else 1e-12)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_graph_fused_bound_equals_the_sweep(dtype):
    """anisotropic backward passes of the register kernel leave LP::LowerBound behind (one wide sum instead of a sweep over all
    records); any other pass or state change falls back to the sweep"""
    rng = np.random.default_rng(77)
    n, L = 500, 8
    perm = rng.permutation(n)
    edges = sorted({(min(i, (i + 1) % n), max(i, (i + 1) % n)) for i in range(n)} | {(min(int(perm[k]), int(perm[k + 1])), max(int(perm[k]), int(perm[k + 1]))) for k in range(0, n, 2)})
    ei, ej = np.array([e[0] for e in edges], dtype=np.uint64), np.array([e[1] for e in edges], dtype=np.uint64)
    theta, tables = rng.random((n, L)), rng.random((len(edges), L, L)) - 0.2
    lps = []
    for fused in (1, 0):
        lp = cuda_lp(dtype)
        lp.set_tuning(graph_fused_bound=fused)
        lp.add_unaries(theta); lp.add_pairwise_dense(ei, ej, tables)
        colour, nc = lp.greedy_colouring(ei, ej)
        lp.set_node_ranks(colour.astype(np.uint64) * n + np.arange(n, dtype=np.uint64))
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
        lps.append(lp)
    f, s = lps
    for it in range(3):
        f.ComputePass(it); s.ComputePass(it)
        before = f.stats()["kernel_launches"]
        lb = f.LowerBound()
        assert f.stats()["kernel_launches"] - before == 1, "one wide sum, no sweep"
        assert np.isclose(lb, s.LowerBound(), rtol=1e-13 if dtype == "f64" else 1e-6)
        assert same_state(f, s)
    f.ComputeForwardPass(); s.ComputeForwardPass()          # a forward pass invalidates the terms
    assert np.isclose(f.LowerBound(), s.LowerBound(), rtol=1e-13 if dtype == "f64" else 1e-6)
    f.ComputeBackwardPass(); s.ComputeBackwardPass()
    th = f.unaries(); th[0, 0] += 1.0
    f.set_unaries(th); s.set_unaries(th)                      # ... and so does a state change
    assert np.isclose(f.LowerBound(), s.LowerBound(), rtol=1e-13 if dtype == "f64" else 1e-6)
    f.set_reparametrization("uniform"); s.set_reparametrization("uniform")
    f.ComputePass(9); s.ComputePass(9)
    assert np.isclose(f.LowerBound(), s.LowerBound(), rtol=1e-13 if dtype == "f64" else 1e-6) and same_state(f, s)
