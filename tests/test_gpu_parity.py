"""GPU parity tests: the CUDA library, called through the C ABI (include/lpmp_b200.h), against
  (1) the committed fixtures recorded from the reference's own scheduler (tests/golden),
  (2) the CPU oracle on the same seeded inputs -- bit-exact state (theta, a, b, labels) in f64 AND f32,
  (3) size-independent properties at BASELINE.json's full sizes.
Tolerances: state and labelings bit-exact; lower bound / primal cost relative 1e-12 (f64) resp. 1e-6 (f32) because the
global sums are accumulated in a different (deterministic tree) order than the reference's sequential sum.
"""
import ctypes as C

import numpy as np
import pytest

from checkers import grid_edges, oracle_lp, potts_table, trunclin_table
from common import GOLDEN_CASES, build_from_golden, check_against_golden, load_golden, run_lockstep, same_state

pytestmark = pytest.mark.gpu

MODES = ["anisotropic", "uniform", "damped_uniform"]
LB_RTOL = {"f64": 1e-12, "f32": 1e-6}


def cuda_lp(dtype="f64"):
    import lp_mp_b200
    return lp_mp_b200.LP(dtype=dtype, device=0)


def _pd(x):
    return None if x is None else x.ctypes.data_as(C.POINTER(C.c_double))


# ---- (1) committed fixtures from the reference core ------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("native", [True, False])
def test_cuda_reproduces_reference_fixture(name, native):
    g = load_golden(name)
    check_against_golden(build_from_golden(cuda_lp("f64"), g, native=native), g, exact=True)


@pytest.mark.parametrize("name", ["grid_potts", "grid_trunclin"])
@pytest.mark.parametrize("schedule", ["levels", "wavefront"])
def test_cuda_grid_api_reproduces_reference_fixture(name, schedule):
    g = load_golden(name)
    lp = build_from_golden(cuda_lp("f64"), g, use_grid_api=True)
    lp.set_schedule(schedule)
    check_against_golden(lp, g, exact=True)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cuda_f32_tracks_reference_fixture(name):
    g = load_golden(name)
    check_against_golden(build_from_golden(cuda_lp("f32"), g), g, exact=False, rtol=1e-5)


# ---- single-factor kernels (north_star: 1e-5 relative; we get bit-exact) -----------------------------------------
@pytest.mark.parametrize("dtype", [0, 1])
def test_single_factor_kernels(cuda_api, oracle_api, dtype):
    rng = np.random.default_rng(1)
    n, L1, L2 = 37, 5, 7
    t, m1, m2 = rng.normal(size=(n, L1, L2)), rng.normal(size=(n, L1)), rng.normal(size=(n, L2))
    for side in (0, 1):
        outs = []
        for api in (cuda_api, oracle_api):
            out, lb = np.zeros((n, L1 if side == 0 else L2)), np.zeros(n)
            api.check(api.minmarginal_dense(dtype, n, L1, L2, _pd(t), _pd(m1), _pd(m2), side, _pd(out), _pd(lb)))
            outs.append((out, lb))
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    L = 9
    lam, slope, trunc = rng.uniform(-1, 1, size=n), rng.uniform(0, 0.5, size=n), rng.uniform(0.2, 2.0, size=n)
    m1, m2 = rng.normal(size=(n, L)), rng.normal(size=(n, L))
    for side in (0, 1):
        res = []
        for api in (cuda_api, oracle_api):
            o1, l1, o2, l2 = np.zeros((n, L)), np.zeros(n), np.zeros((n, L)), np.zeros(n)
            api.check(api.minmarginal_potts(dtype, n, L, _pd(lam), _pd(m1), _pd(m2), side, _pd(o1), _pd(l1)))
            api.check(api.minmarginal_trunclin(dtype, n, L, _pd(slope), _pd(trunc), _pd(m1), _pd(m2), side, _pd(o2), _pd(l2)))
            res.append((o1, l1, o2, l2))
        for x, y in zip(*res):
            assert np.array_equal(x, y)
    theta = rng.normal(size=(n, 11))
    theta[3, 4] = theta[3, 9] = theta[3].min() - 1.0  # tie: first minimum wins
    res = []
    for api in (cuda_api, oracle_api):
        lb, arg = np.zeros(n), np.zeros(n, dtype=np.uint32)
        api.check(api.unary_min(dtype, n, 11, _pd(theta), _pd(lb), arg.ctypes.data_as(C.POINTER(C.c_uint32))))
        res.append((lb, arg))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and res[0][1][3] == 4


def test_reference_goldens_on_device(cuda_api):
    """test/simplex.cpp:8-11,52-64 and test/simplex_marginalization.cpp:9-41 through the CUDA kernels"""
    theta = np.array([[0.1, 0.2, 0.05, 1.0]])
    lb, arg = np.zeros(1), np.zeros(1, dtype=np.uint32)
    cuda_api.check(cuda_api.unary_min(1, 1, 4, _pd(theta), _pd(lb), arg.ctypes.data_as(C.POINTER(C.c_uint32))))
    assert lb[0] == 0.05
    t = np.zeros((3, 3))
    for x in range(3):
        t[x, x] = -float(x) - 1.0
    out, lb = np.zeros(3), np.zeros(1)
    cuda_api.check(cuda_api.minmarginal_dense(1, 1, 3, 3, _pd(t), None, None, 0, _pd(out), _pd(lb)))
    assert lb[0] == -3.0
    cost = np.array([[0.1, 0.2, 0.05], [0.3, 0.001, 0.2], [-0.3, -0.001, -0.2], [0.3, 0.001, 0.2]])
    out = np.zeros(4)
    cuda_api.check(cuda_api.minmarginal_dense(1, 1, 4, 3, _pd(cost), None, None, 0, _pd(out), None))
    assert list(-out) == [-0.05, -0.001, 0.3, -0.001]
    out = np.zeros(3)
    cuda_api.check(cuda_api.minmarginal_dense(1, 1, 4, 3, _pd(cost), None, None, 1, _pd(out), None))
    assert list(-out) == [0.3, 0.001, 0.2]
    for lam in (1.0, -1.0):  # test/potts_factor.cpp:34-66
        m1, m2 = np.array([[-0.1, 0.5, 0.8]]), np.array([[1.5, 1.0, 0.0]])
        for side in (0, 1):
            d, p, ld, lp_ = np.zeros(3), np.zeros(3), np.zeros(1), np.zeros(1)
            cuda_api.check(cuda_api.minmarginal_dense(1, 1, 3, 3, _pd(potts_table(3, lam)), _pd(m1), _pd(m2), side, _pd(d), _pd(ld)))
            cuda_api.check(cuda_api.minmarginal_potts(1, 1, 3, _pd(np.array([lam])), _pd(m1), _pd(m2), side, _pd(p), _pd(lp_)))
            assert np.array_equal(d, p) and ld[0] == lp_[0]


# ---- (2) general graphs, level schedule ---------------------------------------------------------------------------
def random_graph(rng, n, m):
    edges = set()
    while len(edges) < m:
        i, j = rng.integers(0, n, size=2)
        if i != j:
            edges.add((int(min(i, j)), int(max(i, j))))
    edges = sorted(edges)
    rng.shuffle(edges)
    return np.array([e[0] for e in edges], dtype=np.uint64), np.array([e[1] for e in edges], dtype=np.uint64)


def fill(lp, kind, theta, ei, ej, rng_seed, L):
    rng = np.random.default_rng(rng_seed)
    lp.add_unaries(theta)
    if kind == "dense":
        lp.add_pairwise_dense(ei, ej, rng.random((len(ei), L, L)) - 0.2)
    elif kind == "potts":
        lp.add_pairwise_potts(ei, ej, rng.uniform(-0.5, 1.0, size=len(ei)))
    else:
        lp.add_pairwise_trunclin(ei, ej, rng.uniform(0.05, 0.5, size=len(ei)), rng.uniform(0.2, 1.5, size=len(ei)))
    return lp


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("kind,L", [("dense", 2), ("dense", 5), ("dense", 8), ("potts", 3), ("potts", 32), ("potts", 40),
                                    ("trunclin", 7), ("trunclin", 16), ("dense", 64)])
def test_general_graph_bit_parity(dtype, mode, kind, L):
    rng = np.random.default_rng(L * 7 + len(mode))
    n = 60
    ei, ej = random_graph(rng, n, 130)
    theta = rng.random((n, L))
    lps = [fill(oracle_lp(dtype), kind, theta, ei, ej, 5, L), fill(cuda_lp(dtype), kind, theta, ei, ej, 5, L)]
    run_lockstep(lps, mode, iters=4, lb_rtol=LB_RTOL[dtype])


@pytest.mark.parametrize("edge_order", ["lifo", "fifo"])
def test_edge_order_parameter(edge_order):
    rng = np.random.default_rng(9)
    ei, ej = grid_edges(5, 7)
    theta = rng.random((35, 6))
    lps = [fill(oracle_lp("f64"), "dense", theta, ei, ej, 2, 6), fill(cuda_lp("f64"), "dense", theta, ei, ej, 2, 6)]
    for lp in lps:
        lp.set_edge_order(edge_order)
    run_lockstep(lps, "damped_uniform", iters=3)


def test_degenerate_graphs():
    """isolated nodes, a single edge, a star (degree 12), duplicate-free multi-component graph"""
    rng = np.random.default_rng(2)
    theta = rng.random((14, 4))
    ei = np.array([0] * 12 + [12], dtype=np.uint64)          # node 13 stays isolated
    ej = np.array(list(range(1, 13)) + [1], dtype=np.uint64)
    ej[12] = 11
    ei[12] = 12
    for mode in MODES:
        lps = [fill(oracle_lp("f64"), "dense", theta, ei, ej, 3, 4), fill(cuda_lp("f64"), "dense", theta, ei, ej, 3, 4)]
        run_lockstep(lps, mode, iters=3)
    lp = cuda_lp("f64")          # no edges at all
    lp.add_unaries(theta)
    lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
    lp.ComputePass(0)
    assert np.isclose(lp.LowerBound(), theta.min(axis=1).sum(), rtol=1e-14)
    lp.ComputeForwardPassAndPrimal(1)
    assert np.array_equal(lp.labeling(), theta.argmin(axis=1))


def test_colouring_schedule():
    """graph-colouring schedule (BASELINE config 4 shape, small): exact parity with the oracle run on the same
    colour ranks; the final bound of the natural-order device run equals the oracle's (<= 1e-12, criterion
    1e-4), the colour-order run's gap to it is measured and bounded by 2e-3 (see the comment at the assertion)."""
    rng = np.random.default_rng(17)
    n, L = 400, 8
    # random 3-regular-ish graph: union of a ring and a random perfect matching
    perm = rng.permutation(n)
    edges = {(min(i, (i + 1) % n), max(i, (i + 1) % n)) for i in range(n)}
    for k in range(0, n, 2):
        i, j = int(perm[k]), int(perm[k + 1])
        edges.add((min(i, j), max(i, j)))
    edges = sorted(edges)
    ei, ej = np.array([e[0] for e in edges], dtype=np.uint64), np.array([e[1] for e in edges], dtype=np.uint64)
    theta = rng.random((n, L))
    cu = fill(cuda_lp("f64"), "dense", theta, ei, ej, 4, L)
    colour, nc = cu.greedy_colouring(ei, ej)
    assert nc <= 4 and all(colour[i] != colour[j] for i, j in zip(ei, ej))
    rank = colour.astype(np.uint64) * n + np.arange(n, dtype=np.uint64)
    orc = fill(oracle_lp("f64"), "dense", theta, ei, ej, 4, L)
    ocol, onc = orc.greedy_colouring(ei, ej)
    assert np.array_equal(ocol, colour) and onc == nc
    cu.set_node_ranks(rank); orc.set_node_ranks(rank)
    run_lockstep([orc, cu], "anisotropic", iters=5)
    assert cu.stats()["num_levels"] <= nc
    nat = fill(oracle_lp("f64"), "dense", theta, ei, ej, 4, L)
    nat.finalize(); nat.Begin(); nat.set_reparametrization("anisotropic")
    for it in range(5, 200):
        cu.ComputePass(it)
    for it in range(200):
        nat.ComputePass(it)
    # SRMP fixed points depend on the orientation of the relation DAG, so the colour-oriented run and the index-oriented
    # run converge to different (both valid) bounds; on this instance they differ by 8e-4 relative, i.e. BASELINE.json's
    # 1e-4 is NOT met by the colouring schedule here (DESIGN.md 4.3 records the measured gaps; bench.py's C4 block reports
    # the gap of every run).  The colour-oriented ORACLE reaches exactly the device's bound (bit parity above): the gap is
    # a property of the schedule, not of the kernel.  What does meet the criterion -- exactly -- is the device run in the
    # reference's own order (level sets of the index-oriented DAG, a dozen levels on cubic graphs):
    gap = abs(cu.LowerBound() - nat.LowerBound()) / abs(nat.LowerBound())
    assert gap <= 2e-3, gap
    exact = fill(cuda_lp("f64"), "dense", theta, ei, ej, 4, L)
    exact.finalize(); exact.Begin(); exact.set_reparametrization("anisotropic")
    for it in range(200):
        exact.ComputePass(it)
    assert abs(exact.LowerBound() - nat.LowerBound()) <= 1e-12 * abs(nat.LowerBound())
    assert np.array_equal(exact.unaries(), nat.unaries())


# ---- grids: wavefront kernel ------------------------------------------------------------------------------------------
def grid_pair(dtype, H, W, L, kind, seed, schedule="wavefront", edge_order="lifo", **tuning):
    rng = np.random.default_rng(seed)
    theta = rng.random((H, W, L))
    if kind == "potts":
        hz, vt = rng.uniform(-0.3, 1.0, size=(H, W - 1, 1)), rng.uniform(-0.3, 1.0, size=(H - 1, W, 1))
    elif kind == "dense":
        hz, vt = rng.uniform(-0.5, 1.0, size=(H, W - 1, L * L)), rng.uniform(-0.5, 1.0, size=(H - 1, W, L * L))
    else:
        hz = np.stack([rng.uniform(0.05, 0.5, size=(H, W - 1)), rng.uniform(0.2, 1.5, size=(H, W - 1))], axis=-1)
        vt = np.stack([rng.uniform(0.05, 0.5, size=(H - 1, W)), rng.uniform(0.2, 1.5, size=(H - 1, W))], axis=-1)
    lps = []
    for make in (oracle_lp, cuda_lp):
        lp = make(dtype).grid_create(H, W, L, kind)
        lp.grid_set_unaries(theta)
        lp.grid_set_pairwise(hz, vt)
        lp.set_edge_order(edge_order)
        lps.append(lp)
    lps[1].set_schedule(schedule)
    if tuning:
        lps[1].set_tuning(**tuning)
    return lps


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("kind,L", [("potts", 8), ("potts", 32), ("trunclin", 16), ("potts", 64), ("trunclin", 8)])
def test_wavefront_bit_parity(dtype, mode, kind, L):
    lps = grid_pair(dtype, 37, 29, L, kind, seed=L + len(mode))
    run_lockstep(lps, mode, iters=3, lb_rtol=LB_RTOL[dtype])


@pytest.mark.parametrize("mode", ["uniform", "damped_uniform"])
@pytest.mark.parametrize("edge_order", ["lifo", "fifo"])
@pytest.mark.parametrize("kind,L,dtype", [("potts", 16, "f32"), ("potts", 32, "f32"), ("trunclin", 16, "f32"), ("potts", 16, "f64")])
def test_split_kernel_uniform_modes_bit_parity(mode, edge_order, kind, L, dtype):
    """the uniform / damped-uniform instances of the split wavefront kernel (grids with W % 8 == 0: rows of whole tiles, the later
    neighbours' vectors read from the next warp's tiles / the halo ring): bit parity with the oracle, rounding passes included;
    bands that do not divide H, a single band, one row, several bands per CTA, clusters of 1 / 2 CTAs"""
    for (H, W), tuning in (((37, 32), {}), ((8, 16), {}), ((1, 24), {}), ((9, 40), dict(wavefront_cluster=1)),
                           ((70, 48), dict(wavefront_max_ctas=2)), ((45, 64), dict(wavefront_cluster=2, wavefront_max_ctas=4))):
        lps = grid_pair(dtype, H, W, L, kind, seed=H + W + L, edge_order=edge_order, **tuning)
        run_lockstep(lps, mode, iters=3, lb_rtol=LB_RTOL[dtype])
        assert lps[1].stats()["kernel_launches"] <= 40, "pass kernels (8 launches + bound / primal sweeps), not the level schedule"


@pytest.mark.parametrize("kind,L,dtype", [("potts", 16, "f32"), ("potts", 32, "f32"), ("trunclin", 16, "f32"), ("trunclin", 16, "f64")])
def test_split_kernel_rows_of_whole_tiles(kind, L, dtype):
    """anisotropic passes on grids with W % 8 == 0: a row is tiles only (the tiles of column 0 and of the last column run the fast
    steps, hand-off by one generic store for the last row of a band); both message orders, rounding passes included"""
    for (H, W), tuning in (((37, 32), {}), ((8, 16), {}), ((1, 24), {}), ((9, 40), dict(wavefront_cluster=1)), ((2, 8), {}),
                           ((70, 48), dict(wavefront_max_ctas=2)), ((45, 64), dict(wavefront_cluster=2, wavefront_max_ctas=4))):
        for edge_order in ("lifo", "fifo"):
            lps = grid_pair(dtype, H, W, L, kind, seed=H + W + L, edge_order=edge_order, **tuning)
            run_lockstep(lps, "anisotropic", iters=3, lb_rtol=LB_RTOL[dtype])


def test_split_kernel_uniform_modes_match_the_one_warp_kernel():
    """same state from the split instances and from the one-warp wavefront kernel (wavefront_split_uni=0) on a larger grid"""
    rng = np.random.default_rng(5)
    H, W, L = 130, 256, 16
    theta = rng.random((H, W, L))
    hz = np.stack([rng.uniform(0.05, 0.5, size=(H, W - 1)), rng.uniform(0.2, 1.5, size=(H, W - 1))], axis=-1)
    vt = np.stack([rng.uniform(0.05, 0.5, size=(H - 1, W)), rng.uniform(0.2, 1.5, size=(H - 1, W))], axis=-1)
    lps = []
    for uni in (0, -1):
        lp = cuda_lp("f32").grid_create(H, W, L, "trunclin")
        lp.grid_set_unaries(theta); lp.grid_set_pairwise(hz, vt)
        lp.set_tuning(wavefront_split_uni=uni)
        lps.append(lp)
    run_lockstep(lps, "damped_uniform", iters=3, lb_rtol=1e-6)


@pytest.mark.parametrize("H,W", [(1, 1), (1, 9), (9, 1), (2, 2), (3, 50), (50, 3), (33, 33), (64, 17)])
def test_wavefront_ragged_shapes(H, W):
    """bands that do not divide H, grids narrower than a band, single rows / columns"""
    for kind, L in (("potts", 16), ("trunclin", 32)):
        lps = grid_pair("f64", H, W, L, kind, seed=H * 100 + W)
        run_lockstep(lps, "anisotropic", iters=3)
        lps = grid_pair("f64", H, W, L, kind, seed=H * 100 + W)
        run_lockstep(lps, "damped_uniform", iters=2)


@pytest.mark.parametrize("tuning", [dict(wavefront_max_ctas=1), dict(wavefront_max_ctas=3), dict(wavefront_ctas_per_sm=1),
                                    dict(wavefront_max_ctas=2, wavefront_ctas_per_sm=1), dict(wavefront_lean=0), dict(wavefront_lean=0, wavefront_max_ctas=2),
                                    dict(wavefront_warps_per_cta=1),
                                    dict(wavefront_warps_per_cta=2), dict(wavefront_warps_per_cta=3, wavefront_tile_cols=4),
                                    dict(wavefront_warps_per_cta=3, wavefront_max_ctas=1),
                                    dict(wavefront_warps_per_cta=8, wavefront_tile_cols=4, wavefront_max_ctas=2)])
def test_wavefront_launch_shapes_do_not_change_results(tuning):
    lps = grid_pair("f32", 70, 45, 16, "trunclin", seed=3, **tuning)
    run_lockstep(lps, "anisotropic", iters=3, lb_rtol=1e-6)
    lps = grid_pair("f64", 45, 70, 32, "potts", seed=4, **tuning)
    run_lockstep(lps, "uniform", iters=2)


@pytest.mark.parametrize("kind,L,tuning", [("trunclin", 16, dict(wavefront_tile_cols=4, wavefront_warps_per_cta=4)),
                                           ("trunclin", 16, dict(wavefront_tile_cols=8, wavefront_warps_per_cta=2)),
                                           ("potts", 32, dict(wavefront_warps_per_cta=8)),
                                           ("potts", 8, dict(wavefront_max_ctas=3, wavefront_warps_per_cta=3))])
def test_wavefront_super_bands(kind, L, tuning):
    """several consecutive bands per CTA: hand-off through the shared-memory rings (they wrap many times: W >> ring),
    partial last super-band, CTAs that take several super-bands"""
    lps = grid_pair("f32", 203, 157, L, kind, seed=21, **tuning)
    run_lockstep(lps, "anisotropic", iters=3, lb_rtol=1e-6)


@pytest.mark.parametrize("kind,L,dtype,tuning", [("trunclin", 16, "f32", dict(wavefront_lean=1)), ("trunclin", 16, "f32", dict(wavefront_cluster=1)),
                                                 ("trunclin", 16, "f32", dict(wavefront_cluster=2)), ("potts", 32, "f64", dict(wavefront_cluster=4, wavefront_max_ctas=8)),
                                                 ("potts", 8, "f32", dict(wavefront_cluster=8, wavefront_max_ctas=8)), ("potts", 16, "f32", dict(wavefront_cluster=4, wavefront_ctas_per_sm=1)), ("trunclin", 16, "f64", dict(wavefront_lean=1, wavefront_tile_cols=4)),
                                                 ("potts", 32, "f32", dict(wavefront_lean=1)), ("potts", 8, "f64", dict(wavefront_lean=1, wavefront_max_ctas=3)),
                                                 ("potts", 64, "f32", dict(wavefront_lean=1, wavefront_max_ctas=1))])
def test_wavefront_lean_variant(kind, L, dtype, tuning):
    """compute warp + mover warps per band (kernels_grid_lean.cuh): anisotropic passes and rounding, several bands per CTA
    in sequence, halo ring wrapping many times, ragged last band / tile"""
    lps = grid_pair(dtype, 203, 157, L, kind, seed=22, **tuning)
    run_lockstep(lps, "anisotropic", iters=3, lb_rtol=1e-6 if dtype == "f32" else 1e-12)
    for H, W in ((1, 40), (40, 1), (3, 3), (9, 2), (17, 70)):
        lps = grid_pair(dtype, H, W, L, kind, seed=H * 100 + W, **tuning)
        run_lockstep(lps, "anisotropic", iters=2, lb_rtol=1e-6 if dtype == "f32" else 1e-12)


@pytest.mark.parametrize("kind,L,dtype,tuning", [("trunclin", 16, "f32", dict(wavefront_rows=1)), ("trunclin", 16, "f64", dict(wavefront_rows=1)),
                                                 ("trunclin", 16, "f32", dict(wavefront_rows=1, wavefront_max_ctas=2)),
                                                 ("potts", 16, "f32", dict(wavefront_rows=1)), ("potts", 32, "f32", dict(wavefront_rows=1)),
                                                 ("potts", 32, "f64", dict(wavefront_rows=1, wavefront_max_ctas=3)),
                                                 ("potts", 8, "f32", dict(wavefront_rows=1)), ("trunclin", 8, "f64", dict(wavefront_rows=1)),
                                                 ("trunclin", 32, "f32", dict(wavefront_rows=1, wavefront_max_ctas=1))])
def test_wavefront_rows_variant(kind, L, dtype, tuning):
    """one label per lane (kernels_grid_rows.cuh): anisotropic passes and rounding; rings wrapping many times, several bands per
    CTA in sequence, ragged last band / tile, single rows / columns, negative Potts couplings (two smallest entries)"""
    rt = 1e-6 if dtype == "f32" else 1e-12
    lps = grid_pair(dtype, 203, 157, L, kind, seed=23, **tuning)
    run_lockstep(lps, "anisotropic", iters=3, lb_rtol=rt)
    for H, W in ((1, 40), (40, 1), (3, 3), (9, 2), (17, 70), (2, 5), (5, 9), (33, 8)):
        lps = grid_pair(dtype, H, W, L, kind, seed=H * 100 + W, **tuning)
        run_lockstep(lps, "anisotropic", iters=2, lb_rtol=rt)


@pytest.mark.parametrize("kind,L,dtype,tuning", [("trunclin", 16, "f32", dict(wavefront_split=1)), ("trunclin", 16, "f64", dict(wavefront_split=1)),
                                                 ("trunclin", 16, "f32", dict(wavefront_split=1, wavefront_max_ctas=2)),
                                                 ("potts", 16, "f32", dict(wavefront_split=1)), ("potts", 32, "f32", dict(wavefront_split=1)),
                                                 ("potts", 16, "f64", dict(wavefront_split=1, wavefront_max_ctas=3)),
                                                 ("potts", 32, "f32", dict(wavefront_split=1, wavefront_max_ctas=1)),
                                                 ("trunclin", 16, "f32", dict(wavefront_split=1, wavefront_cluster=1)),
                                                 ("trunclin", 16, "f32", dict(wavefront_split=1, wavefront_cluster=2)),
                                                 ("potts", 16, "f64", dict(wavefront_split=1, wavefront_cluster=4)),
                                                 ("potts", 32, "f32", dict(wavefront_split=1, wavefront_cluster=4, wavefront_max_ctas=8)),
                                                 ("trunclin", 16, "f32", dict(wavefront_split=1, wavefront_cluster=8, wavefront_max_ctas=8))])
def test_wavefront_split_variant(kind, L, dtype, tuning):
    """one row per warp, the two receive edges split over the half-warps (kernels_grid_split.cuh): anisotropic passes and
    rounding; rings wrapping many times, several bands per CTA in sequence, ragged last band / tile, single rows / columns,
    negative Potts couplings (two smallest entries)"""
    rt = 1e-6 if dtype == "f32" else 1e-12
    lps = grid_pair(dtype, 203, 157, L, kind, seed=23, **tuning)
    run_lockstep(lps, "anisotropic", iters=3, lb_rtol=rt)
    for H, W in ((1, 40), (40, 1), (3, 3), (9, 2), (17, 70), (2, 5), (5, 9), (33, 8), (4, 4), (11, 17)):
        lps = grid_pair(dtype, H, W, L, kind, seed=H * 100 + W, **tuning)
        run_lockstep(lps, "anisotropic", iters=2, lb_rtol=rt)


def test_wavefront_split_variant_fifo_and_synthetic():
    lps = grid_pair("f64", 21, 23, 16, "trunclin", seed=5, edge_order="fifo", wavefront_split=1)
    run_lockstep(lps, "anisotropic", iters=3)
    for kind, lo, hi, L in (("trunclin", 0.25, 1.0, 16), ("potts", 0.1, 1.0, 32), ("potts", 0.1, 1.0, 16)):
        lps = [oracle_lp("f32").grid_create(67, 81, L, kind), cuda_lp("f32").grid_create(67, 81, L, kind)]
        lps[1].set_tuning(wavefront_split=1)
        for lp in lps:
            lp.grid_fill_synthetic(0x5EED, lo, hi)
        run_lockstep(lps, "anisotropic", iters=3, lb_rtol=1e-6)


def test_wavefront_rows_variant_fifo_and_synthetic():
    lps = grid_pair("f64", 21, 23, 16, "trunclin", seed=5, edge_order="fifo", wavefront_rows=1)
    run_lockstep(lps, "anisotropic", iters=3)
    for kind, lo, hi, L in (("potts", 0.1, 1.0, 8), ("trunclin", 0.25, 1.0, 16), ("potts", 0.1, 1.0, 32)):
        lps = [oracle_lp("f32").grid_create(67, 81, L, kind), cuda_lp("f32").grid_create(67, 81, L, kind)]
        lps[1].set_tuning(wavefront_rows=1)
        for lp in lps:
            lp.grid_fill_synthetic(0x5EED, lo, hi)
        run_lockstep(lps, "anisotropic", iters=3, lb_rtol=1e-6)


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("kind,L,H,W", [("potts", 32, 37, 29), ("potts", 16, 64, 33), ("trunclin", 16, 21, 70), ("potts", 32, 8, 8), ("potts", 16, 3, 5)])
def test_wavefront_fused_bound_equals_bound_kernel(dtype, kind, L, H, W):
    """split variant: the backward pass accumulates LP::LowerBound (include/LP_MP.h:1507-1518) on the fly; the result must equal
    the separate sweep over the records (same terms, other summation order) and the oracle's bound -- plain passes and rounding
    passes, LIFO and FIFO message order, negative Potts weights (grid_pair draws lambda from [-0.3, 1))"""
    for edge_order in ("lifo", "fifo"):
        if kind == "potts":
            orc, fused = grid_pair(dtype, H, W, L, kind, seed=H + W, edge_order=edge_order, wavefront_fused_bound=1)
            _, plain = grid_pair(dtype, H, W, L, kind, seed=H + W, edge_order=edge_order, wavefront_fused_bound=0)
        else:   # truncated linear: the split variant needs a short label window (slope 0.25, truncation 1: reach 3)
            orc, fused, plain = (make(dtype).grid_create(H, W, L, kind) for make in (oracle_lp, cuda_lp, cuda_lp))
            fused.set_tuning(wavefront_fused_bound=1); plain.set_tuning(wavefront_fused_bound=0)
            for lp in (orc, fused, plain):
                lp.grid_fill_synthetic(H + W, 0.25, 1.0)
                lp.set_edge_order(edge_order)
        for lp in (orc, fused, plain):
            lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
        for it in range(3):
            for lp in (orc, fused, plain):
                lp.ComputePass(it)
            launches = fused.stats()["kernel_launches"]
            lb = fused.LowerBound()
            if not (dtype == "f64" and L == 32):   # (f64 rows of 32 labels do not fit the split kernel's shared-memory rings: sweep)
                assert fused.stats()["kernel_launches"] - launches == 1, "the fused bound is one final-sum launch, no sweep"
            assert np.isclose(lb, plain.LowerBound(), rtol=1e-13, atol=1e-9)
            assert np.isclose(lb, orc.LowerBound(), rtol=LB_RTOL[dtype], atol=1e-9)
            assert same_state(fused, plain)
        for lp in (orc, fused, plain):
            lp.ComputeBackwardPassAndPrimal(3)
        assert np.isclose(fused.LowerBound(), plain.LowerBound(), rtol=1e-13, atol=1e-9)
        assert np.isclose(fused.LowerBound(), orc.LowerBound(), rtol=LB_RTOL[dtype], atol=1e-9)
        assert np.array_equal(fused.labeling(), orc.labeling())
        # a forward pass invalidates the rows of the last backward pass: the sweep runs again
        fused.ComputeForwardPass(); plain.ComputeForwardPass()
        assert np.isclose(fused.LowerBound(), plain.LowerBound(), rtol=1e-13, atol=1e-9)


def test_wavefront_fused_bound_on_stripes():
    """stripes: the rows' terms come from the pass, the edges to the stripe below from the neighbour's vectors when the bound is asked for"""
    for dtype, kind, L, lo, hi, H, cut, W in (("f64", "potts", 32, 0.1, 1.0, 41, 17, 39), ("f32", "trunclin", 16, 0.25, 1.0, 70, 30, 45)):
        full, (top, bot) = _stripe_pair(dtype, H, cut, W, L, kind, lo, hi, tuning=dict(wavefront_fused_bound=1))
        top.grid_set_peer_boundary(1, bot.grid_boundary_ptrs(0)[0])
        bot.grid_set_peer_boundary(0, top.grid_boundary_ptrs(1)[0])
        for it in range(3):
            full.ComputePass(it)
            top.ComputeForwardPass(); bot.ComputeForwardPass()
            bot.ComputeBackwardPass(); top.ComputeBackwardPass()
            top.synchronize(); bot.synchronize()
            assert np.isclose(_striped_bound((top, bot)), full.LowerBound(), rtol=LB_RTOL[dtype])


def test_wavefront_fifo_order():
    lps = grid_pair("f64", 20, 23, 8, "potts", seed=5, edge_order="fifo")
    run_lockstep(lps, "damped_uniform", iters=3)
    lps = grid_pair("f64", 20, 23, 8, "potts", seed=5, edge_order="fifo")
    run_lockstep(lps, "anisotropic", iters=3)


def test_wavefront_equals_level_schedule_on_device():
    """two device schedules of the same DAG must agree bit for bit (order invariance)"""
    for dtype in ("f32", "f64"):
        a = grid_pair(dtype, 130, 97, 16, "trunclin", seed=8, schedule="wavefront")[1]
        b = grid_pair(dtype, 130, 97, 16, "trunclin", seed=8, schedule="levels")[1]
        run_lockstep([a, b], "anisotropic", iters=3, lb_rtol=1e-13)


def test_synthetic_generator_matches_oracle():
    for kind, lo, hi, L in (("potts", 0.1, 1.0, 8), ("trunclin", 0.25, 1.0, 16)):
        lps = [oracle_lp("f32").grid_create(19, 21, L, kind), cuda_lp("f32").grid_create(19, 21, L, kind)]
        for lp in lps:
            lp.grid_fill_synthetic(0x5EED, lo, hi)
        run_lockstep(lps, "anisotropic", iters=3, lb_rtol=1e-6)


# ---- grids with dense tables: row-dataflow kernel (BASELINE config 5 shape) --------------------------------------------
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("L", [16, 32, 64])
def test_dense_grid_bit_parity(dtype, L):
    lps = grid_pair(dtype, 23, 19, L, "dense", seed=L)
    run_lockstep(lps, "anisotropic", iters=3, lb_rtol=LB_RTOL[dtype])


@pytest.mark.parametrize("H,W", [(1, 1), (1, 7), (7, 1), (2, 2), (3, 40), (40, 3)])
def test_dense_grid_ragged_shapes(H, W):
    lps = grid_pair("f64", H, W, 16, "dense", seed=H * 100 + W)
    run_lockstep(lps, "anisotropic", iters=3)


@pytest.mark.parametrize("tuning", [dict(wavefront_max_ctas=1), dict(wavefront_max_ctas=5), dict(wavefront_ctas_per_sm=1)])
def test_dense_grid_launch_shapes_do_not_change_results(tuning):
    lps = grid_pair("f32", 31, 45, 64, "dense", seed=9, **tuning)
    run_lockstep(lps, "anisotropic", iters=2, lb_rtol=1e-6)


def test_dense_grid_fifo_and_uniform_modes():
    """FIFO message order; the uniform modes of the rounding iterations run on the level schedule of the same state"""
    lps = grid_pair("f64", 12, 14, 32, "dense", seed=6, edge_order="fifo")
    run_lockstep(lps, "anisotropic", iters=3)
    lps = grid_pair("f64", 12, 14, 16, "dense", seed=7)
    run_lockstep(lps, "anisotropic", iters=2)
    for lp in lps:
        lp.set_reparametrization("damped_uniform")
        lp.ComputePass(2)
    assert np.array_equal(lps[0].unaries(), lps[1].unaries())
    for lp in lps:
        lp.set_reparametrization("anisotropic")
        lp.ComputePass(3)
    assert np.array_equal(lps[0].unaries(), lps[1].unaries())
    a0, b0 = lps[0].messages(); a1, b1 = lps[1].messages()
    assert np.array_equal(a0, a1) and np.array_equal(b0, b1)


def test_dense_grid_fused_bound_equals_bound_kernel():
    """the backward pass accumulates LP::LowerBound on the fly; the stand-alone bound kernel must give the same value"""
    for dtype in ("f32", "f64"):
        lp = cuda_lp(dtype).grid_create(21, 33, 32, "dense")
        lp.grid_fill_synthetic(7, -0.2, 1.0)
        lp.add_to_constant(3.5)
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
        for it in range(3):
            lp.ComputePass(it)
            fused = lp.LowerBound()
            lp.set_unaries(lp.unaries())          # any state write invalidates the fused rows
            assert np.isclose(lp.LowerBound(), fused, rtol=1e-13, atol=0)
        lp.ComputeForwardPass()                    # after a forward pass the bound comes from the bound kernel
        orc = oracle_lp(dtype).grid_create(21, 33, 32, "dense")
        orc.grid_fill_synthetic(7, -0.2, 1.0)
        orc.add_to_constant(3.5)
        orc.finalize(); orc.Begin(); orc.set_reparametrization("anisotropic")
        for it in range(3):
            orc.ComputePass(it)
        orc.ComputeForwardPass()
        assert np.isclose(lp.LowerBound(), orc.LowerBound(), rtol=LB_RTOL[dtype])


def test_dense_grid_equals_level_schedule_and_synthetic_generator():
    for dtype in ("f32", "f64"):
        lps = [oracle_lp(dtype).grid_create(17, 22, 64, "dense"), cuda_lp(dtype).grid_create(17, 22, 64, "dense"),
               cuda_lp(dtype).grid_create(17, 22, 64, "dense")]
        lps[2].set_schedule("levels")
        for lp in lps:
            lp.grid_fill_synthetic(0x5EED, 0.0, 1.0)
        run_lockstep(lps, "anisotropic", iters=3, lb_rtol=LB_RTOL[dtype])


# ---- row stripes (multi-GPU path) exercised with two handles on ONE device ------------------------------------------------
def _stripe_pair(dtype, H, cut, W, L, kind, lo, hi, seed=0x5EED, tuning=None):
    """full LP + the two stripes [0, cut), [cut, H) of the same synthetic problem"""
    full = cuda_lp(dtype).grid_create(H, W, L, kind)
    full.grid_fill_synthetic(seed, lo, hi)
    stripes = []
    for row0, rows in ((0, cut), (cut, H - cut)):
        lp = cuda_lp(dtype).grid_create(rows, W, L, kind)
        if tuning:
            lp.set_tuning(**tuning)
        lp.grid_set_stripe(row0, H)
        lp.grid_fill_synthetic(seed, lo, hi)
        stripes.append(lp)
    for lp in [full] + stripes:
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
    return full, stripes


def _striped_bound(stripes):
    import torch
    top, bot = stripes
    W, L = top.grid_shape[1], top.L
    buf = torch.empty(W * L, dtype=torch.float32 if top.dtype == 0 else torch.float64, device="cuda")
    bot.grid_get_boundary_vectors(0, buf.data_ptr())
    bot.synchronize()
    return top.grid_lower_bound_partial(buf.data_ptr()) + bot.grid_lower_bound_partial(None)


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("kind,L,lo,hi,H,cut,W", [("trunclin", 16, 0.25, 1.0, 70, 30, 45), ("potts", 32, 0.1, 1.0, 41, 17, 39),
                                                  ("potts", 8, 0.1, 1.0, 33, 16, 20), ("dense", 16, 0.0, 1.0, 29, 13, 21), ("dense", 64, 0.0, 1.0, 9, 4, 12)])
def test_stripes_pipelined_are_bit_identical_to_the_unstriped_lp(dtype, kind, L, lo, hi, H, cut, W):
    """exact coupling: the last band of the upper stripe stores its outgoing vectors straight into the lower stripe's
    in-buffer while that stripe's pass kernel polls them (on multi-GPU boxes the pointer is a cudaIpc mapping)"""
    full, (top, bot) = _stripe_pair(dtype, H, cut, W, L, kind, lo, hi)
    top.grid_set_peer_boundary(1, bot.grid_boundary_ptrs(0)[0])   # forward: top's last row -> bot.in(top)
    bot.grid_set_peer_boundary(0, top.grid_boundary_ptrs(1)[0])   # backward: bot's first row -> top.in(bottom)
    for it in range(3):
        full.ComputePass(it)
        top.ComputeForwardPass(); bot.ComputeForwardPass()        # both kernels in flight, coupled through the buffers
        bot.ComputeBackwardPass(); top.ComputeBackwardPass()
        top.synchronize(); bot.synchronize()
        assert np.array_equal(np.concatenate([top.unaries(), bot.unaries()]), full.unaries()), f"theta differs after pass {it}"
        assert np.isclose(_striped_bound((top, bot)), full.LowerBound(), rtol=LB_RTOL[dtype])
    fa, fb = full.messages()
    ta, tb = top.messages()
    # edges of the top stripe come first in insertion order, except that its last row has no vertical edges
    n_top_inner = len(ta)
    rows_full = []
    e = 0
    for r in range(H):
        for c in range(W):
            if c + 1 < W:
                rows_full.append((r, e)); e += 1
            if r + 1 < H:
                rows_full.append((r if r + 1 != cut else -1, e)); e += 1
    keep = [k for (r, k) in rows_full if 0 <= r < cut]
    assert len(keep) == n_top_inner
    assert np.array_equal(fa[keep], ta) and np.array_equal(fb[keep], tb)


@pytest.mark.parametrize("mode", ["uniform", "damped_uniform"])
@pytest.mark.parametrize("dtype,kind,L,lo,hi,H,cut,W", [("f64", "trunclin", 16, 0.25, 1.0, 70, 30, 45), ("f32", "potts", 32, 0.1, 1.0, 41, 17, 39),
                                                        ("f64", "potts", 8, 0.1, 1.0, 33, 16, 20), ("f32", "trunclin", 16, 0.25, 1.0, 19, 8, 64),
                                                        ("f32", "potts", 16, -0.3, 1.0, 12, 11, 24),
                                                        # stripes of whole bands of 8 rows, W % 8 == 0: the split kernel's uniform instances
                                                        ("f32", "potts", 32, 0.1, 1.0, 48, 16, 40), ("f32", "trunclin", 16, 0.25, 1.0, 40, 24, 64),
                                                        ("f64", "trunclin", 16, 0.25, 1.0, 32, 16, 32), ("f32", "potts", 16, 0.1, 1.0, 16, 8, 16)])
def test_stripes_uniform_modes_pipelined_are_bit_identical_to_the_unstriped_lp(mode, dtype, kind, L, lo, hi, H, cut, W):
    """uniform / damped-uniform passes on stripes (the rounding iterations of the default control law,
    include/visitors/standard_visitor.hxx:44): the last virtual row of a stripe also receives from the other stripe's first
    virtual row, whose vectors -- as they were before the pass -- are pushed into the OTHER in-buffer before the pass."""
    full, (top, bot) = _stripe_pair(dtype, H, cut, W, L, kind, lo, hi)
    bot_in_top, top_in_bot = bot.grid_boundary_ptrs(0)[0], top.grid_boundary_ptrs(1)[0]
    top.grid_set_peer_boundary(1, bot_in_top)
    bot.grid_set_peer_boundary(0, top_in_bot)
    for lp in (full, top, bot):
        lp.set_reparametrization(mode)

    def both(fwd, primal, it):
        name = "Compute" + ("Forward" if fwd else "Backward") + "Pass" + ("AndPrimal" if primal else "")
        args = (it,) if primal else ()
        getattr(full, name)(*args)
        if fwd:
            bot.grid_push_boundary(0, top_in_bot, False)      # bU of the lower stripe's first row, state before the pass
            getattr(top, name)(*args); getattr(bot, name)(*args)
        else:
            top.grid_push_boundary(1, bot_in_top, True)       # aD of the upper stripe's last row (backward pass: slots right to left)
            getattr(bot, name)(*args); getattr(top, name)(*args)
        top.synchronize(); bot.synchronize()

    for it in range(3):
        for fwd in (True, False):
            both(fwd, False, it)
            assert np.array_equal(np.concatenate([top.unaries(), bot.unaries()]), full.unaries()), f"theta differs after pass {it} fwd={fwd}"
        assert np.isclose(_striped_bound((top, bot)), full.LowerBound(), rtol=LB_RTOL[dtype])
    if dtype == "f32" and W % 8 == 0 and W >= 16 and cut % 8 == 0 and (H - cut) % 8 == 0 and (kind == "potts" or L == 16) and L in (16, 32):
        # the split instances ran: a forward pass of the upper stripe is the gather of the lower stripe's vectors + the pass kernel
        n0 = top.stats()["kernel_launches"]
        both(True, False, 3)
        assert top.stats()["kernel_launches"] - n0 == 2
        assert np.array_equal(np.concatenate([top.unaries(), bot.unaries()]), full.unaries())
    # anisotropic passes in between (the default law alternates), then a rounding iteration in the uniform mode
    for lp in (full, top, bot):
        lp.set_reparametrization("anisotropic")
    full.ComputePass(3)
    top.ComputeForwardPass(); bot.ComputeForwardPass(); bot.ComputeBackwardPass(); top.ComputeBackwardPass()
    top.synchronize(); bot.synchronize()
    for lp in (full, top, bot):
        lp.set_reparametrization(mode)
    for fwd in (True, False):
        both(fwd, True, 4)
        assert np.array_equal(np.concatenate([top.labeling(), bot.labeling()]), full.labeling()), "labeling differs"
        assert np.array_equal(np.concatenate([top.unaries(), bot.unaries()]), full.unaries())
    assert np.isclose(_striped_bound((top, bot)), full.LowerBound(), rtol=LB_RTOL[dtype])


def test_stripes_uniform_modes_per_pass_halo():
    """block-Jacobi coupling in the damped-uniform mode: both boundaries of every stripe are pushed before a pass.  Both endpoints
    of a boundary edge receive from it on the state of the previous pass, so the bound is that of a valid reparametrisation (below
    every primal cost) but not necessarily monotone; it stays close to the exact run's"""
    H, cut, W, L = 96, 48, 64, 16
    full, (top, bot) = _stripe_pair("f64", H, cut, W, L, "trunclin", 0.25, 1.0)
    bot_in_top, top_in_bot = bot.grid_boundary_ptrs(0)[0], top.grid_boundary_ptrs(1)[0]
    for lp in (full, top, bot):
        lp.set_reparametrization("damped_uniform")
    prev = _striped_bound((top, bot))
    for it in range(10):
        full.ComputePass(it)
        for fwd in (True, False):
            top.grid_push_boundary(1, bot_in_top, not fwd); bot.grid_push_boundary(0, top_in_bot, not fwd)
            if fwd:
                top.ComputeForwardPass(); bot.ComputeForwardPass()
            else:
                bot.ComputeBackwardPass(); top.ComputeBackwardPass()
        top.synchronize(); bot.synchronize()
        prev = _striped_bound((top, bot))
        assert abs(prev - full.LowerBound()) <= 2e-2 * abs(prev), (it, prev, full.LowerBound())
    full.ComputeForwardPassAndPrimal(10)
    top.grid_push_boundary(1, bot_in_top, False); bot.grid_push_boundary(0, top_in_bot, False)
    top.ComputeForwardPassAndPrimal(10); bot.ComputeForwardPassAndPrimal(10)
    top.synchronize(); bot.synchronize()
    assert prev <= full.EvaluatePrimal() + 1e-9
    assert np.mean(np.concatenate([top.labeling(), bot.labeling()]) == full.labeling()) > 0.9


def test_stripes_per_pass_halo_is_monotone_and_close():
    """block-Jacobi coupling: each stripe sweeps with the neighbour's vectors of the previous pass.  Valid
    reparametrisation (monotone bound, below the exact one or equal) with a slightly weaker fixed point."""
    H, cut, W, L = 96, 48, 64, 16
    full, (top, bot) = _stripe_pair("f64", H, cut, W, L, "trunclin", 0.25, 1.0)
    prev = _striped_bound((top, bot))
    assert np.isclose(prev, full.LowerBound(), rtol=1e-12)
    for it in range(20):
        full.ComputePass(it)
        top.grid_push_boundary(1, bot.grid_boundary_ptrs(0)[0], False); top.synchronize()
        top.ComputeForwardPass(); bot.ComputeForwardPass()
        top.synchronize(); bot.synchronize()
        bot.grid_push_boundary(0, top.grid_boundary_ptrs(1)[0], True); bot.synchronize()
        bot.ComputeBackwardPass(); top.ComputeBackwardPass()
        top.synchronize(); bot.synchronize()
        lb = _striped_bound((top, bot))
        assert lb >= prev - 1e-9 * abs(prev)
        prev = lb
    exact = full.LowerBound()
    assert prev <= exact + 1e-9 * abs(exact)
    assert abs(prev - exact) <= 5e-3 * abs(exact), (prev, exact)


@pytest.mark.parametrize("dtype,kind,L,lo,hi,H,cut,W", [("f64", "trunclin", 16, 0.25, 1.0, 37, 16, 29), ("f32", "potts", 32, 0.1, 1.0, 48, 16, 40),
                                                        ("f32", "potts", 8, -0.3, 1.0, 21, 9, 17), ("f32", "trunclin", 16, 0.25, 1.0, 40, 24, 64),
                                                        ("f32", "dense", 16, 0.0, 1.0, 24, 11, 19), ("f64", "dense", 32, 0.0, 1.0, 10, 6, 13)])
def test_stripes_push_after_the_neighbours_pass_is_exact(dtype, kind, L, lo, hi, H, cut, W):
    """the push of the per-pass halo coupling delivers what the pipelined coupling delivers (slot q = q-th column in processing
    order: right to left for a backward pass): pushed AFTER the neighbour's pass has finished, the stripes reproduce the unstriped
    LP bit for bit.  (Pushed before it -- the coupling proper -- the same vectors arrive one pass late.)"""
    full, (top, bot) = _stripe_pair(dtype, H, cut, W, L, kind, lo, hi)
    bot_in_top, top_in_bot = bot.grid_boundary_ptrs(0)[0], top.grid_boundary_ptrs(1)[0]
    for it in range(3):
        full.ComputeForwardPass()
        top.ComputeForwardPass(); top.grid_push_boundary(1, bot_in_top, False); top.synchronize()
        bot.ComputeForwardPass(); bot.synchronize()
        assert np.array_equal(np.concatenate([top.unaries(), bot.unaries()]), full.unaries()), f"theta differs after forward pass {it}"
        full.ComputeBackwardPass()
        bot.ComputeBackwardPass(); bot.grid_push_boundary(0, top_in_bot, True); bot.synchronize()
        top.ComputeBackwardPass(); top.synchronize()
        assert np.array_equal(np.concatenate([top.unaries(), bot.unaries()]), full.unaries()), f"theta differs after backward pass {it}"
        assert np.isclose(_striped_bound((top, bot)), full.LowerBound(), rtol=LB_RTOL[dtype])


# ---- solver loop: reference control law driving the device ---------------------------------------------------------
def test_kept_labeling_is_a_device_side_snapshot():
    """Solver::RegisterPrimal (include/solver.hxx:321-340) keeps the best labeling: lpmp_keep_labeling copies it aside on the device,
    later rounding passes do not touch the copy"""
    chain = fill(cuda_lp("f64"), "dense", np.random.default_rng(1).random((300, 4)), np.arange(299, dtype=np.uint64), np.arange(1, 300, dtype=np.uint64), 4, 4)
    for lp in (grid_pair("f32", 40, 64, 16, "trunclin", seed=3)[1], chain):
        if not lp.finalized:
            lp.finalize()
        lp.Begin(); lp.set_reparametrization("anisotropic")
        with pytest.raises(Exception):
            lp.keep_labeling()            # no rounding pass yet
        lp.ComputeForwardPassAndPrimal(0)
        first = lp.labeling()
        lp.keep_labeling()
        lp.set_reparametrization("damped_uniform")
        lp.ComputeBackwardPassAndPrimal(0)
        lp.ComputePass(1)
        lp.ComputeForwardPassAndPrimal(2)
        assert lp is chain or not np.array_equal(lp.labeling(), first)
        assert np.array_equal(lp.kept_labeling(), first)


def test_frustrated_cycle_goldens_on_device():
    """test/graphical_model.cpp:98-136 (lower bounds are goldens; see tests/test_oracle_goldens.py for the primal note)"""
    from lp_mp_b200 import MpRoundingSolver, StandardVisitor
    NEG, POS = np.array([[1.0, 0.0], [0.0, 1.0]]), np.array([[0.0, 1.0], [1.0, 0.0]])
    for n, edges, ub in ((5, [(0, 1, NEG), (1, 2, POS), (2, 3, POS), (3, 4, POS)], 1.0),
                         (4, [(0, 1, NEG), (1, 2, POS), (2, 3, POS), (0, 3, POS)], 3.0)):
        lp = cuda_lp("f64")
        for _ in range(n):
            lp.AddUnaryFactor(np.zeros(2))
        for i, j, t in edges:
            lp.AddPairwiseFactor(i, j, t)
        s = MpRoundingSolver(lp, StandardVisitor(maxIter=100))
        s.Solve()
        assert abs(s.lower_bound()) <= 1e-8 and s.primal_cost() <= ub + 1e-8


def test_solve_trajectory_matches_reference_solver():
    """MpRoundingSolver + StandardVisitor defaults (rounding in damped_uniform every 5th iteration) on the device vs the
    reference's own Solve() (oracle/_ref) if it travelled to this box, else vs the oracle under the same Python loop."""
    from checkers import RefLP, ref_library
    from lp_mp_b200 import MpRoundingSolver, StandardVisitor
    rng = np.random.default_rng(8)
    H, W, L = 8, 8, 4
    ei, ej = grid_edges(H, W)
    theta, tables = rng.random((H * W, L)), rng.random((len(ei), L, L))
    lp = cuda_lp("f64")
    lp.add_unaries(theta); lp.add_pairwise_dense(ei, ej, tables)
    s = MpRoundingSolver(lp, StandardVisitor(maxIter=13))
    s.Solve()
    lib = ref_library()
    if lib is not None:
        ref = RefLP(lib, ["-v", "0", "--maxIter", "13"])
        for t in theta:
            ref.AddUnaryFactor(t)
        for e in range(len(ei)):
            ref.AddPairwiseFactor(int(ei[e]), int(ej[e]), tables[e])
        r = ref.solve()
        want_lb, want_ub, flb, fub = r["lb"], r["ub"], r["final_lb"], r["final_ub"]
    else:
        o = oracle_lp("f64")
        o.add_unaries(theta); o.add_pairwise_dense(ei, ej, tables)
        so = MpRoundingSolver(o, StandardVisitor(maxIter=13))
        so.Solve()
        want_lb, want_ub = [h[0] for h in so.history], [h[1] for h in so.history]
        flb, fub = so.lower_bound(), so.primal_cost()
    assert len(s.history) == len(want_lb) == 13
    for k, (lb, ub, c) in enumerate(s.history):
        assert np.isclose(lb, want_lb[k], rtol=1e-12)
        assert ub == want_ub[k] or np.isclose(ub, want_ub[k], rtol=1e-12)
    assert np.isclose(s.lower_bound(), flb, rtol=1e-12) and np.isclose(s.primal_cost(), fub, rtol=1e-12)
    assert [k for k, h in enumerate(s.history) if h[2].computePrimal] == [1, 6, 11, 12]


# ---- (3) properties at BASELINE.json's full sizes -------------------------------------------------------------------
@pytest.mark.parametrize("H,W,L,kind,lo,hi", [(1024, 1024, 16, "trunclin", 0.25, 1.0), (4096, 4096, 32, "potts", 0.1, 1.0)])
def test_full_size_properties(H, W, L, kind, lo, hi):
    """configs 2 and 3: dual ascent is monotone, bound <= primal, reruns are bit-identical (deterministic schedule),
    and the bound equals the oracle's on a crop-independent checksum: the first pass' bound of the top-left 64x64 crop
    cannot be compared (the grid couples), so we compare the full-size run with a second run under another launch shape."""
    lp = cuda_lp("f32").grid_create(H, W, L, kind)
    lp.grid_fill_synthetic(0x5EED, lo, hi)
    lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
    prev = lp.LowerBound()
    lbs = []
    for it in range(3):
        lp.ComputePass(it)
        lb = lp.LowerBound()
        assert lb >= prev - 1e-6 * abs(prev)
        prev = lb
        lbs.append(lb)
    lp.ComputeForwardPassAndPrimal(3)
    primal = lp.EvaluatePrimal()
    assert np.isfinite(primal) and primal >= prev - 1e-6 * abs(prev)
    labels = lp.labeling()
    assert labels.max() < L
    other = cuda_lp("f32").grid_create(H, W, L, kind)
    other.grid_fill_synthetic(0x5EED, lo, hi)
    other.set_tuning(wavefront_max_ctas=97, wavefront_ctas_per_sm=1)
    other.finalize(); other.Begin(); other.set_reparametrization("anisotropic")
    for it in range(3):
        other.ComputePass(it)
        assert other.LowerBound() == lbs[it]
    other.ComputeForwardPassAndPrimal(3)
    assert np.array_equal(other.labeling(), labels) and other.EvaluatePrimal() == primal


@pytest.mark.parametrize("H,W,L,kind,lo,hi", [(1024, 1024, 16, "trunclin", 0.25, 1.0), (1024, 4096, 32, "potts", 0.1, 1.0)])
def test_full_size_kernel_variants_are_bit_identical(H, W, L, kind, lo, hi):
    """config 2 at full size and config 3 at full width: the one-warp kernel, the lean variant without clusters (every band
    boundary through the L2 hand-off slots) and with clusters of 4 / 8 (distributed shared memory) leave bit-identical
    theta / messages / labels after two iterations and a rounding pass: the schedule is exact whatever moves the data."""
    ref = None
    for tuning in (dict(wavefront_lean=0), dict(wavefront_cluster=1), dict(wavefront_cluster=4, wavefront_tile_cols=4), dict()):
        lp = cuda_lp("f32").grid_create(H, W, L, kind)
        lp.set_tuning(**tuning)
        lp.grid_fill_synthetic(0x5EED, lo, hi)
        lp.finalize(); lp.Begin(); lp.set_reparametrization("anisotropic")
        for it in range(2):
            lp.ComputePass(it)
        lb = lp.LowerBound()
        lp.ComputeForwardPassAndPrimal(2)
        state = (lp.unaries(), *lp.messages(), lp.labeling(), lb)
        del lp
        if ref is None:
            ref = state
        else:
            for x, y in zip(ref[:4], state[:4]):
                assert np.array_equal(x, y), tuning
            assert state[4] == ref[4]


def test_mid_size_grid_against_oracle():
    """256x256 crops of configs 2 / 3 against the oracle (seconds on the CPU), f32 bit-exact"""
    for kind, L, lo, hi in (("trunclin", 16, 0.25, 1.0), ("potts", 32, 0.1, 1.0)):
        lps = [oracle_lp("f32").grid_create(256, 256, L, kind), cuda_lp("f32").grid_create(256, 256, L, kind)]
        for lp in lps:
            lp.grid_fill_synthetic(7, lo, hi)
        run_lockstep(lps, "anisotropic", iters=2, lb_rtol=1e-6)


# ---- the oracle at BASELINE.json's sizes (it does a C2 iteration in ~0.6 s and a 1 M-node C4 iteration in ~1 s on one core) ---------
@pytest.mark.parametrize("H,W,L,kind,lo,hi", [(1024, 1024, 16, "trunclin", 0.25, 1.0),      # config 2, full size
                                              (512, 4096, 32, "potts", 0.1, 1.0),            # config 3: full width, 1/8 of the rows (one stripe at 8 GPUs)
                                              (12, 2048, 64, "dense", 0.0, 1.0)])            # config 5: full width, 12 rows (oracle tables are f64: 1.5 GB)
def test_baseline_sizes_bit_exact_against_oracle(H, W, L, kind, lo, hi):
    """f32 (the bench's storage type): one iteration, the bound, then a forward and a backward rounding pass -- theta, a, b and the
    labeling bit-identical to the oracle after every step, bound / primal cost to summation order"""
    lps = [oracle_lp("f32").grid_create(H, W, L, kind), cuda_lp("f32").grid_create(H, W, L, kind)]
    for lp in lps:
        lp.grid_fill_synthetic(0x5EED, lo, hi)
    run_lockstep(lps, "anisotropic", iters=1, lb_rtol=1e-6)


def test_config4_full_size_bit_exact_against_oracle():
    """BASELINE config 4 at full size (random 3-regular graph, 1 M nodes, 8 labels, dense tables, f32): two iterations + rounding on
    the colour ranks AND in the natural (reference) order, bit-exact against the oracle run on the same ranks"""
    from bench import random_regular_graph
    n, L = 1_000_000, 8
    ei, ej = random_regular_graph(n, 3, 0x5EED)
    rng = np.random.default_rng(0x5EED)
    theta, tables = rng.random((n, L)), rng.random((len(ei), L, L))
    for colouring in (True, False):
        lps = []
        for make in (oracle_lp, cuda_lp):
            lp = make("f32")
            lp.add_unaries(theta); lp.add_pairwise_dense(ei, ej, tables)
            lps.append(lp)
        if colouring:
            colour, nc = lps[1].greedy_colouring(ei, ej)
            assert nc <= 4
            for lp in lps:
                lp.set_node_ranks(colour.astype(np.uint64) * n + np.arange(n, dtype=np.uint64))
        run_lockstep(lps, "anisotropic", iters=2, lb_rtol=1e-6)
        levels = lps[1].stats()["num_levels"]
        assert levels <= (4 if colouring else 16)   # natural order: the index-oriented DAG of a random cubic graph is 12 levels deep
        del lps


def test_error_behaviour_on_device():
    from lp_mp_b200 import LpmpError
    lp = cuda_lp("f64")
    lp.add_unaries(np.zeros((3, 2)))
    with pytest.raises(LpmpError):
        lp.add_pairwise_potts([0], [7], [1.0])
    lp.add_pairwise_potts([0, 1], [1, 2], [1.0, 1.0])
    with pytest.raises(LpmpError):
        lp.ComputePass(0)
    lp.finalize(); lp.Begin()
    with pytest.raises(LpmpError, match="no reparametrization mode set"):
        lp.ComputePass(0)
    with pytest.raises(LpmpError):
        lp.add_unaries(np.zeros((1, 2)))
    assert lp.EvaluatePrimal() == np.inf
    g = cuda_lp("f32").grid_create(4, 4, 5, "potts")      # L=5: no wavefront instance
    g.set_schedule("wavefront")
    with pytest.raises(LpmpError):
        g.finalize()
