"""Size-independent properties of the message-passing sweep, checked on the CPU restatement (oracle/lpmp_oracle.cpp) over random
small graphs (hypothesis).  A pass only REPARAMETRISES the problem (include/factors_messages.hxx:688-760: what a message takes
from one factor it gives to the other), so

  * the energy of every labeling, sum_i theta_i(x_i) + sum_e [c_e(x_i, x_j) + a_e(x_i) + b_e(x_j)], never changes;
  * LP::LowerBound (include/LP_MP.h:1507-1518) stays below the energy of every labeling, and below the primal cost of a rounding
    pass (LP::EvaluatePrimal, :1520-1536);
  * with the anisotropic (SRMP) weights every node update is a block-coordinate ascent step: the bound never decreases.

The same identities hold for the device library by bit parity with this oracle (tests/test_gpu_*.py)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from checkers import oracle_lp

MODES = ["anisotropic", "anisotropic2", "uniform", "damped_uniform"]


def random_instance(seed, n, m, L):
    rng = np.random.default_rng(seed)
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    pick = rng.permutation(len(pairs))[:min(m, len(pairs))]
    ei = np.array([pairs[k][0] for k in pick], dtype=np.uint64)
    ej = np.array([pairs[k][1] for k in pick], dtype=np.uint64)
    theta = rng.normal(size=(n, L))
    tables = rng.normal(size=(len(ei), L, L))
    return rng, theta, ei, ej, tables


def energy(theta, ei, ej, tables, a, b, x):
    e = theta[np.arange(len(x)), x].sum()
    for k, (i, j) in enumerate(zip(ei, ej)):
        e += tables[k][x[i], x[j]] + a[k][x[i]] + b[k][x[j]]
    return e


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(2, 9), m=st.integers(1, 14), L=st.integers(2, 4),
       mode=st.sampled_from(MODES), send=st.sampled_from(["shared", "residual"]), fifo=st.booleans())
def test_passes_reparametrise_and_bound_from_below(seed, n, m, L, mode, send, fifo):
    rng, theta, ei, ej, tables = random_instance(seed, n, m, L)
    lp = oracle_lp("f64")
    if fifo:
        lp.set_edge_order("fifo")
    lp.add_unaries(theta); lp.add_pairwise_dense(ei, ej, tables)
    lp.finalize(); lp.Begin(); lp.set_send_mode(send); lp.set_reparametrization(mode)
    labelings = [rng.integers(0, L, size=n) for _ in range(6)]
    zero = np.zeros((len(ei), L))
    e0 = [energy(theta, ei, ej, tables, zero, zero, x) for x in labelings]
    prev = lp.LowerBound()
    for it in range(4):
        for fwd in (True, False):
            lp.ComputeForwardPass() if fwd else lp.ComputeBackwardPass()
            th = lp.unaries()
            a, b = lp.messages()
            for x, want in zip(labelings, e0):
                assert abs(energy(th, ei, ej, tables, a, b, x) - want) <= 1e-9 * (1.0 + abs(want))
            lb = lp.LowerBound()
            assert lb <= min(e0) + 1e-9 * (1.0 + abs(min(e0)))
            if mode == "anisotropic" and send == "shared":
                assert lb >= prev - 1e-9 * (1.0 + abs(prev)), (it, fwd, prev, lb)
            prev = lb
    lp.ComputeForwardPassAndPrimal(4)
    x = lp.labeling().astype(np.int64)
    cost = lp.EvaluatePrimal()
    # the primal cost is the energy of the rounded labeling in the ORIGINAL problem, and the bound stays below it
    assert abs(cost - energy(theta, ei, ej, tables, zero, zero, x)) <= 1e-9 * (1.0 + abs(cost))
    assert lp.LowerBound() <= cost + 1e-9 * (1.0 + abs(cost))
