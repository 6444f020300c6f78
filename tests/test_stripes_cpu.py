"""Host-side logic of the row-stripe partition (lp_mp_b200/stripes.py) on CPU: the partition plan and, with
world_size 2 and 3 over the gloo backend, the neighbour exchange that carries the halo vectors and the partial
lower bounds.  The kernels themselves are covered by the -m gpu suite (two stripes on one device, bit parity)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lp_mp_b200.stripes import BOTTOM, TOP, StripePlan, neighbour_exchange


def test_plan_partitions_rows_contiguously():
    for H, world in ((1024, 8), (10, 3), (7, 7), (4096, 5)):
        plan = StripePlan(H, world)
        nxt = 0
        for r in range(world):
            row0, n = plan.rows(r)
            assert row0 == nxt and n >= 1
            nxt += n
            assert plan.owner(row0) == r and plan.owner(row0 + n - 1) == r
        assert nxt == H
        sizes = [plan.rows(r)[1] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        StripePlan(3, 4)
    assert not StripePlan(8, 2).has_above(0) and StripePlan(8, 2).has_below(0) and not StripePlan(8, 2).has_below(1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = StripePlan(40, world)
        row0, rows = plan.rows(rank)
        W = 6
        mine_bottom = torch.full((W,), float(100 + rank))     # stands for aD of my last row
        mine_top = torch.full((W,), float(200 + rank))        # stands for bU of my first row
        in_top, in_bot = torch.full((W,), -1.0), torch.full((W,), -1.0)
        # forward pass halo: boundary BOTTOM travels down into the lower neighbour's in_top
        neighbour_exchange(dist, rank, world, BOTTOM, mine_bottom if plan.has_below(rank) else None, in_top if plan.has_above(rank) else None)
        # backward pass halo: boundary TOP travels up into the upper neighbour's in_bot
        neighbour_exchange(dist, rank, world, TOP, mine_top if plan.has_above(rank) else None, in_bot if plan.has_below(rank) else None)
        ok = True
        ok &= bool((in_top == (100 + rank - 1)).all()) if plan.has_above(rank) else bool((in_top == -1).all())
        ok &= bool((in_bot == (200 + rank + 1)).all()) if plan.has_below(rank) else bool((in_bot == -1).all())
        # lower bound: partial sums all-reduced in double
        t = torch.tensor([float(rows)], dtype=torch.float64)
        dist.all_reduce(t)
        ok &= t.item() == 40.0
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_exchange_over_gloo(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert list(out) == [1] * world


# ---- StripedLP driven end to end over gloo with a host stand-in for the per-rank LP ------------------------------------------
class HostStripe:
    """Stand-in for lp_mp_b200.LP on a stripe (host memory, numpy): enough behaviour to pin StripedLP's host logic -- which
    vectors travel where before which pass, what enters the bound / primal reduction, how the labeling is gathered."""
    dtype = 1   # f64

    def __init__(self, rank, rows, W, L, has_above, has_below):
        self.rank, self.grid_shape, self.L = rank, (rows, W), L
        self.in_buf = {0: np.full(W * L, -1.0) if has_above else None, 1: np.full(W * L, -1.0) if has_below else None}
        self.passes, self.seen, self.labelled, self.repam = 0, [], False, None
        self.finalized = False

    def _vec(self, which):      # my side of boundary `which` after `passes` passes; + 0.25 * column
        W, L = self.grid_shape[1], self.L
        return 1000.0 * self.rank + 100.0 * which + self.passes + 0.25 * np.repeat(np.arange(W, dtype=np.float64), L)

    def finalize(self): self.finalized = True
    def Begin(self): pass
    def End(self): self.ended = True
    def set_reparametrization(self, mode): self.repam = mode
    def synchronize(self): pass

    def grid_boundary_ptrs(self, which):
        return self.in_buf[which].ctypes.data, self.in_buf[which].nbytes

    def grid_get_boundary_vectors(self, which, dst):
        import ctypes
        v = self._vec(which)
        ctypes.memmove(dst, v.ctypes.data, v.nbytes)

    def grid_get_boundary_labels(self, which, dst):
        import ctypes
        v = np.full(self.grid_shape[1], 10 * self.rank + which, dtype=np.int32)
        ctypes.memmove(dst, v.ctypes.data, v.nbytes)

    def _pass(self, which_in, primal):
        buf = self.in_buf[which_in]
        self.seen.append(None if buf is None else float(buf[0]))
        other = self.in_buf[1 - which_in]
        self.seen_later = None if other is None else float(other[0])
        self.passes += 1
        self.labelled |= primal

    def ComputeForwardPass(self): self._pass(0, False)
    def ComputeBackwardPass(self): self._pass(1, False)
    def ComputeForwardPassAndPrimal(self, it): self._pass(0, True)
    def ComputeBackwardPassAndPrimal(self, it): self._pass(1, True)

    def grid_lower_bound_partial_device(self, below, out):
        import ctypes
        n = self.grid_shape[1] * self.L
        val = float(self.rank + 1) + (float(np.ctypeslib.as_array((ctypes.c_double * n).from_address(below)).sum()) / n if below else 0.0)
        ctypes.c_double.from_address(out).value = val

    def grid_evaluate_primal_partial_device(self, below, below_labels, out):
        import ctypes
        val = float("inf") if not self.labelled else 7.0 * (self.rank + 1) + (float(ctypes.c_int32.from_address(below_labels).value) if below_labels else 0.0)
        ctypes.c_double.from_address(out).value = val

    def CheckPrimalConsistency(self): return self.labelled
    def labeling(self): return np.full(self.grid_shape[0] * self.grid_shape[1], self.rank, dtype=np.uint32)
    def keep_labeling(self): self.kept = self.passes                     # device-side copy in the product: nothing travels
    def kept_labeling(self): return self.labeling() + 100 * self.kept    # (+ when it was kept)


def _striped_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lp_mp_b200 import MpRoundingSolver, StandardVisitor
        from lp_mp_b200.stripes import StripedLP
        H, W, L = 11, 3, 2
        plan = StripePlan(H, world)
        rows = plan.rows(rank)[1]
        lp = HostStripe(rank, rows, W, L, plan.has_above(rank), plan.has_below(rank))
        slp = StripedLP(lp, plan, rank, dist, coupling="halo_nccl", device="cpu").finalize()
        ok = True
        # bound: sum over ranks of (rank + 1) + first entry of the lower neighbour's TOP vector (none for the last rank)
        want = sum(r + 1 for r in range(world)) + sum(1000.0 * (r + 1) + 0.25 for r in range(world - 1))   # (mean of 0.25 * column, W = 3)
        ok &= slp.LowerBound() == want
        ok &= slp.EvaluatePrimal() == float("inf") and not slp.CheckPrimalConsistency()
        slp.set_reparametrization("anisotropic")
        slp.ComputeForwardPass()     # consumes the upper neighbour's BOTTOM vector of pass count 0
        ok &= lp.seen[-1] == (None if rank == 0 else 1000.0 * (rank - 1) + 100.0)
        slp.ComputeBackwardPass()    # consumes the lower neighbour's TOP vector after its one pass, slots right to left: slot 0 = column W - 1
        ok &= lp.seen[-1] == (None if rank == world - 1 else 1000.0 * (rank + 1) + 1.0 + 0.25 * (W - 1))
        slp.ComputeForwardPassAndPrimal(1)
        ok &= slp.CheckPrimalConsistency()
        ok &= slp.EvaluatePrimal() == sum(7.0 * (r + 1) for r in range(world)) + sum(10.0 * (r + 1) for r in range(world - 1))
        lab = slp.labeling()
        ok &= lab.shape == (H * W,) and all((lab[plan.rows(r)[0] * W:(plan.rows(r)[0] + plan.rows(r)[1]) * W] == r).all() for r in range(world))
        ok &= slp.GetNumberOfFactors() == H * W + H * (W - 1) + (H - 1) * W
        # uniform modes: the pass also consumes the NOT yet processed neighbour's vectors across the other boundary, as they are
        # before the pass (3 passes so far on every rank)
        slp.set_reparametrization("damped_uniform")
        slp.ComputeForwardPass()
        ok &= lp.seen[-1] == (None if rank == 0 else 1000.0 * (rank - 1) + 100.0 + 3.0)
        ok &= lp.seen_later == (None if rank == world - 1 else 1000.0 * (rank + 1) + 3.0)
        slp.ComputeBackwardPass()
        ok &= lp.seen[-1] == (None if rank == world - 1 else 1000.0 * (rank + 1) + 4.0 + 0.25 * (W - 1))
        ok &= lp.seen_later == (None if rank == 0 else 1000.0 * (rank - 1) + 100.0 + 4.0 + 0.25 * (W - 1))
        # the reference control law drives a StripedLP (Solver::Solve, include/solver.hxx:230-257), rounding in its default mode
        # (damped_uniform, include/visitors/standard_visitor.hxx:44) and in the anisotropic one
        for opts in (dict(maxIter=4), dict(maxIter=4, roundingReparametrization="anisotropic")):
            lp.ended = False
            s = MpRoundingSolver(slp, StandardVisitor(**opts))
            ok &= bool(s.Solve()) and len(s.history) >= 1 and lp.ended and s.solution_ is not None
            # the solution is the gather of what every stripe KEPT at the last improvement, not of the current labels
            sol = s.solution_
            ok &= sol.shape == (H * W,) and all((sol[plan.rows(r)[0] * W:(plan.rows(r)[0] + plan.rows(r)[1]) * W] % 100 == r).all() for r in range(world))
            ok &= (sol // 100).min() == (sol // 100).max() == lp.kept
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_striped_lp_host_logic_over_gloo(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_striped_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert list(out) == [1] * world
