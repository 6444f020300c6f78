"""Row-stripe partition of one grid LP over the ranks of a ``torch.distributed`` process group (one process per GPU).

The reference has no distribution at all; its closest notion is the partition pass (include/LP_MP.h:1716-1963), which
still runs the groups sequentially.  Here a grid of ``H_total`` rows is cut into ``world`` contiguous row stripes
(SURVEY.md 8e); every rank owns one stripe as an ordinary ``LP`` plus, per stripe boundary, one "in" buffer of
``W*L`` reals that the neighbouring stripe's outgoing message vectors are delivered to.  Three couplings:

``pipelined`` (exact)
    The pass kernel of stripe g stores the outgoing vectors of its last band straight into stripe g+1's in-buffer
    through NVLink peer memory (cudaIpc mapping of the buffer), column by column, while stripe g+1's first band polls
    them: the anti-diagonal wavefront simply continues across GPUs.  Results are bit-identical to the unstriped LP; the
    pass time is the dependency chain through ALL rows, so it does not scale.

``halo`` (per-pass halo, block-Jacobi across stripes; default)
    Before a pass every rank stores the current vectors of its boundary row into the in-buffer of the neighbour that will
    consume them -- one small kernel writing through NVLink peer memory (``lpmp_grid_push_boundary``).  The in-buffer
    slots carry their own validity (sentinel protocol of the pass kernels), so pushes and passes of neighbouring stripes
    order themselves through the data: no collective, no host synchronisation on the data path (slot q = the q-th column
    in the processing order of the consuming pass, i.e. right to left for a backward pass).  All stripes sweep
    concurrently, each using the neighbour's state of the previous pass: always a valid reparametrisation (monotone
    lower bound), but a different fixed point than the sequential order (hard part 2).

``halo_nccl``
    The same semantics with the boundary vectors carried by NCCL send/recv (stream-ordered on the library's stream).

The lower bound / primal cost of the LP is the NCCL all-reduce of the stripes' partial values; everything is enqueued on
the library's stream and read back with ONE stream-ordered copy into pinned memory (LP::LowerBound,
include/LP_MP.h:1507-1518; LP::EvaluatePrimal, :1520-1536).

Uniform / damped-uniform passes (the rounding iterations of the default control law,
include/visitors/standard_visitor.hxx:44): a node also receives from its two not-yet-processed neighbours, so the last
virtual row of a stripe needs the neighbouring stripe's in-side vectors as they were before the pass.  Every rank delivers
them (same push, into the neighbour's OTHER in-buffer) before it launches its pass, in every coupling; with the pipelined
coupling the result stays bit-identical to the unstriped LP.  Potts / truncated-linear grids only (dense-table stripes
are anisotropic-only and raise ``LpmpError``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

TOP, BOTTOM = 0, 1
COUPLINGS = ("halo", "halo_nccl", "pipelined")


@dataclass(frozen=True)
class StripePlan:
    """Which rows a rank owns: contiguous blocks, the first ``H_total % world`` ranks get one row more."""
    H_total: int
    world: int

    def __post_init__(self):
        if self.world < 1 or self.H_total < self.world:
            raise ValueError("need at least one grid row per rank")

    def rows(self, rank: int) -> tuple[int, int]:
        base, extra = divmod(self.H_total, self.world)
        row0 = rank * base + min(rank, extra)
        return row0, base + (1 if rank < extra else 0)

    def owner(self, row: int) -> int:
        for r in range(self.world):
            row0, n = self.rows(r)
            if row0 <= row < row0 + n:
                return r
        raise IndexError(row)

    def has_above(self, rank: int) -> bool:
        return rank > 0

    def has_below(self, rank: int) -> bool:
        return rank + 1 < self.world


class _DevicePtr:
    """Exposes a raw device pointer of the library as a CUDA array so that torch / NCCL can write into it."""

    def __init__(self, ptr: int, count: int, itemsize: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4" if itemsize == 4 else "<f8",
                                         "data": (int(ptr), False), "version": 2}


def neighbour_exchange(dist, rank: int, world: int, send_which: int, send, recv):
    """One shift along the stripe chain: ``send`` (or None) goes to the neighbour across boundary ``send_which``
    (BOTTOM: rank+1, TOP: rank-1); ``recv`` (or None) is filled by the neighbour on the opposite side.  A single
    batched isend/irecv group, so there is no ordering deadlock for any world size.  With NCCL the waits are
    stream-level (the current stream waits for the transfer), not host-level."""
    dst = rank + 1 if send_which == BOTTOM else rank - 1
    src = rank - 1 if send_which == BOTTOM else rank + 1
    ops = []
    if send is not None and 0 <= dst < world:
        ops.append(dist.P2POp(dist.isend, send, dst))
    if recv is not None and 0 <= src < world:
        ops.append(dist.P2POp(dist.irecv, recv, src))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class StripedLP:
    """One stripe per rank behind the LP_TYPE surface that ``Solver`` / ``MpRoundingSolver`` drive (Begin, End,
    set_reparametrization, Compute*Pass[AndPrimal], LowerBound, EvaluatePrimal, CheckPrimalConsistency, labeling).

    ``dist`` is ``torch.distributed`` (initialised, NCCL for GPU runs) or None for a single rank; ``lp`` is this
    rank's ``LP`` with ``grid_create`` + ``grid_set_stripe`` already applied (``make``).  ``device`` is "cuda" for the
    product; "cpu" exists for the host-logic tests (gloo, a stand-in ``lp``)."""

    def __init__(self, lp, plan: StripePlan, rank: int, dist=None, coupling: str = "halo", device: str = "cuda"):
        if coupling not in COUPLINGS:
            raise ValueError(f"coupling must be one of {COUPLINGS}")
        self.lp, self.plan, self.rank, self.dist, self.coupling = lp, plan, rank, dist, coupling
        self.world = plan.world
        self._cuda = device == "cuda"
        self._dev = device
        self._ts = None            # the library's stream as a torch stream
        self._in = {}              # which -> (device pointer, tensor view) of my in-buffers
        self._stage = {}           # which -> compact [W*L] staging tensor (NCCL halo / bound exchange)
        self._peer = {}            # which -> mapped pointer of the neighbour's in-buffer across boundary `which`
        self._ipc_opened = []
        self._setup_done = False
        self._uni = False          # current reparametrisation is uniform / damped-uniform

    # ---- construction -----------------------------------------------------------------------------------------
    @classmethod
    def make(cls, H_total, W, L, pairwise, rank, world, dist=None, dtype="f32", device=0, coupling="halo", tuning=None):
        from .lp import LP
        plan = StripePlan(H_total, world)
        row0, rows = plan.rows(rank)
        lp = LP(dtype=dtype, device=device).grid_create(rows, W, L, pairwise)
        if tuning:
            lp.set_tuning(**tuning)
        if world > 1:
            lp.grid_set_stripe(row0, H_total)
        return cls(lp, plan, rank, dist, coupling)

    def finalize(self):
        self.lp.finalize()
        if self.world > 1:
            self._setup()
        return self

    def _stream_ctx(self):
        if not self._cuda:
            return _NullCtx()
        import torch
        return torch.cuda.stream(self._ts)

    def _setup(self):
        import torch
        lp, plan, rank = self.lp, self.plan, self.rank
        self._itemsize = 4 if lp.dtype == 0 else 8
        self._tdtype = torch.float32 if self._itemsize == 4 else torch.float64
        self._W = lp.grid_shape[1]
        self._count = self._W * lp.L
        if self._cuda:
            self._ts = torch.cuda.ExternalStream(lp.stream(), device=torch.cuda.current_device())
            self._ev = torch.cuda.Event()
        with self._stream_ctx():     # buffers that only ever live on the library's stream are allocated on it
            self._alloc_buffers(torch)
        if self.coupling in ("halo", "pipelined"):
            self._map_peers()
        self._setup_done = True

    def _alloc_buffers(self, torch):
        lp, plan, rank = self.lp, self.plan, self.rank
        for which, present in ((TOP, plan.has_above(rank)), (BOTTOM, plan.has_below(rank))):
            if not present:
                continue
            ptr, nbytes = lp.grid_boundary_ptrs(which)
            assert nbytes == self._count * self._itemsize
            if self._cuda:
                view = torch.as_tensor(_DevicePtr(ptr, self._count, self._itemsize), device="cuda")
            else:   # host-logic tests: the stand-in LP's in-buffer is host memory
                ctype = C.c_float if self._itemsize == 4 else C.c_double
                view = torch.from_numpy(np.ctypeslib.as_array((ctype * self._count).from_address(ptr)))
            self._in[which] = (ptr, view)
            self._stage[which] = torch.empty(self._count, dtype=self._tdtype, device=self._dev)
        # neighbour's side of the boundary edges below me (bound / primal evaluation), labels of its first row
        self._below_b = torch.empty(self._count, dtype=self._tdtype, device=self._dev) if plan.has_below(rank) else None
        self._below_lab = torch.empty(self._W, dtype=torch.int32, device=self._dev) if plan.has_below(rank) else None
        self._my_lab = torch.empty(self._W, dtype=torch.int32, device=self._dev) if plan.has_above(rank) else None
        self._acc = torch.zeros(2, dtype=torch.float64, device=self._dev)
        self._host = torch.zeros(2, dtype=torch.float64, pin_memory=True) if self._cuda else torch.zeros(2, dtype=torch.float64)

    def _map_peers(self):
        """exchange cudaIpc handles of the in-buffers and map the neighbours' buffers"""
        lp, rank = self.lp, self.rank
        mine = {}
        for which, (ptr, _) in self._in.items():
            handle = (C.c_ubyte * 64)()
            lp.api.check(lp.api.ipc_export(C.c_void_p(ptr), handle))
            mine[which] = bytes(handle)
        everyone = [None] * self.world
        self.dist.all_gather_object(everyone, mine)
        for which, nb, nb_which in ((TOP, rank - 1, BOTTOM), (BOTTOM, rank + 1, TOP)):
            if 0 <= nb < self.world and which in self._in:
                peer = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(everyone[nb][nb_which])
                lp.api.check(lp.api.ipc_open(buf, C.byref(peer)))
                self._peer[which] = peer.value
                self._ipc_opened.append(peer.value)
                if self.coupling == "pipelined":   # the pass kernel itself delivers the vectors of its last band
                    lp.grid_set_peer_boundary(which, peer.value)
        self.dist.barrier()

    # ---- halo delivery before a pass -------------------------------------------------------------------------------
    def _deliver(self, send_which: int):
        """what the neighbours need from me before a pass.  ``send_which`` is the boundary whose vectors the pass of the
        neighbour across it CARRIES on (forward: BOTTOM, the lower neighbour continues from my last row; backward: TOP); the
        pipelined coupling delivers those from inside the pass kernel.  Uniform modes: the neighbour across the OTHER boundary
        processes its last virtual row before I touch mine and receives from it -- my vectors there, as they are now, go out
        in every coupling."""
        if self.world == 1:
            return
        backward = send_which == TOP      # the pass about to run, here and on the neighbours
        if self._uni:
            self._deliver_one(TOP if send_which == BOTTOM else BOTTOM, backward, force_push=self.coupling == "pipelined")
        if self.coupling != "pipelined":
            self._deliver_one(send_which, backward)

    def _deliver_one(self, send_which: int, backward: bool, force_push: bool = False):
        """in-buffer slots are in the processing order of the consuming pass: right to left for a backward pass
        (include/lpmp_b200.h, lpmp_grid_push_boundary)"""
        if self.coupling == "halo" or force_push:
            if send_which in self._peer:
                self.lp.grid_push_boundary(send_which, self._peer[send_which], backward)
            return
        recv_which = TOP if send_which == BOTTOM else BOTTOM
        with self._stream_ctx():
            send = None
            if send_which in self._stage:
                self.lp.grid_get_boundary_vectors(send_which, self._stage[send_which].data_ptr())
                send = self._stage[send_which]
                if backward:
                    send = send.view(-1, self.lp.L).flip(0).contiguous().view(-1)
            recv = self._in[recv_which][1] if recv_which in self._in else None
            neighbour_exchange(self.dist, self.rank, self.world, send_which, send, recv)

    # ---- the LP_TYPE surface --------------------------------------------------------------------------------------
    def Begin(self):
        self.lp.Begin()

    def End(self):
        self.lp.End()

    UNIFORM_MODES = True

    def set_reparametrization(self, mode):
        uni = mode in ("uniform", "damped_uniform", 1, 2)
        if self.world > 1 and not uni and mode not in ("anisotropic", 0):
            from ._capi import LpmpError
            raise LpmpError("unsupported: row stripes run the anisotropic, uniform and damped_uniform reparametrizations")
        if self.world > 1 and uni and getattr(self.lp, "pairwise", None) == "dense":
            from ._capi import LpmpError
            raise LpmpError("unsupported: dense-table stripes run the anisotropic reparametrization only "
                            "(use roundingReparametrization='anisotropic' with StandardVisitor)")
        self.lp.set_reparametrization(mode)
        self._uni = uni

    def ComputeForwardPass(self):
        self._deliver(BOTTOM)
        self.lp.ComputeForwardPass()

    def ComputeBackwardPass(self):
        self._deliver(TOP)
        self.lp.ComputeBackwardPass()

    def ComputePass(self, iteration=0):
        self.ComputeForwardPass()
        self.ComputeBackwardPass()

    def ComputeForwardPassAndPrimal(self, iteration=0):
        self._deliver(BOTTOM)
        self.lp.ComputeForwardPassAndPrimal(iteration)

    def ComputeBackwardPassAndPrimal(self, iteration=0):
        self._deliver(TOP)
        self.lp.ComputeBackwardPassAndPrimal(iteration)

    def _exchange_below(self, with_labels: bool):
        """the stripe below sends its side of the shared boundary edges (and the labels of its first row) up"""
        import torch
        send = send_lab = None
        if TOP in self._stage:      # my first row's b side completes the upper neighbour's boundary edges
            self.lp.grid_get_boundary_vectors(TOP, self._stage[TOP].data_ptr())
            send = self._stage[TOP]
            if with_labels:
                self.lp.grid_get_boundary_labels(TOP, self._my_lab.data_ptr())
                send_lab = self._my_lab
        neighbour_exchange(self.dist, self.rank, self.world, TOP, send, self._below_b)
        if with_labels:
            neighbour_exchange(self.dist, self.rank, self.world, TOP, send_lab, self._below_lab)

    def _reduce_and_read(self, n: int):
        """all-reduce the first ``n`` doubles of the accumulator on the library's stream; one stream-ordered copy to the host"""
        self.dist.all_reduce(self._acc[:n])
        self._host.copy_(self._acc, non_blocking=True)
        if self._cuda:
            self._ev.record()

    def LowerBound(self) -> float:
        """LP::LowerBound of the whole grid: partial bounds of the stripes, all-reduced (1 double over NCCL)"""
        if self.world == 1:
            return self.lp.LowerBound()
        with self._stream_ctx():
            self._exchange_below(False)
            self.lp.grid_lower_bound_partial_device(self._below_b.data_ptr() if self._below_b is not None else None,
                                                    self._acc.data_ptr())
            self._reduce_and_read(1)
        if self._cuda:
            self._ev.synchronize()
        return float(self._host[0])

    def EvaluatePrimal(self) -> float:
        """LP::EvaluatePrimal of the whole grid (+inf until a rounding pass has labelled every stripe)"""
        if self.world == 1:
            return self.lp.EvaluatePrimal()
        with self._stream_ctx():
            self._exchange_below(True)
            self.lp.grid_evaluate_primal_partial_device(
                self._below_b.data_ptr() if self._below_b is not None else None,
                self._below_lab.data_ptr() if self._below_lab is not None else None, self._acc.data_ptr())
            self._reduce_and_read(1)
        if self._cuda:
            self._ev.synchronize()
        return float(self._host[0])

    def CheckPrimalConsistency(self) -> bool:
        ok = 1 if self.lp.CheckPrimalConsistency() else 0
        if self.world == 1:
            return bool(ok)
        import torch
        with self._stream_ctx():
            t = torch.tensor([ok], dtype=torch.int32, device=self._dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            res = int(t.item())      # inside the stream context: the copy is ordered after the reduction
        return bool(res)

    def keep_labeling(self):
        """every stripe keeps its part on its device (Solver::RegisterPrimal, include/solver.hxx:321-340): nothing travels"""
        self.lp.keep_labeling()

    def kept_labeling(self):
        return self._gather_labeling(self.lp.kept_labeling())

    def labeling(self):
        """the labeling of the whole grid on every rank"""
        return self._gather_labeling(self.lp.labeling())

    def _gather_labeling(self, mine):
        if self.world == 1:
            return mine
        import torch
        W = self._W
        rows_max = max(self.plan.rows(r)[1] for r in range(self.world))
        pad = torch.zeros(rows_max * W, dtype=torch.int32, device=self._dev)
        pad[:mine.size] = torch.from_numpy(mine.astype(np.int32)).to(self._dev)
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(parts, pad)
        out = [parts[r][: self.plan.rows(r)[1] * W].cpu().numpy().astype(np.uint32) for r in range(self.world)]
        return np.concatenate(out)

    def add_to_constant(self, x):
        if self.rank == 0:       # the constant enters the bound once (stripe 0 adds it to its partial value)
            self.lp.add_to_constant(x)

    def GetNumberOfFactors(self) -> int:
        H, W = self.plan.H_total, self.lp.grid_shape[1]
        return H * W + H * (W - 1) + (H - 1) * W

    def GetNumberOfMessages(self) -> int:
        H, W = self.plan.H_total, self.lp.grid_shape[1]
        return 2 * (H * (W - 1) + (H - 1) * W)

    def synchronize(self):
        self.lp.synchronize()

    def close(self):
        if self._cuda and self._ts is not None:
            self._ts.synchronize()
        for which in list(self._peer):
            if self.coupling == "pipelined":
                self.lp.grid_set_peer_boundary(which, None)
        for ptr in self._ipc_opened:
            self.lp.api.ipc_close(C.c_void_p(ptr))
        self._peer, self._ipc_opened = {}, []
