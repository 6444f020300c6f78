"""Row-stripe partition of one grid LP over the ranks of a ``torch.distributed`` process group (one process per GPU).

The reference has no distribution at all; its closest notion is the partition pass (include/LP_MP.h:1716-1963), which
still runs the groups sequentially.  Here a grid of ``H_total`` rows is cut into ``world`` contiguous row stripes
(SURVEY.md 8e); every rank owns one stripe as an ordinary ``LP`` plus, per stripe boundary, one "in" buffer of
``W*L`` reals that the neighbouring stripe's outgoing message vectors are delivered to.  Two couplings:

``pipelined`` (exact)
    The pass kernel of stripe g stores the outgoing vectors of its last band straight into stripe g+1's in-buffer
    through NVLink peer memory (cudaIpc mapping of the buffer), column by column, while stripe g+1's first band polls
    them: the anti-diagonal wavefront simply continues across GPUs.  Results are bit-identical to the unstriped LP.
    No collective on the data path; NCCL is only used for the lower-bound all-reduce.

``halo`` (per-pass halo, block-Jacobi across stripes)
    Before a pass every rank sends the current vectors of its boundary row to the neighbour that will consume them
    (NCCL send/recv into the in-buffer); all stripes then sweep concurrently, each using the neighbour's state of the
    previous pass.  Always a valid reparametrisation (monotone lower bound), but a different fixed point than the
    sequential order (hard part 2): parity is on the final bound, not per iteration.

Only the anisotropic (SRMP) reparametrisation is supported on stripes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

TOP, BOTTOM = 0, 1


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
    batched isend/irecv group, so there is no ordering deadlock for any world size."""
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


class StripedLP:
    """One stripe per rank behind the LP_TYPE surface (Begin / ComputePass / LowerBound ...).

    ``dist`` is ``torch.distributed`` (initialised, NCCL for GPU runs) or None for a single rank; ``lp`` is this
    rank's ``LP`` with ``grid_create`` + ``grid_set_stripe`` already applied (``make``)."""

    def __init__(self, lp, plan: StripePlan, rank: int, dist=None, coupling: str = "halo"):
        if coupling not in ("halo", "pipelined"):
            raise ValueError("coupling must be 'halo' or 'pipelined'")
        self.lp, self.plan, self.rank, self.dist, self.coupling = lp, plan, rank, dist, coupling
        self.world = plan.world
        self._torch_stream = None
        self._in = {}
        self._stage = {}
        self._peer = {}
        self._setup_done = False

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

    def _setup(self):
        import torch
        lp, plan, rank = self.lp, self.plan, self.rank
        self._itemsize = 4 if lp.dtype == 0 else 8
        self._tdtype = torch.float32 if self._itemsize == 4 else torch.float64
        self._count = lp.grid_shape[1] * lp.L
        self._torch_stream = torch.cuda.ExternalStream(lp.stream())
        for which, present in ((TOP, plan.has_above(rank)), (BOTTOM, plan.has_below(rank))):
            if not present:
                continue
            ptr, nbytes = lp.grid_boundary_ptrs(which)
            assert nbytes == self._count * self._itemsize
            self._in[which] = (ptr, torch.as_tensor(_DevicePtr(ptr, self._count, self._itemsize), device="cuda"))
            self._stage[which] = torch.empty(self._count, dtype=self._tdtype, device="cuda")
        if self.coupling == "pipelined":
            # exchange cudaIpc handles of the in-buffers; map the neighbours' buffers and hand them to the pass kernel
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
                    lp.grid_set_peer_boundary(which, peer.value)
            self.dist.barrier()
        self._setup_done = True

    # ---- halo exchange (coupling == "halo") ---------------------------------------------------------------------
    def _exchange(self, send_which: int):
        """every rank sends the vectors of boundary ``send_which`` to the neighbour across it, which receives them
        into its in-buffer of the opposite boundary; enqueued on the library's stream"""
        import torch
        recv_which = TOP if send_which == BOTTOM else BOTTOM
        with torch.cuda.stream(self._torch_stream):
            send = None
            if send_which in self._stage:
                self.lp.grid_get_boundary_vectors(send_which, self._stage[send_which].data_ptr())
                send = self._stage[send_which]
            recv = self._in[recv_which][1] if recv_which in self._in else None
            neighbour_exchange(self.dist, self.rank, self.world, send_which, send, recv)

    # ---- the LP_TYPE surface --------------------------------------------------------------------------------------
    def Begin(self):
        self.lp.Begin()

    def set_reparametrization(self, mode):
        if self.world > 1 and mode not in ("anisotropic", 0):
            raise ValueError("stripes support the anisotropic reparametrization only")
        self.lp.set_reparametrization(mode)

    def ComputeForwardPass(self):
        if self.world > 1 and self.coupling == "halo":
            self._exchange(BOTTOM)      # the forward pass consumes the upper neighbour's last-row vectors
        self.lp.ComputeForwardPass()

    def ComputeBackwardPass(self):
        if self.world > 1 and self.coupling == "halo":
            self._exchange(TOP)         # the backward pass consumes the lower neighbour's first-row vectors
        self.lp.ComputeBackwardPass()

    def ComputePass(self, iteration=0):
        self.ComputeForwardPass()
        self.ComputeBackwardPass()

    def LowerBound(self) -> float:
        """LP::LowerBound of the whole grid: partial bounds of the stripes, all-reduced (1 double over NCCL)"""
        if self.world == 1:
            return self.lp.LowerBound()
        import torch
        dist, rank = self.dist, self.rank
        below = None
        with torch.cuda.stream(self._torch_stream):
            send = None
            if TOP in self._stage:      # my first row's b side completes the upper neighbour's boundary edges
                self.lp.grid_get_boundary_vectors(TOP, self._stage[TOP].data_ptr())
                send = self._stage[TOP]
            if BOTTOM in self._stage:
                below = self._stage[BOTTOM]
            neighbour_exchange(dist, rank, self.world, TOP, send, below)
            part = self.lp.grid_lower_bound_partial(below.data_ptr() if below is not None else None)
            t = torch.tensor([part], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
        return float(t.item())

    def synchronize(self):
        self.lp.synchronize()

    def close(self):
        for which, ptr in self._peer.items():
            self.lp.grid_set_peer_boundary(which, None)
            self.lp.api.ipc_close(C.c_void_p(ptr))
        self._peer = {}
