"""Host-side mirror of the reference's ``LP<FMC>`` (the LP_TYPE consumed by ``Solver<LP_TYPE,VISITOR>``,
/root/reference/include/solver.hxx:29-30) for the SRMP simplex configuration.

Method names, argument meaning and error behaviour follow the reference (include/LP_MP.h:200-585) so that the
parity tests read like the reference's own tests (test/graphical_model.cpp:98-136).  Every method is one call
into the C ABI (include/lpmp_b200.h); nothing is computed in Python.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as K


def _pd(x):
    return None if x is None else x.ctypes.data_as(C.POINTER(C.c_double))


def _pu64(x):
    return x.ctypes.data_as(C.POINTER(C.c_uint64))


def _f64(x, shape=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return x if shape is None else x.reshape(shape)


class LP:
    """LP<FMC_SRMP>: unary + pairwise simplex factors tied by marginalisation messages."""

    def __init__(self, dtype="f64", device=0, capi=None):
        if capi is None:
            from . import load_library
            capi = load_library()
        self.api = capi
        self.dtype = {"f32": K.F32, "f64": K.F64, K.F32: K.F32, K.F64: K.F64}[dtype]
        h = C.c_void_p()
        capi.check(capi.create(self.dtype, int(device), C.byref(h)))
        self.h = h
        self.N = self.E = self.L = 0
        self.label_counts = None        # per-variable label counts of an LP with mixed counts (add_unaries_var), else None
        self._edge_list = []            # (i, j) per edge, kept for the packed layouts of mixed-count LPs
        self._pending_unaries, self._pending_pairwise = [], []
        self.grid_shape = None
        self.finalized = False

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.api.destroy(h)

    def _ck(self, status):
        self.api.check(status, self.h)

    # ---- problem construction ---------------------------------------------------------------------------
    def add_unaries(self, theta):
        theta = _f64(theta)
        n, L = theta.shape
        self._ck(self.api.add_unaries(self.h, n, L, _pd(theta)))
        self.N += n
        self.L = L

    def add_unaries_var(self, costs):
        """variables with their own label counts (test/graphical_model.cpp:74-96): `costs` is a list of 1-D arrays"""
        costs = [np.asarray(c, dtype=np.float64).ravel() for c in costs]
        counts = np.array([len(c) for c in costs], dtype=np.uint32)
        packed = _f64(np.concatenate(costs)) if costs else np.zeros(0)
        self._ck(self.api.add_unaries_var(self.h, len(costs), counts.ctypes.data_as(C.POINTER(C.c_uint32)), _pd(packed)))
        if self.label_counts is None:
            self.label_counts = [self.L] * self.N
        self.label_counts += [int(c) for c in counts]
        self.N += len(costs)
        self.L = max(self.label_counts)

    def _edges(self, i, j, record=False):
        i = np.ascontiguousarray(i, dtype=np.uint64)
        j = np.ascontiguousarray(j, dtype=np.uint64)
        assert i.shape == j.shape and i.ndim == 1
        if record and self.label_counts is not None:
            self._edge_list += list(zip(i.tolist(), j.tolist()))
        return i, j

    def add_pairwise_dense(self, i, j, tables):
        """tables: [m][L][L]; with mixed label counts a list of L_i x L_j arrays (passed packed)"""
        i, j = self._edges(i, j, record=True)
        if self.label_counts is not None:
            tabs = [np.asarray(t, dtype=np.float64) for t in tables]
            for t, a, b in zip(tabs, i, j):
                assert t.shape == (self.label_counts[int(a)], self.label_counts[int(b)])
            tables = _f64(np.concatenate([t.ravel() for t in tabs]))
        else:
            tables = _f64(tables, (len(i), self.L, self.L))
        self._ck(self.api.add_pairwise_dense(self.h, len(i), _pu64(i), _pu64(j), _pd(tables)))
        self.E += len(i)

    def add_pairwise_potts(self, i, j, lam):
        i, j = self._edges(i, j, record=True)
        lam = _f64(lam, (len(i),))
        self._ck(self.api.add_pairwise_potts(self.h, len(i), _pu64(i), _pu64(j), _pd(lam)))
        self.E += len(i)

    def add_pairwise_trunclin(self, i, j, slope, trunc):
        i, j = self._edges(i, j, record=True)
        slope = _f64(np.broadcast_to(slope, (len(i),)))
        trunc = _f64(np.broadcast_to(trunc, (len(i),)))
        self._ck(self.api.add_pairwise_trunclin(self.h, len(i), _pu64(i), _pu64(j), _pd(slope), _pd(trunc)))
        self.E += len(i)

    # mrf.AddUnaryFactor / mrf.AddPairwiseFactor of the MRF problem constructor (test/graphical_model.cpp:98-108)
    def AddUnaryFactor(self, cost):
        assert not self._pending_pairwise, "add all unary factors first"
        self._pending_unaries.append(np.asarray(cost, dtype=np.float64))
        return len(self._pending_unaries) - 1

    def AddPairwiseFactor(self, i, j, cost):
        self._pending_pairwise.append((int(i), int(j), np.asarray(cost, dtype=np.float64)))
        return len(self._pending_pairwise) - 1

    def _flush_pending(self):
        if self._pending_unaries:
            if len({len(c) for c in self._pending_unaries}) == 1 and self.label_counts is None:
                self.add_unaries(np.stack(self._pending_unaries))
            else:
                self.add_unaries_var(self._pending_unaries)
            self._pending_unaries = []
        if self._pending_pairwise:
            i, j, t = zip(*self._pending_pairwise)
            self.add_pairwise_dense(np.array(i), np.array(j), list(t) if self.label_counts is not None else np.stack(t))
            self._pending_pairwise = []

    def set_node_ranks(self, rank):
        self._flush_pending()
        rank = np.ascontiguousarray(rank, dtype=np.uint64)
        assert rank.shape == (self.N,)
        self._ck(self.api.set_node_ranks(self.h, _pu64(rank)))

    def set_edge_order(self, order):
        self._ck(self.api.set_edge_order(self.h, K.EDGE_ORDER[order]))

    def set_schedule(self, schedule):
        self._ck(self.api.set_schedule(self.h, K.SCHEDULE[schedule]))

    def set_tuning(self, **kw):
        for k, v in kw.items():
            self._ck(self.api.set_tuning(self.h, k.encode(), int(v)))

    def greedy_colouring(self, i, j):
        i, j = self._edges(i, j)
        colour = np.zeros(self.N + len(self._pending_unaries), dtype=np.uint32)
        nc = C.c_uint32()
        self._ck(self.api.greedy_colouring(len(colour), len(i), _pu64(i), _pu64(j),
                                           colour.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(nc)))
        return colour, nc.value

    # ---- grids ------------------------------------------------------------------------------------------
    def grid_create(self, H, W, L, pairwise):
        self._ck(self.api.grid_create(self.h, H, W, L, K.PAIRWISE[pairwise]))
        self.grid_shape = (H, W)
        self.N, self.L = H * W, L
        self.E = H * (W - 1) + (H - 1) * W
        self.pairwise = pairwise
        return self

    def _src(self, x):
        x = np.ascontiguousarray(x)
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
        return x, (K.F32 if x.dtype == np.float32 else K.F64)

    def grid_set_unaries(self, theta):
        H, W = self.grid_shape
        theta, src = self._src(theta)
        assert theta.size == H * W * self.L
        self._ck(self.api.grid_set_unaries(self.h, theta.ctypes.data, src))

    def grid_set_pairwise(self, horizontal, vertical):
        H, W = self.grid_shape
        P = {"dense": self.L * self.L, "potts": 1, "trunclin": 2}[self.pairwise]
        horizontal, s0 = self._src(horizontal)
        vertical, s1 = self._src(vertical)
        if s0 != s1:
            horizontal, vertical, s0 = horizontal.astype(np.float64), vertical.astype(np.float64), K.F64
        assert horizontal.size == H * (W - 1) * P and vertical.size == (H - 1) * W * P
        self._ck(self.api.grid_set_pairwise(self.h, horizontal.ctypes.data, vertical.ctypes.data, s0))

    def grid_fill_synthetic(self, seed, p_lo, p_hi):
        self._ck(self.api.grid_fill_synthetic(self.h, int(seed), float(p_lo), float(p_hi)))

    # ---- row stripes (multi-GPU) -------------------------------------------------------------------------
    def grid_set_stripe(self, row0, H_total):
        self._ck(self.api.grid_set_stripe(self.h, int(row0), int(H_total)))

    def grid_set_boundary_pairwise(self, top, bottom):
        top = None if top is None else np.ascontiguousarray(top, dtype=np.float64)
        bottom = None if bottom is None else np.ascontiguousarray(bottom, dtype=np.float64)
        self._ck(self.api.grid_set_boundary_pairwise(self.h, None if top is None else top.ctypes.data,
                                                     None if bottom is None else bottom.ctypes.data, K.F64))

    def grid_boundary_ptrs(self, which):
        """(device pointer, bytes) of the in-buffer of boundary `which` (0 = top, 1 = bottom)"""
        buf, nbytes = C.c_void_p(), C.c_uint64()
        self._ck(self.api.grid_boundary_ptrs(self.h, int(which), C.byref(buf), C.byref(nbytes)))
        return buf.value, nbytes.value

    def grid_set_peer_boundary(self, which, peer_ptr):
        self._ck(self.api.grid_set_peer_boundary(self.h, int(which), C.c_void_p(peer_ptr or 0)))

    def grid_get_boundary_vectors(self, which, dst_ptr):
        self._ck(self.api.grid_get_boundary_vectors(self.h, int(which), C.c_void_p(dst_ptr)))

    def grid_lower_bound_partial(self, below_ptr=None) -> float:
        out = C.c_double()
        self._ck(self.api.grid_lower_bound_partial(self.h, C.c_void_p(below_ptr or 0), C.byref(out)))
        return out.value

    def grid_lower_bound_partial_device(self, below_ptr, out_ptr):
        """asynchronous: the stripe's share of LP::LowerBound is left in the device double at `out_ptr`"""
        self._ck(self.api.grid_lower_bound_partial_device(self.h, C.c_void_p(below_ptr or 0), C.c_void_p(out_ptr)))

    def grid_evaluate_primal_partial_device(self, below_ptr, below_labels_ptr, out_ptr):
        self._ck(self.api.grid_evaluate_primal_partial_device(self.h, C.c_void_p(below_ptr or 0), C.c_void_p(below_labels_ptr or 0), C.c_void_p(out_ptr)))

    def grid_get_boundary_labels(self, which, dst_ptr):
        self._ck(self.api.grid_get_boundary_labels(self.h, int(which), C.c_void_p(dst_ptr)))

    def grid_push_boundary(self, which, peer_ptr, backward):
        """``backward``: the pass of the neighbour that consumes the vectors is a backward pass (slots right to left)"""
        self._ck(self.api.grid_push_boundary(self.h, int(which), C.c_void_p(peer_ptr), int(bool(backward))))

    def finalize(self):
        self._flush_pending()
        self._ck(self.api.finalize(self.h))
        self.finalized = True
        return self

    # ---- the LP_TYPE surface (include/LP_MP.h) ------------------------------------------------------------
    def Begin(self):
        if not self.finalized:
            self.finalize()
        self._ck(self.api.begin(self.h))

    def End(self):
        self._ck(self.api.end(self.h))

    def set_reparametrization(self, mode):
        self._ck(self.api.set_reparametrization(self.h, K.REPAM[mode] if isinstance(mode, str) else int(mode)))

    def set_send_mode(self, mode):
        """--reparametrizationType of the reference LP: "shared" (default) or "residual" (include/LP_MP.h:704-722)"""
        self._ck(self.api.set_send_mode(self.h, K.SEND_MODE[mode] if isinstance(mode, str) else int(mode)))

    def put_in_same_partition(self, i, j):
        """LP::put_in_same_partition on the unary factors of variables i[k], j[k] (include/LP_MP.h:465)"""
        self._flush_pending()
        i, j = self._edges(np.atleast_1d(i), np.atleast_1d(j))
        self._ck(self.api.put_in_same_partition(self.h, len(i), _pu64(i), _pu64(j)))

    def compute_partition_pass(self, inner_iterations=5, overlapping=False):
        """LP::compute_partition_pass / compute_overlapping_partition_pass (include/LP_MP.h:1961-2051)"""
        self._ck(self.api.compute_partition_pass(self.h, int(inner_iterations), 1 if overlapping else 0))

    def ComputePass(self, iteration=0):
        self._ck(self.api.compute_pass(self.h, int(iteration)))

    def ComputeForwardPass(self):
        self._ck(self.api.compute_forward_pass(self.h))

    def ComputeBackwardPass(self):
        self._ck(self.api.compute_backward_pass(self.h))

    def ComputeForwardPassAndPrimal(self, iteration):
        self._ck(self.api.compute_forward_pass_and_primal(self.h, int(iteration)))

    def ComputeBackwardPassAndPrimal(self, iteration):
        self._ck(self.api.compute_backward_pass_and_primal(self.h, int(iteration)))

    def LowerBound(self) -> float:
        out = C.c_double()
        self._ck(self.api.lower_bound(self.h, C.byref(out)))
        return out.value

    def EvaluatePrimal(self) -> float:
        out = C.c_double()
        self._ck(self.api.evaluate_primal(self.h, C.byref(out)))
        return out.value

    def CheckPrimalConsistency(self) -> bool:
        out = C.c_int()
        self._ck(self.api.check_primal_consistency(self.h, C.byref(out)))
        return bool(out.value)

    def add_to_constant(self, x):
        self._ck(self.api.add_to_constant(self.h, float(x)))

    def GetNumberOfFactors(self) -> int:
        out = C.c_uint64()
        self._ck(self.api.num_factors(self.h, C.byref(out)))
        return out.value

    def GetNumberOfMessages(self) -> int:
        out = C.c_uint64()
        self._ck(self.api.num_messages(self.h, C.byref(out)))
        return out.value

    # ---- state ------------------------------------------------------------------------------------------
    def labeling(self, out=None):
        """``out``: a contiguous uint32 array of N entries to fill (e.g. the numpy view of a pinned tensor), else a new array"""
        if out is None:
            out = np.empty(self.N, dtype=np.uint32)
        assert out.dtype == np.uint32 and out.size == self.N and out.flags["C_CONTIGUOUS"]
        self._ck(self.api.get_labeling(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def keep_labeling(self):
        """device-side copy of the current labeling (what Solver::RegisterPrimal keeps, include/solver.hxx:321-340)"""
        self._ck(self.api.keep_labeling(self.h))

    def kept_labeling(self):
        out = np.zeros(self.N, dtype=np.uint32)
        self._ck(self.api.get_kept_labeling(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def _unpack(self, packed, counts):
        """packed per-variable / per-edge vectors of a mixed-count LP -> [n][L] padded with zeros (as checkers.RefLP reports them)"""
        out = np.zeros((len(counts), self.L), dtype=np.float64)
        k = 0
        for r, c in enumerate(counts):
            out[r, :c] = packed[k:k + c]
            k += c
        return out

    def unaries(self):
        if self.label_counts is not None:
            packed = np.zeros(sum(self.label_counts), dtype=np.float64)
            self._ck(self.api.get_unaries(self.h, _pd(packed)))
            return self._unpack(packed, self.label_counts)
        out = np.zeros((self.N, self.L), dtype=np.float64)
        self._ck(self.api.get_unaries(self.h, _pd(out)))
        return out

    def messages(self):
        if self.label_counts is not None:
            ca = [self.label_counts[i] for i, _ in self._edge_list]
            cb = [self.label_counts[j] for _, j in self._edge_list]
            pa, pb = np.zeros(sum(ca), dtype=np.float64), np.zeros(sum(cb), dtype=np.float64)
            self._ck(self.api.get_messages(self.h, _pd(pa), _pd(pb)))
            return self._unpack(pa, ca), self._unpack(pb, cb)
        a = np.zeros((self.E, self.L), dtype=np.float64)
        b = np.zeros((self.E, self.L), dtype=np.float64)
        self._ck(self.api.get_messages(self.h, _pd(a), _pd(b)))
        return a, b

    # dual variables in the reference's archive layout (include/factor_archive.hxx:34-48): checkpoint / warm start
    def dual_size(self) -> int:
        out = C.c_uint64()
        self._ck(self.api.dual_size(self.h, C.byref(out)))
        return out.value

    def save_dual(self) -> bytes:
        buf = np.zeros(self.dual_size() // 8, dtype=np.float64)
        self._ck(self.api.save_dual(self.h, buf.ctypes.data, buf.nbytes))
        return buf.tobytes()

    def load_dual(self, archive: bytes):
        buf = np.frombuffer(archive, dtype=np.float64).copy()
        self._ck(self.api.load_dual(self.h, buf.ctypes.data, buf.nbytes))

    def set_messages(self, a, b):
        a, b = _f64(a, (self.E, self.L)), _f64(b, (self.E, self.L))
        self._ck(self.api.set_messages(self.h, _pd(a), _pd(b)))

    def set_unaries(self, theta):
        theta = _f64(theta, (self.N, self.L))
        self._ck(self.api.set_unaries(self.h, _pd(theta)))

    def stats(self):
        s = K.Stats()
        self._ck(self.api.get_stats(self.h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in K.Stats._fields_}

    def synchronize(self):
        self._ck(self.api.synchronize(self.h))

    def trace(self, bands):
        out = np.zeros((bands, 4), dtype=np.uint64)
        self._ck(self.api.get_trace(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size))
        return out

    def trace_steps(self, bands):
        """(wavefront_trace=2) end-of-step globaltimer stamps [2][8192] of bands ``bands // 2`` and ``bands // 2 + 1``"""
        out = np.zeros(bands * 4 + 2 * 8192, dtype=np.uint64)
        self._ck(self.api.get_trace(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size))
        return out[bands * 4:].reshape(2, 8192)

    def stream(self) -> int:
        out = C.c_void_p()
        self._ck(self.api.stream(self.h, C.byref(out)))
        return out.value or 0
