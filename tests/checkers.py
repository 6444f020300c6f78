"""Test-side access to the CPU checkers under oracle/ (the product package never imports this)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_oracle = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def oracle_capi():
    """Standalone restatement (oracle/lpmp_oracle.cpp) behind the product's ctypes binding."""
    global _oracle
    if _oracle is None:
        from lp_mp_b200._capi import CApi
        path = os.path.join(ORACLE_DIR, "liblpmp_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        _oracle = CApi(lib, "lpmp_oracle_")
        lib.lpmp_oracle_force_dense.argtypes = [C.c_void_p, C.c_int]
    return _oracle


def oracle_lp(dtype="f64", **kw):
    from lp_mp_b200.lp import LP
    return LP(dtype=dtype, capi=oracle_capi(), **kw)


def ref_library(fifo=False, cuda_adapter=False):
    """the reference core (oracle/ref_driver.cpp); cuda_adapter: the same driver -- the reference's unmodified Solver /
    MpRoundingSolver / StandardVisitor -- around the product's binding include/LP_cuda.hxx (needs a GPU at run time)"""
    name = "liblpmp_ref_cuda.so" if cuda_adapter else ("liblpmp_ref_fifo.so" if fifo else "liblpmp_ref.so")
    path = os.path.join(ORACLE_DIR, "_ref", name)
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/include"):
            build_oracle()
        if not os.path.exists(path):
            return None
    lib = C.CDLL(path)
    vp, u64, pd, d = C.c_void_p, C.c_uint64, C.POINTER(C.c_double), C.c_double
    sig = {
        "lpmp_ref_create": ([C.c_int, C.POINTER(C.c_char_p)], vp),
        "lpmp_ref_destroy": ([vp], None),
        "lpmp_ref_add_unary": ([vp, u64, pd], C.c_int64),
        "lpmp_ref_add_pairwise": ([vp, u64, u64, pd], C.c_int64),
        "lpmp_ref_add_to_constant": ([vp, d], None),
        "lpmp_ref_begin": ([vp], None),
        "lpmp_ref_set_reparametrization": ([vp, C.c_int], None),
        "lpmp_ref_compute_pass": ([vp, u64], None),
        "lpmp_ref_forward_pass": ([vp], None),
        "lpmp_ref_backward_pass": ([vp], None),
        "lpmp_ref_forward_pass_and_primal": ([vp, u64], None),
        "lpmp_ref_backward_pass_and_primal": ([vp, u64], None),
        "lpmp_ref_lower_bound": ([vp], d),
        "lpmp_ref_evaluate_primal": ([vp], d),
        "lpmp_ref_num_factors": ([vp], u64),
        "lpmp_ref_num_messages": ([vp], u64),
        "lpmp_ref_no_send_messages": ([vp, u64], u64),
        "lpmp_ref_get_unary": ([vp, u64, pd], None),
        "lpmp_ref_get_pairwise_msgs": ([vp, u64, pd, pd], None),
        "lpmp_ref_get_labels": ([vp, C.POINTER(u64)], None),
        "lpmp_ref_unary_lower_bound": ([u64, pd], d),
        "lpmp_ref_pairwise_min_marginal": ([u64, u64, pd, pd, pd, C.c_int, pd, pd], None),
        "lpmp_ref_solve": ([vp, pd, pd, C.POINTER(C.c_int32), u64, pd, pd, pd], u64),
        "lpmp_ref_put_in_same_partition": ([vp, u64, u64], None),
        "lpmp_ref_dual_size": ([vp], u64),
        "lpmp_ref_save_dual": ([vp, vp], None),
        "lpmp_ref_load_dual": ([vp, vp], None),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, res
    return lib


def _pd(x):
    return None if x is None else x.ctypes.data_as(C.POINTER(C.c_double))


class RefLP:
    """The reference's own LP<FMC>/Solver (oracle/ref_driver.cpp) with the method names of lp_mp_b200.LP."""

    MODES = {"anisotropic": 0, "uniform": 1, "damped_uniform": 2, "anisotropic2": 3}

    def __init__(self, lib, options=()):
        self.lib = lib
        opts = [b"lpmp_ref"] + [str(o).encode() for o in options]
        arr = (C.c_char_p * len(opts))(*opts)
        self.h = lib.lpmp_ref_create(len(opts), arr)
        self.N = self.E = 0
        self.Ls = []

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.lpmp_ref_destroy(self.h)
            self.h = None

    def AddUnaryFactor(self, cost):
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        self.Ls.append(len(cost))
        self.N += 1
        return self.lib.lpmp_ref_add_unary(self.h, len(cost), _pd(cost))

    def AddPairwiseFactor(self, i, j, table):
        table = np.ascontiguousarray(table, dtype=np.float64)
        assert table.shape == (self.Ls[i], self.Ls[j])
        r = self.lib.lpmp_ref_add_pairwise(self.h, i, j, _pd(table))
        assert r >= 0
        self.E += 1
        self.edges = getattr(self, "edges", []) + [(i, j)]
        return r

    def put_in_same_partition(self, i, j):
        for a, b in zip(np.atleast_1d(i), np.atleast_1d(j)):
            self.lib.lpmp_ref_put_in_same_partition(self.h, int(a), int(b))

    def Begin(self): self.lib.lpmp_ref_begin(self.h)
    def End(self): pass
    def set_reparametrization(self, mode): self.lib.lpmp_ref_set_reparametrization(self.h, self.MODES[mode])
    def ComputePass(self, it=0): self.lib.lpmp_ref_compute_pass(self.h, it)
    def ComputeForwardPass(self): self.lib.lpmp_ref_forward_pass(self.h)
    def ComputeBackwardPass(self): self.lib.lpmp_ref_backward_pass(self.h)
    def ComputeForwardPassAndPrimal(self, it): self.lib.lpmp_ref_forward_pass_and_primal(self.h, it)
    def ComputeBackwardPassAndPrimal(self, it): self.lib.lpmp_ref_backward_pass_and_primal(self.h, it)
    def LowerBound(self): return self.lib.lpmp_ref_lower_bound(self.h)
    def EvaluatePrimal(self): return self.lib.lpmp_ref_evaluate_primal(self.h)
    def CheckPrimalConsistency(self): return True
    def GetNumberOfFactors(self): return self.lib.lpmp_ref_num_factors(self.h)
    def GetNumberOfMessages(self): return self.lib.lpmp_ref_num_messages(self.h)

    def unaries(self):
        out = np.zeros((self.N, max(self.Ls)))
        for i in range(self.N):
            row = np.zeros(self.Ls[i])
            self.lib.lpmp_ref_get_unary(self.h, i, _pd(row))
            out[i, :self.Ls[i]] = row
        return out

    def messages(self):
        L = max(self.Ls)
        a, b = np.zeros((self.E, L)), np.zeros((self.E, L))
        for e, (i, j) in enumerate(self.edges):
            ra, rb = np.zeros(self.Ls[i]), np.zeros(self.Ls[j])
            self.lib.lpmp_ref_get_pairwise_msgs(self.h, e, _pd(ra), _pd(rb))
            a[e, :len(ra)], b[e, :len(rb)] = ra, rb
        return a, b

    def labeling(self):
        out = np.zeros(self.N, dtype=np.uint64)
        self.lib.lpmp_ref_get_labels(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out

    def save_dual(self) -> bytes:
        buf = np.zeros(self.lib.lpmp_ref_dual_size(self.h) // 8, dtype=np.float64)
        self.lib.lpmp_ref_save_dual(self.h, buf.ctypes.data)
        return buf.tobytes()

    def load_dual(self, archive: bytes):
        buf = np.frombuffer(archive, dtype=np.float64).copy()
        assert buf.nbytes == self.lib.lpmp_ref_dual_size(self.h)
        self.lib.lpmp_ref_load_dual(self.h, buf.ctypes.data)

    def solve(self, cap=100000):
        lb, ub = np.zeros(cap), np.zeros(cap)
        flags = np.zeros(cap, dtype=np.int32)
        flb, fub, sec = C.c_double(), C.c_double(), C.c_double()
        n = self.lib.lpmp_ref_solve(self.h, _pd(lb), _pd(ub), flags.ctypes.data_as(C.POINTER(C.c_int32)), cap,
                                    C.byref(flb), C.byref(fub), C.byref(sec))
        return dict(lb=lb[:n].copy(), ub=ub[:n].copy(), flags=flags[:n].copy(), final_lb=flb.value,
                    final_ub=fub.value, seconds=sec.value)


# ---- problem generators shared by the tests ---------------------------------------------------------------
def grid_edges(H, W):
    """Edge list in the canonical insertion order: per node (row-major) horizontal edge, then vertical edge."""
    i, j = [], []
    for r in range(H):
        for c in range(W):
            n = r * W + c
            if c + 1 < W:
                i.append(n); j.append(n + 1)
            if r + 1 < H:
                i.append(n); j.append(n + W)
    return np.array(i, dtype=np.uint64), np.array(j, dtype=np.uint64)


def potts_table(L, lam):
    t = np.full((L, L), float(lam))
    np.fill_diagonal(t, 0.0)
    return t


def trunclin_table(L, slope, trunc, dtype=np.float64):
    d = np.abs(np.arange(L)[:, None] - np.arange(L)[None, :]).astype(dtype)
    return np.minimum(dtype(slope) * d, dtype(trunc)).astype(np.float64)
