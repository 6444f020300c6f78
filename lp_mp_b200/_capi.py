"""ctypes binding of the C ABI declared in include/lpmp_b200.h.

The binding is parametrised by (shared library, symbol prefix) because the same header is implemented by the
CUDA library (prefix ``lpmp_``) and -- for the tests only -- by the CPU checker.  This module never names or
loads the checker; ``lp_mp_b200.load_library()`` only ever opens ``lp_mp_b200/liblpmp_b200.so`` and raises if
it is missing (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

F32, F64 = 0, 1
REPAM = {"anisotropic": 0, "uniform": 1, "damped_uniform": 2, "anisotropic2": 3}
SEND_MODE = {"shared": 0, "residual": 1, "adaptive": 2}
PAIRWISE = {"dense": 0, "potts": 1, "trunclin": 2}
SCHEDULE = {"levels": 0, "wavefront": 1}
EDGE_ORDER = {"lifo": 0, "fifo": 1}
STATUS = {0: "ok", 1: "invalid argument", 2: "bad state", 3: "CUDA error", 4: "graph not a dag", 5: "unsupported"}


class Stats(C.Structure):
    _fields_ = [
        ("num_nodes", C.c_uint64), ("num_edges", C.c_uint64), ("num_labels", C.c_uint64),
        ("num_levels", C.c_uint64), ("kernel_launches", C.c_uint64), ("message_updates", C.c_uint64),
        ("device_bytes", C.c_uint64), ("algorithmic_bytes_per_pass", C.c_double),
    ]


_vp, _u64, _u32, _i, _d = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_double
_pd, _pu64, _pu32 = C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)

# name -> argtypes (restype is int unless listed in _RESTYPE); mirrors include/lpmp_b200.h line by line
SIGNATURES = {
    "create": [_i, _i, C.POINTER(_vp)],
    "destroy": [_vp],
    "last_error": [_vp],
    "add_unaries": [_vp, _u64, _u32, _pd],
    "add_unaries_var": [_vp, _u64, _pu32, _pd],
    "add_pairwise_dense": [_vp, _u64, _pu64, _pu64, _pd],
    "add_pairwise_potts": [_vp, _u64, _pu64, _pu64, _pd],
    "add_pairwise_trunclin": [_vp, _u64, _pu64, _pu64, _pd, _pd],
    "set_node_ranks": [_vp, _pu64],
    "set_edge_order": [_vp, _i],
    "set_schedule": [_vp, _i],
    "set_tuning": [_vp, C.c_char_p, _i],
    "greedy_colouring": [_u64, _u64, _pu64, _pu64, _pu32, _pu32],
    "grid_create": [_vp, _u32, _u32, _u32, _i],
    "grid_set_unaries": [_vp, _vp, _i],
    "grid_set_pairwise": [_vp, _vp, _vp, _i],
    "grid_fill_synthetic": [_vp, _u64, _d, _d],
    "grid_set_stripe": [_vp, _u32, _u32],
    "grid_set_boundary_pairwise": [_vp, _vp, _vp, _i],
    "grid_boundary_ptrs": [_vp, _i, C.POINTER(_vp), _pu64],
    "grid_set_peer_boundary": [_vp, _i, _vp],
    "grid_get_boundary_vectors": [_vp, _i, _vp],
    "grid_lower_bound_partial": [_vp, _vp, _pd],
    "grid_lower_bound_partial_device": [_vp, _vp, _vp],
    "grid_evaluate_primal_partial_device": [_vp, _vp, _vp, _vp],
    "grid_get_boundary_labels": [_vp, _i, _vp],
    "grid_push_boundary": [_vp, _i, _vp, _i],
    "ipc_export": [_vp, _vp],
    "ipc_open": [_vp, C.POINTER(_vp)],
    "ipc_close": [_vp],
    "finalize": [_vp],
    "begin": [_vp],
    "end": [_vp],
    "set_reparametrization": [_vp, _i],
    "set_send_mode": [_vp, _i],
    "compute_pass": [_vp, _u64],
    "compute_forward_pass": [_vp],
    "compute_backward_pass": [_vp],
    "compute_forward_pass_and_primal": [_vp, _u64],
    "compute_backward_pass_and_primal": [_vp, _u64],
    "put_in_same_partition": [_vp, _u64, _pu64, _pu64],
    "compute_partition_pass": [_vp, _u64, _i],
    "lower_bound": [_vp, _pd],
    "evaluate_primal": [_vp, _pd],
    "check_primal_consistency": [_vp, C.POINTER(_i)],
    "add_to_constant": [_vp, _d],
    "num_factors": [_vp, _pu64],
    "num_messages": [_vp, _pu64],
    "get_labeling": [_vp, _pu32],
    "keep_labeling": [_vp],
    "get_kept_labeling": [_vp, _pu32],
    "get_unaries": [_vp, _pd],
    "get_messages": [_vp, _pd, _pd],
    "set_messages": [_vp, _pd, _pd],
    "set_unaries": [_vp, _pd],
    "dual_size": [_vp, _pu64],
    "save_dual": [_vp, _vp, _u64],
    "load_dual": [_vp, _vp, _u64],
    "get_stats": [_vp, C.POINTER(Stats)],
    "synchronize": [_vp],
    "get_trace": [_vp, _pu64, _u64],
    "stream": [_vp, C.POINTER(_vp)],
    "minmarginal_dense": [_i, _u64, _u32, _u32, _pd, _pd, _pd, _i, _pd, _pd],
    "minmarginal_potts": [_i, _u64, _u32, _pd, _pd, _pd, _i, _pd, _pd],
    "minmarginal_trunclin": [_i, _u64, _u32, _pd, _pd, _pd, _pd, _i, _pd, _pd],
    "node_update": [_i, _u64, _u32, _u32, _i, _i, _i, C.POINTER(C.c_uint8), _pd, _pd, _pd, _pd, _pd, _pd, _pu32],
    "unary_min": [_i, _u64, _u32, _pd, _pd, _pu32],
}
_RESTYPE = {"last_error": C.c_char_p}


class LpmpError(RuntimeError):
    """Raised for any non-zero status; the reference throws std::runtime_error in the same situations."""


class CApi:
    """All entry points of include/lpmp_b200.h bound from one shared library."""

    def __init__(self, lib: C.CDLL, prefix: str = "lpmp_"):
        self.lib, self.prefix = lib, prefix
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, prefix + name)  # AttributeError here == a declared symbol is not exported
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, C.c_int)
            setattr(self, name, fn)

    def check(self, status: int, handle=None):
        if status != 0:
            msg = self.last_error(handle).decode() if handle else ""
            raise LpmpError(f"{STATUS.get(status, status)}: {msg}")
