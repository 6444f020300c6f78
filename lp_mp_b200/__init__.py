"""lp_mp_b200 -- B200-native (sm_100a CUDA) implementation of LP_MP's message-passing hot path.

Host side = a thin mirror of the reference's LP<>/Solver/StandardVisitor interface over the C ABI of
``liblpmp_b200.so`` (include/lpmp_b200.h).  There is no CPU fallback: importing works without a GPU, but
creating an ``LP`` loads the CUDA library and raises if it is absent.
"""
from __future__ import annotations

import ctypes
import os

from ._capi import CApi, LpmpError, F32, F64  # noqa: F401

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblpmp_b200.so")
_loaded = None


def library_path() -> str:
    return _LIB_PATH


def load_library() -> CApi:
    """Open lp_mp_b200/liblpmp_b200.so (built by __graft_entry__.build()).  Fails loudly if missing."""
    global _loaded
    if _loaded is None:
        if not os.path.exists(_LIB_PATH):
            raise LpmpError(
                f"CUDA library {_LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback.")
        _loaded = CApi(ctypes.CDLL(_LIB_PATH), "lpmp_")
    return _loaded


from .lp import LP  # noqa: E402
from .solver import LpControl, MpRoundingSolver, Solver, StandardVisitor  # noqa: E402

__all__ = ["LP", "Solver", "MpRoundingSolver", "StandardVisitor", "LpControl", "load_library", "LpmpError"]
