"""ctypes binding of libbsm_b200.so (include/bsm_b200.h). Fails loudly when the CUDA library is
missing: there is no CPU fallback behind this package."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("BSM_B200_LIB", _HERE / "libbsm_b200.so"))

F32, F64, C64 = 0, 1, 2
OP_N, OP_T, OP_C = 0, 1, 2
KIND_BLOCKSPARSE, KIND_SYMMETRIC, KIND_VBCRS = 0, 1, 2
VARIANT_AUTO, VARIANT_GATHER, VARIANT_FUSED, VARIANT_COLOR, VARIANT_FUSED_TMA = 0, 1, 2, 3, 4
DEVICE_NONE = -2

(TAB_ARENA, TAB_BLOCK_OFF, TAB_BLOCK_M, TAB_BLOCK_N, TAB_SET_LEN, TAB_SET_START, TAB_SET_POOL_OFF,
 TAB_POOL, TAB_CONTRIB, TAB_SLICE, TAB_GATHER_ROWS, TAB_GATHER_PTR, TAB_GATHER_POS, TAB_GROUP_PTR,
 TAB_GROUP_SET, TAB_CONTRIB_TOFF, TAB_WCHUNK, TAB_WITEM_PTR, TAB_COLOR_PTR) = range(19)


class Options(ctypes.Structure):
    _fields_ = [("device", c_int32), ("variant", c_int32),
                ("own_row_lo", c_int64), ("own_row_hi", c_int64),
                ("own_col_lo", c_int64), ("own_col_hi", c_int64),
                ("plan_hints", c_int64), ("blocks_on_device", c_int64), ("reserved", c_int64 * 2)]


class CgOptions(ctypes.Structure):
    _fields_ = [("rtol", c_double), ("maxit", c_int64), ("hermitian", c_int32), ("check_every", c_int32)]


class BsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbsm_b200 error {code}: {msg}")
        self.code = code


# every symbol include/bsm_b200.h declares: (name, restype, argtypes)
_P64 = POINTER(c_int64)
SIGNATURES = [
    ("bsm_default_options", None, [POINTER(Options)]),
    ("bsm_create_blocksparse", c_int, [c_int, c_int64, c_int64, c_int64, POINTER(c_void_p), _P64, _P64,
                                       _P64, _P64, _P64, _P64, POINTER(Options), POINTER(c_void_p)]),
    ("bsm_create_symmetric", c_int, [c_int, c_int64, c_int64, c_int64, POINTER(c_void_p), _P64, _P64, _P64,
                                     c_int64, POINTER(c_void_p), _P64, _P64, _P64, _P64, _P64, _P64,
                                     POINTER(Options), POINTER(c_void_p)]),
    ("bsm_create_vbcrs", c_int, [c_int, c_int64, c_int64, c_int64, c_int64, _P64, _P64, _P64,
                                 POINTER(c_void_p), _P64, _P64, POINTER(c_uint8), POINTER(Options),
                                 POINTER(c_void_p)]),
    ("bsm_update_values", c_int, [c_void_p, POINTER(c_void_p), c_int64]),
    ("bsm_update_values_dev", c_int, [c_void_p, POINTER(c_void_p), c_int64]),
    ("bsm_vbcrs_sort_dev", c_int, [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, _P64, c_void_p]),
    ("bsm_destroy", c_int, [c_void_p]),
    ("bsm_mul", c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64,
                        c_int64, c_void_p]),
    ("bsm_mul_host", c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p,
                             c_int64, c_int64]),
    ("bsm_set_variant", c_int, [c_void_p, c_int]),
    ("bsm_set_profiling", c_int, [c_void_p, c_int]),
    ("bsm_get_profile", c_int, [c_void_p, POINTER(c_double), POINTER(c_double)]),
    ("bsm_nnz", c_int64, [c_void_p]),
    ("bsm_stored_entries", c_int64, [c_void_p]),
    ("bsm_size", c_int, [c_void_p, _P64, _P64]),
    ("bsm_dtype_of", c_int, [c_void_p]),
    ("bsm_kind_of", c_int, [c_void_p]),
    ("bsm_work", c_int, [c_void_p, c_int, c_int64, c_int, POINTER(c_double), POINTER(c_double),
                         POINTER(c_double)]),
    ("bsm_launch_count", c_int, [c_void_p, c_int]),
    ("bsm_plan_stats", c_int, [c_void_p, c_int, _P64]),
    ("bsm_table_count", c_int64, [c_void_p, c_int, c_int]),
    ("bsm_table_copy", c_int, [c_void_p, c_int, c_int, c_void_p, c_int64]),
    ("bsm_sparse_build", c_int, [c_void_p, c_int, _P64]),
    ("bsm_sparse_fetch", c_int, [c_void_p, _P64, _P64, c_void_p]),
    ("bsm_sparse_device_pointers", c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), _P64]),
    ("bsm_dist_unique_id", c_int, [c_void_p]),
    ("bsm_dist_init", c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_void_p)]),
    ("bsm_dist_destroy", c_int, [c_void_p]),
    ("bsm_dist_set_overlap", c_int, [c_void_p, c_int]),
    ("bsm_dist_set_collective", c_int, [c_void_p, c_int]),
    ("bsm_dist_set_debug", c_int, [c_void_p, c_int]),
    ("bsm_dist_debug_read", c_int, [c_void_p, _P64]),
    ("bsm_dist_info", c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    ("bsm_dist_allgather_rows", c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64, _P64, c_void_p]),
    ("bsm_dist_allreduce_max_f64", c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    ("bsm_dist_alloc", c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    ("bsm_dist_free", c_int, [c_void_p, c_void_p]),
    ("bsm_mul_dist_peer", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, _P64,
                                  c_void_p]),
    ("bsm_mul_dist_peer_host", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, _P64, c_int64, c_int64, c_void_p]),
    ("bsm_mul_dist", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p,
                             c_int64, c_int64, _P64, c_void_p]),
    ("bsm_dist_allreduce_sum_f64", c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    ("bsm_cg_default_options", None, [POINTER(CgOptions)]),
    ("bsm_cg", c_int, [c_void_p, c_void_p, c_void_p, POINTER(CgOptions), _P64, POINTER(c_double), c_void_p]),
    ("bsm_cg_dist", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, _P64, POINTER(CgOptions), _P64, POINTER(c_double),
                            c_void_p]),
    ("bsm_device_count", c_int, [POINTER(c_int)]),
    ("bsm_malloc", c_int, [c_int, c_size_t, POINTER(c_void_p)]),
    ("bsm_free", c_int, [c_int, c_void_p]),
    ("bsm_memcpy_h2d", c_int, [c_void_p, c_void_p, c_size_t]),
    ("bsm_memcpy_d2h", c_int, [c_void_p, c_void_p, c_size_t]),
    ("bsm_synchronize", c_int, [c_int]),
    ("bsm_last_error", c_char_p, []),
    ("bsm_version", c_char_p, []),
]

_lib = None


def lib() -> ctypes.CDLL:
    """The loaded library. Raises if it has not been built (python blocksparsematrices.jl_b200/build.py)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python blocksparsematrices.jl_b200/build.py` "
                "(nvcc, sm_100a). This package has no CPU fallback.")
        L = ctypes.CDLL(str(LIB_PATH))
        for name, res, args in SIGNATURES:
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise BsmError(rc, (lib().bsm_last_error() or b"").decode(errors="replace"))


def device_count() -> int:
    n = c_int(0)
    rc = lib().bsm_device_count(ctypes.byref(n))
    return int(n.value) if rc == 0 else 0
