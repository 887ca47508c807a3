"""ctypes binding of the C ABI declared in include/vinum_b200.h.

This is the only place where Python touches the native library.  There is NO CPU
fallback: if the shared object is missing the import fails loudly, and every entry
point raises `VinumB200Error` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_C" / "libvinum_b200.so"
if os.environ.get("VINUM_B200_LIB"):  # tuning builds (build.py --variant=...), same C ABI
    LIB_PATH = Path(os.environ["VINUM_B200_LIB"]).resolve()


class VinumB200Error(RuntimeError):
    """A C-ABI call failed (mirrors std::runtime_error -> RuntimeError of the
    reference's pybind11 module, vinum_cpp/src/common/util.h:4-11)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"vinum_b200 [{code}]: {message}")
        self.code = code
        self.message = message


VK_OK = 0
VK_ERR_CUDA, VK_ERR_ARG, VK_ERR_UNSUPPORTED, VK_ERR_OOM, VK_ERR_STATE = -1, -2, -3, -4, -5

# VkDType
I8, I16, I32, I64, U8, U16, U32, U64, F32, F64, BOOL8 = range(1, 12)
# VkCmpOp
EQ, NE, GT, GE, LT, LE = range(6)
# VkMaskOp
MASK_AND, MASK_OR, MASK_NOT = range(3)
# VkPredKind
PRED_NONE, PRED_MASK, PRED_CMP, PRED_EXPR = range(4)
# VkArithOp
ADD, SUB, MUL, DIV, MOD, BITAND, BITOR, BITXOR, NEG, BITNOT = range(10)
# VkAggFunc
AGG_COUNT_STAR, AGG_COUNT, AGG_MIN, AGG_MAX, AGG_SUM, AGG_AVG = range(6)
# VkSortOrder
ASC, DESC = 0, 1
# VkGenKind
GEN_I0, GEN_I1, GEN_I2, GEN_I3, GEN_F0, GEN_F1, GEN_F2, GEN_F3, GEN_K32 = range(9)

AGG_MAX_KEYS = 8
AGG_MAX_FUNCS = 16


class VkColumn(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("validity", C.c_void_p),
        ("offset", C.c_int64),
        ("length", C.c_int64),
        ("dtype", C.c_int32),
        ("nulls_as_nan", C.c_int32),
    ]


class _ScalarValue(C.Union):
    _fields_ = [("i", C.c_int64), ("u", C.c_uint64), ("f", C.c_double)]


class VkScalar(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("_pad", C.c_int32), ("v", _ScalarValue)]


VK_EXPR_MAX_TERMS = 4


class VkExprTerm(C.Structure):
    _fields_ = [("column", VkColumn), ("scalar", VkScalar), ("is_column", C.c_int32), ("op", C.c_int32)]


class VkExprChain(C.Structure):
    _fields_ = [("n_terms", C.c_int32), ("_pad", C.c_int32), ("terms", VkExprTerm * VK_EXPR_MAX_TERMS)]


class VkExprCompare(C.Structure):
    _fields_ = [("lhs", VkExprChain), ("rhs", VkExprChain), ("op", C.c_int32), ("_pad", C.c_int32)]


class VkPredicate(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("op", C.c_int32),
        ("mask", C.c_void_p),
        ("column", VkColumn),
        ("scalar", VkScalar),
        ("expr", C.POINTER(VkExprCompare)),
    ]


def make_scalar(value) -> VkScalar:
    """Python / NumPy scalar -> VkScalar (int -> I64 or U64, float -> F64)."""
    import numpy as np

    s = VkScalar()
    if isinstance(value, (bool, np.bool_)):
        s.dtype = I64
        s.v.i = int(value)
    elif isinstance(value, (int, np.integer)):
        v = int(value)
        if v > 0x7FFFFFFFFFFFFFFF:
            if v > 0xFFFFFFFFFFFFFFFF:
                raise OverflowError("integer literal does not fit in 64 bits")
            s.dtype = U64
            s.v.u = v
        else:
            if v < -0x8000000000000000:
                raise OverflowError("integer literal does not fit in 64 bits")
            s.dtype = I64
            s.v.i = v
    elif isinstance(value, (float, np.floating)):
        s.dtype = F64
        s.v.f = float(value)
    else:
        raise TypeError(f"unsupported scalar type for the device path: {type(value)!r}")
    return s


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: the vinum_b200 CUDA library has not been built. "
            "Run `python vinum_b200/build.py` (needs nvcc); there is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH), mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
    return lib


_lib = _load()

_p = C.c_void_p
_i64 = C.c_int64
_u64 = C.c_uint64
_int = C.c_int
_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/vinum_b200.h declares.
SIGNATURES = {
    "vk_abi_version": (_int, []),
    "vk_last_error": (C.c_char_p, []),
    "vk_device_count": (_int, [C.POINTER(_int)]),
    "vk_set_device": (_int, [_int]),
    "vk_get_device": (_int, [C.POINTER(_int)]),
    "vk_device_info": (_int, [_int, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), C.POINTER(_u64), C.POINTER(_u64)]),
    "vk_malloc": (_int, [_PP, _u64, _p]),
    "vk_free": (_int, [_p, _p]),
    "vk_host_alloc": (_int, [_PP, _u64]),
    "vk_host_free": (_int, [_p]),
    "vk_host_register": (_int, [_p, _u64]),
    "vk_host_unregister": (_int, [_p]),
    "vk_memcpy_h2d": (_int, [_p, _p, _u64, _p]),
    "vk_memcpy_h2d_staged": (_int, [_p, _p, _u64, _p]),
    "vk_memcpy_h2d_auto": (_int, [_p, _p, _u64, _p]),
    "vk_ingest_threads": (_int, []),
    "vk_memcpy_d2h": (_int, [_p, _p, _u64, _p]),
    "vk_memcpy_d2d": (_int, [_p, _p, _u64, _p]),
    "vk_memset": (_int, [_p, _int, _u64, _p]),
    "vk_stream_create": (_int, [_PP]),
    "vk_stream_destroy": (_int, [_p]),
    "vk_stream_sync": (_int, [_p]),
    "vk_device_sync": (_int, []),
    "vk_event_create": (_int, [_PP]),
    "vk_event_destroy": (_int, [_p]),
    "vk_event_record": (_int, [_p, _p]),
    "vk_event_sync": (_int, [_p]),
    "vk_stream_wait_event": (_int, [_p, _p]),
    "vk_event_elapsed_ms": (_int, [_p, _p, C.POINTER(C.c_float)]),
    "vk_launch_count": (_u64, []),
    "vk_set_option": (_int, [C.c_char_p, _i64]),
    "vk_get_option": (_int, [C.c_char_p, C.POINTER(_i64)]),
    "vk_reset_options": (_int, []),
    "vk_datagen": (_int, [_int, _u64, _i64, _i64, _p, _p]),
    "vk_compare_scalar": (_int, [C.POINTER(VkColumn), _int, C.POINTER(VkScalar), _p, _p]),
    "vk_compare_columns": (_int, [C.POINTER(VkColumn), _int, C.POINTER(VkColumn), _p, _p]),
    "vk_between_scalar": (_int, [C.POINTER(VkColumn), C.POINTER(VkScalar), C.POINTER(VkScalar), _int, _p, _p]),
    "vk_isin_scalars": (_int, [C.POINTER(VkColumn), C.POINTER(VkScalar), _int, _int, _p, _p]),
    "vk_mask_combine": (_int, [_int, _p, _p, _i64, _p, _p]),
    "vk_is_null": (_int, [C.POINTER(VkColumn), _int, _p, _p]),
    "vk_mask_to_bits": (_int, [_p, _i64, _p, _p]),
    "vk_bits_to_mask": (_int, [_p, _i64, _i64, _p, _p]),
    "vk_filter_scratch_bytes": (_u64, [_i64]),
    "vk_filter": (_int, [C.POINTER(VkPredicate), _i64, C.POINTER(VkColumn), _int, _PP, _PP, _p, _p, _p]),
    "vk_arith": (_int, [_int, C.POINTER(VkColumn), C.POINTER(VkScalar), C.POINTER(VkColumn), C.POINTER(VkScalar),
                        _i64, _int, _p, _p]),
    "vk_expr_eval": (_int, [C.POINTER(VkExprChain), _i64, _p, C.POINTER(C.c_int32), _p]),
    "vk_expr_compare": (_int, [C.POINTER(VkExprCompare), _i64, _p, _p]),
    "vk_agg_create": (_int, [_PP, _int, C.POINTER(C.c_int32), _int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i64]),
    "vk_agg_destroy": (_int, [_p]),
    "vk_agg_update": (_int, [_p, C.POINTER(VkPredicate), _i64, C.POINTER(VkColumn), C.POINTER(VkColumn), _p]),
    "vk_agg_num_groups": (_int, [_p, C.POINTER(_i64), _p]),
    "vk_agg_result": (_int, [_p, _i64, _PP, _PP, _p, _PP, _PP, _PP, _p]),
    "vk_agg_record_words": (_int, [_p, C.POINTER(_int)]),
    "vk_agg_partition_counts": (_int, [_p, _int, _p, _p]),
    "vk_agg_export_partials": (_int, [_p, _int, _p, _p, _p]),
    "vk_agg_merge_partials": (_int, [_p, _p, _i64, _p]),
    "vk_agg_result_packed_bytes": (_u64, [_int, _int, _i64]),
    "vk_agg_result_packed": (_int, [_p, _i64, _p, _p]),
    "vk_peer_create": (_int, [_PP, _int, _int, _i64, _int]),
    "vk_peer_destroy": (_int, [_p]),
    "vk_peer_handle": (_int, [_p, _p]),
    "vk_peer_open": (_int, [_p, _int, _p]),
    "vk_peer_attach_local": (_int, [_p, _int, _p]),
    "vk_peer_decision_ptr": (_p, [_p, _u64]),
    "vk_agg_peer_send": (_int, [_p, _p, _int, _u64, _p]),
    "vk_agg_peer_merge": (_int, [_p, _p, _u64, _p]),
    "vk_agg_last_path": (_int, [_p]),
    "vk_agg_estimate_groups": (C.c_double, [C.c_double, C.c_double]),
    "vk_agg_profile": (_int, [_p, _int]),
    "vk_agg_profile_read": (_int, [_p, _int, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(_i64)]),
    "vk_sort_scratch_bytes": (_u64, [_i64]),
    "vk_sort_indices": (_int, [C.POINTER(VkColumn), C.POINTER(C.c_int32), _int, _i64, _p, _p, _p]),
    "vk_sort_indices_keys": (_int, [C.POINTER(VkColumn), C.POINTER(C.c_int32), _int, _i64, _p, _p, _p, _p]),
    "vk_take": (_int, [C.POINTER(VkColumn), _p, _i64, _p, _p, _p]),
    "vk_topk_scratch_bytes": (_u64, []),
    "vk_topk_candidates": (_int, [C.POINTER(VkColumn), C.c_int32, _i64, _i64, _i64, _p, C.POINTER(_i64), _p, _p]),
}

_STATUS_FUNCS = set()
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(_lib, _name)  # AttributeError here == header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args
    if _res is _int and _name not in ("vk_abi_version", "vk_agg_last_path", "vk_ingest_threads"):
        _STATUS_FUNCS.add(_name)


def last_error() -> str:
    msg = _lib.vk_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    if rc != VK_OK:
        raise VinumB200Error(rc, last_error())


class _Checked:
    """`lib.vk_xxx(...)` raises on failure; the raw CDLL is `lib.raw`."""

    def __init__(self, raw):
        self.raw = raw

    def __getattr__(self, name):
        fn = getattr(self.raw, name)
        if name in _STATUS_FUNCS:
            def wrapped(*args, _fn=fn):
                check(_fn(*args))
            wrapped.__name__ = name
            setattr(self, name, wrapped)
            return wrapped
        setattr(self, name, fn)
        return fn


lib = _Checked(_lib)

if _lib.vk_abi_version() != 1:
    raise ImportError("libvinum_b200.so ABI version mismatch; rebuild with `python vinum_b200/build.py --force`")
