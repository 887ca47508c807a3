"""Device operators: thin, typed Python wrappers over the C ABI.

Each function cites the reference operator it replaces (paths relative to the
reference checkout).  Inputs and outputs are `DeviceColumn`s resident in HBM; no
function here computes on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple, List, Optional, Sequence, Union

import numpy as np
import pyarrow as pa

from . import _lib as L
from ._lib import lib, VinumB200Error
from .device import (DeviceBuffer, DeviceColumn, DeviceBatch, Stream, VK_SIZE, VK_TO_NUMPY, default_stream,
                     arrow_from_numpy, vk_dtype_of_numpy)

Scalar = Union[int, float, bool, np.generic]

CMP_OPS = {"==": L.EQ, "=": L.EQ, "!=": L.NE, "<>": L.NE, ">": L.GT, ">=": L.GE, "<": L.LT, "<=": L.LE}
ARITH_OPS = {"+": L.ADD, "-": L.SUB, "*": L.MUL, "/": L.DIV, "%": L.MOD, "&": L.BITAND, "|": L.BITOR,
             "#": L.BITXOR, "neg": L.NEG, "~": L.BITNOT}
_NUMPY_UFUNC = {L.ADD: np.add, L.SUB: np.subtract, L.MUL: np.multiply, L.DIV: np.divide, L.MOD: np.mod,
                L.BITAND: np.bitwise_and, L.BITOR: np.bitwise_or, L.BITXOR: np.bitwise_xor,
                L.NEG: np.negative, L.BITNOT: np.invert}


def _mask_column(n: int, stream: Optional[Stream]) -> DeviceColumn:
    return DeviceColumn.empty(n, L.BOOL8, pa.bool_(), stream)


def _is_col(x) -> bool:
    return isinstance(x, DeviceColumn)


# ------------------------------------------------------------------ compare ----
def compare(lhs: DeviceColumn, op: str, rhs: Union[DeviceColumn, Scalar], stream: Optional[Stream] = None) -> DeviceColumn:
    """`lhs <op> rhs` -> byte mask.  Replaces the NumPy comparison lambdas of
    vinum/core/expressions.py:30-36, including their NULL behaviour: a column with
    NULLs is compared as float with NaN (vinum/arrow/record_batch.py:100-125)."""
    st = stream or default_stream()
    code = CMP_OPS[op]
    out = _mask_column(lhs.length, st)
    a = lhs.vk(nulls_as_nan=True)
    if _is_col(rhs):
        b = rhs.vk(nulls_as_nan=True)
        lib.vk_compare_columns(C.byref(a), code, C.byref(b), C.c_void_p(out.data_ptr), st.ptr)
    else:
        s = L.make_scalar(rhs)
        lib.vk_compare_scalar(C.byref(a), code, C.byref(s), C.c_void_p(out.data_ptr), st.ptr)
    return out


def between(x: DeviceColumn, low: Scalar, high: Scalar, negate: bool = False,
            stream: Optional[Stream] = None) -> DeviceColumn:
    """BETWEEN / NOT BETWEEN, vinum/core/expressions.py:43-48."""
    st = stream or default_stream()
    out = _mask_column(x.length, st)
    a = x.vk(nulls_as_nan=True)
    lo, hi = L.make_scalar(low), L.make_scalar(high)
    lib.vk_between_scalar(C.byref(a), C.byref(lo), C.byref(hi), int(negate), C.c_void_p(out.data_ptr), st.ptr)
    return out


def isin(x: DeviceColumn, values: Sequence[Scalar], negate: bool = False, stream: Optional[Stream] = None) -> DeviceColumn:
    """IN / NOT IN over a literal list (np.isin), vinum/core/expressions.py:39-40."""
    st = stream or default_stream()
    out = _mask_column(x.length, st)
    a = x.vk(nulls_as_nan=True)
    arr = (L.VkScalar * max(len(values), 1))()
    for i, v in enumerate(values):
        arr[i] = L.make_scalar(v)
    lib.vk_isin_scalars(C.byref(a), arr, len(values), int(negate), C.c_void_p(out.data_ptr), st.ptr)
    return out


# --------------------------------------------------------------- mask algebra ----
def mask_and(a: DeviceColumn, b: DeviceColumn, stream: Optional[Stream] = None) -> DeviceColumn:
    """pc.and_, vinum/core/expressions.py:27."""
    return _mask_binary(L.MASK_AND, a, b, stream)


def mask_or(a: DeviceColumn, b: DeviceColumn, stream: Optional[Stream] = None) -> DeviceColumn:
    """pc.or_, vinum/core/expressions.py:28."""
    return _mask_binary(L.MASK_OR, a, b, stream)


def mask_not(a: DeviceColumn, stream: Optional[Stream] = None) -> DeviceColumn:
    """pc.invert, vinum/core/expressions.py:29."""
    return _mask_binary(L.MASK_NOT, a, None, stream)


def _mask_binary(op: int, a: DeviceColumn, b: Optional[DeviceColumn], stream: Optional[Stream]) -> DeviceColumn:
    st = stream or default_stream()
    if a.dtype != L.BOOL8 or (b is not None and b.dtype != L.BOOL8):
        raise TypeError("boolean operators need boolean operands")
    if b is not None and a.length != b.length:
        raise ValueError("mask length mismatch")
    if a.has_nulls or (b is not None and b.has_nulls):
        raise VinumB200Error(L.VK_ERR_UNSUPPORTED, "Kleene logic over NULL booleans is not on the device path")
    out = _mask_column(a.length, st)
    pa_ = C.c_void_p(a.data_ptr + a.offset)
    pb_ = C.c_void_p(b.data_ptr + b.offset) if b is not None else None
    lib.vk_mask_combine(op, pa_, pb_, a.length, C.c_void_p(out.data_ptr), st.ptr)
    return out


def is_null(x: DeviceColumn, stream: Optional[Stream] = None) -> DeviceColumn:
    """pc.is_null, vinum/core/expressions.py:37."""
    return _is_null(x, False, stream)


def is_valid(x: DeviceColumn, stream: Optional[Stream] = None) -> DeviceColumn:
    """pc.is_valid, vinum/core/expressions.py:38."""
    return _is_null(x, True, stream)


def _is_null(x: DeviceColumn, want_valid: bool, stream: Optional[Stream]) -> DeviceColumn:
    st = stream or default_stream()
    out = _mask_column(x.length, st)
    v = x.vk()
    lib.vk_is_null(C.byref(v), int(want_valid), C.c_void_p(out.data_ptr), st.ptr)
    return out


# ------------------------------------------------------------------- filter ----
class Predicate:
    """WHERE predicate handed to `filter_batch` / `Aggregator.update`: either a byte
    mask or a `column <op> scalar` comparison that the kernel evaluates in registers."""

    def __init__(self, mask: Optional[DeviceColumn] = None, column: Optional[DeviceColumn] = None,
                 op: Optional[str] = None, scalar: Optional[Scalar] = None, chains=None):
        self.mask, self.column, self.op, self.scalar, self.chains = mask, column, op, scalar, chains
        if mask is None and column is None and chains is None:
            raise ValueError("Predicate needs a mask or a comparison")
        self._expr = None

    @staticmethod
    def expr(lhs: "Chain", op: str, rhs: "Chain") -> "Predicate":
        """`<chain> <op> <chain>` fused into the consumer kernel (VK_PRED_EXPR): `WHERE a * 10 > b` reaches
        vk_filter / vk_agg_update as an expression, no intermediate column or mask is written."""
        return Predicate(op=op, chains=(lhs, rhs))

    @staticmethod
    def compare(column: DeviceColumn, op: str, scalar: Scalar) -> "Predicate":
        return Predicate(column=column, op=op, scalar=scalar)

    @staticmethod
    def from_mask(mask: DeviceColumn) -> "Predicate":
        return Predicate(mask=mask)

    def vk(self) -> L.VkPredicate:
        p = L.VkPredicate()
        if self.chains is not None:
            self._expr = _vk_compare(self.chains[0], self.op, self.chains[1])   # kept alive with the predicate
            p.kind = L.PRED_EXPR
            p.expr = C.pointer(self._expr)
            return p
        if self.mask is not None:
            if self.mask.dtype != L.BOOL8:
                raise TypeError("filter mask must be boolean")
            p.kind = L.PRED_MASK
            p.mask = self.mask.data_ptr + self.mask.offset
        else:
            p.kind = L.PRED_CMP
            p.op = CMP_OPS[self.op]
            p.column = self.column.vk(nulls_as_nan=True)
            p.scalar = L.make_scalar(self.scalar)
        return p

    @property
    def length(self) -> int:
        if self.chains is not None:
            return chain_length(self.chains[0]) or chain_length(self.chains[1])
        return self.mask.length if self.mask is not None else self.column.length


# ------------------------------------------------------- expression chains ----
# A chain is a list of (op, term): [(None, a), ("*", 10), ("+", b)] means (a * 10) + b, evaluated left to
# right in ONE pass (vk_expr.cuh).  Terms: null-free int64 / float64 DeviceColumns, Python ints / floats.
Chain = List[Tuple[Optional[str], Union[DeviceColumn, Scalar]]]
CHAIN_OPS = ("+", "-", "*", "/", "%", "&", "|", "#")
MAX_CHAIN_TERMS = L.VK_EXPR_MAX_TERMS


def chain_term_ok(x) -> bool:
    if _is_col(x):
        return x.dtype in (L.I64, L.F64) and not x.has_nulls
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, (bool, np.bool_)) and \
        (not isinstance(x, (int, np.integer)) or -(1 << 63) <= int(x) < (1 << 63))


def chain_length(chain: Chain) -> int:
    for _, t in chain:
        if _is_col(t):
            return t.length
    return 0


def _vk_chain(chain: Chain) -> L.VkExprChain:
    if not 1 <= len(chain) <= MAX_CHAIN_TERMS:
        raise ValueError("an expression chain has 1..4 terms")
    c = L.VkExprChain()
    c.n_terms = len(chain)
    for i, (op, term) in enumerate(chain):
        t = c.terms[i]
        if _is_col(term):
            t.is_column = 1
            t.column = term.vk()
        else:
            t.is_column = 0
            t.scalar = L.make_scalar(float(term) if isinstance(term, (float, np.floating)) else int(term))
        t.op = ARITH_OPS[op] if i else 0
    return c


def _vk_compare(lhs: Chain, op: str, rhs: Chain) -> L.VkExprCompare:
    e = L.VkExprCompare()
    e.lhs = _vk_chain(lhs)
    e.rhs = _vk_chain(rhs)
    e.op = CMP_OPS[op]
    return e


def eval_chain(chain: Chain, stream: Optional[Stream] = None) -> DeviceColumn:
    """The whole arithmetic chain in one pass (vk_expr_eval); the result dtype follows NumPy's promotion."""
    st = stream or default_stream()
    n = chain_length(chain)
    c = _vk_chain(chain)
    out = DeviceBuffer(max(n, 1) * 8, st)
    dt = C.c_int32()
    lib.vk_expr_eval(C.byref(c), n, C.c_void_p(out.ptr), C.byref(dt), st.ptr)
    return DeviceColumn(out, None, 0, n, int(dt.value), pa.float64() if dt.value == L.F64 else pa.int64())


def compare_chains(lhs: Chain, op: str, rhs: Chain, stream: Optional[Stream] = None) -> DeviceColumn:
    """`<chain> <op> <chain>` -> byte mask in one pass (vk_expr_compare)."""
    st = stream or default_stream()
    n = chain_length(lhs) or chain_length(rhs)
    out = _mask_column(n, st)
    e = _vk_compare(lhs, op, rhs)
    lib.vk_expr_compare(C.byref(e), n, C.c_void_p(out.data_ptr), st.ptr)
    return out


def filter_batch(batch: DeviceBatch, pred: Predicate, stream: Optional[Stream] = None) -> DeviceBatch:
    """Order-preserving compaction of EVERY column of the batch.  Replaces
    FilterOperator._kernel -> RecordBatch.filter -> pa.RecordBatch.filter
    (vinum/core/algebra.py:119-123, vinum/arrow/record_batch.py:85-90)."""
    st = stream or default_stream()
    n = batch.num_rows
    if pred.length != n:
        raise ValueError("predicate length != batch length")
    ncols = len(batch.columns)
    vp = pred.vk()
    vcols = (L.VkColumn * max(ncols, 1))()
    out_data = (C.c_void_p * max(ncols, 1))()
    out_valid = (C.c_void_p * max(ncols, 1))()
    outs: List[DeviceColumn] = []
    valid_bytes: List[Optional[DeviceBuffer]] = []
    for i, c in enumerate(batch.columns):
        vcols[i] = c.vk()
        o = DeviceColumn.empty(n, c.dtype, c.arrow_type, st)
        outs.append(o)
        out_data[i] = o.data_ptr
        vb = DeviceBuffer(max(n, 1), st) if c.has_nulls else None
        valid_bytes.append(vb)
        out_valid[i] = vb.ptr if vb is not None else None
    scratch = DeviceBuffer(lib.vk_filter_scratch_bytes(n), st)
    out_rows = DeviceBuffer(8, st)
    lib.vk_filter(C.byref(vp), n, vcols, ncols, out_data, out_valid, C.c_void_p(out_rows.ptr),
                  C.c_void_p(scratch.ptr), st.ptr)
    m = int(out_rows.to_numpy(np.int64, 1, st)[0])  # synchronises: the row count sizes the result
    result_cols = []
    for c, o, vb in zip(batch.columns, outs, valid_bytes):
        validity = None
        null_count = 0
        if vb is not None and m:
            bits = DeviceBuffer((m + 7) // 8 + 8, st)
            lib.vk_mask_to_bits(C.c_void_p(vb.ptr), m, C.c_void_p(bits.ptr), st.ptr)
            validity = bits
            null_count = -1  # unknown without a count; treated as "has nulls"
        col = DeviceColumn(o.data, validity, 0, m, c.dtype, c.arrow_type, null_count)
        result_cols.append(col)
    return DeviceBatch(result_cols, batch.column_names, m)


# --------------------------------------------------------------- arithmetic ----
def _numpy_view_dtype(x) -> np.dtype:
    """dtype of the NumPy view the reference would compute on (NULLs -> float)."""
    if _is_col(x):
        dt = np.dtype(VK_TO_NUMPY[x.dtype])
        if x.has_nulls and dt.kind in "iub":
            return np.dtype(np.float64)
        return dt
    return None


def result_dtype(op: int, lhs, rhs) -> np.dtype:
    """NumPy's result dtype for `ufunc(lhs, rhs)` with columns as arrays and literals as
    Python scalars (weak promotion) -- asked of NumPy itself on empty operands."""
    def operand(x):
        if _is_col(x):
            return np.empty(0, dtype=_numpy_view_dtype(x))
        return x
    uf = _NUMPY_UFUNC[op]
    with np.errstate(all="ignore"):
        if op in (L.NEG, L.BITNOT):
            return uf(operand(lhs)).dtype
        return uf(operand(lhs), operand(rhs)).dtype


def arith(op: str, lhs: Union[DeviceColumn, Scalar], rhs: Union[DeviceColumn, Scalar, None] = None,
          stream: Optional[Stream] = None) -> DeviceColumn:
    """Element-wise arithmetic with NumPy semantics.  Replaces the ufuncs of
    vinum/core/expressions.py:13-24 (np.add/subtract/multiply/divide/mod/negative/
    bitwise_*): `/` is true division, `%` floor-mod, integers wrap, NULL -> NaN."""
    st = stream or default_stream()
    code = ARITH_OPS[op]
    if not _is_col(lhs) and not _is_col(rhs):
        raise TypeError("arith needs at least one column operand")
    n = lhs.length if _is_col(lhs) else rhs.length
    if _is_col(lhs) and _is_col(rhs) and lhs.length != rhs.length:
        raise ValueError("operand length mismatch")
    out_np = result_dtype(code, lhs, rhs)
    if out_np.kind not in "iufb":
        raise VinumB200Error(L.VK_ERR_UNSUPPORTED, f"result dtype {out_np} is not on the device path")
    out_dt = vk_dtype_of_numpy(out_np)
    out = DeviceColumn.empty(n, out_dt, None, st)

    def side(x):
        if x is None:
            return None, None
        if _is_col(x):
            v = x.vk(nulls_as_nan=True)
            return C.byref(v), None
        s = L.make_scalar(x)
        return None, C.byref(s)

    lc, ls = side(lhs)
    rc, rs = side(rhs)
    lib.vk_arith(code, lc, ls, rc, rs, n, out_dt, C.c_void_p(out.data_ptr), st.ptr)
    return out


# ------------------------------------------------------------------- gather ----
def take(col: DeviceColumn, indices: DeviceColumn, stream: Optional[Stream] = None) -> DeviceColumn:
    """out[i] = col[indices[i]] (arrow::compute::Take, sort.cpp:40)."""
    st = stream or default_stream()
    if indices.dtype != L.I64:
        raise TypeError("take indices must be int64")
    n = indices.length
    out = DeviceColumn.empty(n, col.dtype, col.arrow_type, st)
    vb = DeviceBuffer(max(n, 1), st) if col.has_nulls else None
    v = col.vk()
    lib.vk_take(C.byref(v), C.c_void_p(indices.data_ptr + indices.offset * 8), n, C.c_void_p(out.data_ptr),
                C.c_void_p(vb.ptr) if vb is not None else None, st.ptr)
    if vb is not None and n:
        bits = DeviceBuffer((n + 7) // 8 + 8, st)
        lib.vk_mask_to_bits(C.c_void_p(vb.ptr), n, C.c_void_p(bits.ptr), st.ptr)
        out = DeviceColumn(out.data, bits, 0, n, col.dtype, col.arrow_type, -1)
    return out


def sort_indices(keys: Sequence[DeviceColumn], orders: Sequence[int], stream: Optional[Stream] = None) -> DeviceColumn:
    """Stable multi-key sort permutation (arrow::compute::SortIndices, sort.cpp:33):
    NaN after all numbers and NULL last in both directions."""
    st = stream or default_stream()
    if not keys:
        raise ValueError("at least one sort key is required")
    n = keys[0].length
    out = DeviceColumn.empty(n, L.I64, pa.int64(), st)
    if n == 0:
        return out
    vk = (L.VkColumn * len(keys))()
    for i, k in enumerate(keys):
        vk[i] = k.vk()
    ords = (C.c_int32 * len(keys))(*[int(o) for o in orders])
    scratch = DeviceBuffer(lib.vk_sort_scratch_bytes(n), st)
    lib.vk_sort_indices(vk, ords, len(keys), n, C.c_void_p(out.data_ptr), C.c_void_p(scratch.ptr), st.ptr)
    return out


def sort_indices_keys(keys: Sequence[DeviceColumn], orders: Sequence[int],
                      stream: Optional[Stream] = None) -> Tuple[DeviceColumn, Optional[DeviceColumn]]:
    """`sort_indices` that also returns keys[0] in sorted order when the last radix pass can write it
    (plain 8-byte column without NULLs: vk_sort_indices_keys) -- Sort::Sorted gathers every column, the
    sort key included (sort.cpp:40-48), and this saves that column's random gather.  (permutation, None)
    when keys[0] does not qualify."""
    st = stream or default_stream()
    k0 = keys[0]
    n = k0.length
    if n == 0 or k0.has_nulls or k0.dtype not in (L.F64, L.I64, L.U64) or (k0.data_ptr + k0.offset * 8) % 8:
        return sort_indices(keys, orders, st), None
    out = DeviceColumn.empty(n, L.I64, pa.int64(), st)
    sorted0 = DeviceColumn.empty(n, k0.dtype, k0.arrow_type, st)
    vk = (L.VkColumn * len(keys))()
    for i, k in enumerate(keys):
        vk[i] = k.vk()
    ords = (C.c_int32 * len(keys))(*[int(o) for o in orders])
    scratch = DeviceBuffer(lib.vk_sort_scratch_bytes(n), st)
    lib.vk_sort_indices_keys(vk, ords, len(keys), n, C.c_void_p(out.data_ptr), C.c_void_p(sorted0.data_ptr),
                             C.c_void_p(scratch.ptr), st.ptr)
    return out, sorted0


def topk_candidates(key: DeviceColumn, order: int, k: int, stream: Optional[Stream] = None) -> Optional[DeviceColumn]:
    """Row ids (int64, unordered) of a superset of the first `k` rows of `ORDER BY key <order>`, or
    None when selecting does not pay and the caller should sort every row (vk_topk_candidates)."""
    st = stream or default_stream()
    n = key.length
    cap = max(1 << 16, 4 * k, n // 4)
    out = DeviceColumn.empty(cap, L.I64, pa.int64(), st)
    scratch = DeviceBuffer(lib.vk_topk_scratch_bytes(), st)
    count = C.c_int64(-1)
    v = key.vk()
    lib.vk_topk_candidates(C.byref(v), int(order), n, int(k), cap, C.c_void_p(out.data_ptr), C.byref(count),
                           C.c_void_p(scratch.ptr), st.ptr)
    if count.value < 0:
        return None
    return out.slice(0, count.value)


def sort_top(keys: Sequence[DeviceColumn], orders: Sequence[int], k: int, stream: Optional[Stream] = None) -> DeviceColumn:
    """The first `k` entries of sort_indices(keys, orders) -- `ORDER BY ... LIMIT k` -- through a
    radix select on the first key when that beats sorting every row.  Candidates are put back
    into row order before they are sorted, so ties keep their input order exactly as in the
    full stable sort."""
    st = stream or default_stream()
    n = keys[0].length
    k = min(k, n)
    cand = topk_candidates(keys[0], orders[0], k, st) if 0 < k < n else None
    if cand is None:
        return sort_indices(keys, orders, st).slice(0, k)
    ids = take(cand, sort_indices([cand], [L.ASC], st), st)
    perm = sort_indices([take(key, ids, st) for key in keys], orders, st).slice(0, k)
    return take(ids, perm, st)


def sort_batch(batch: DeviceBatch, key_names: Sequence[str], orders: Sequence[int],
               stream: Optional[Stream] = None) -> DeviceBatch:
    """SortIndices + Take of every column (Sort::Sorted, sort.cpp:15-63)."""
    st = stream or default_stream()
    keys = [batch.column(k) for k in key_names]
    idx, sorted0 = sort_indices_keys(keys, orders, st)
    cols = [sorted0 if (sorted0 is not None and c is keys[0]) else take(c, idx, st) for c in batch.columns]
    return DeviceBatch(cols, batch.column_names, batch.num_rows)
