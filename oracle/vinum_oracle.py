"""CPU restatement of the reference's hot-path algorithms -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module, and only as the CHECKER; the product path (vinum_b200/)
never does and fails loudly when its CUDA library is missing.

Every function cites the reference code it restates (paths relative to the reference
checkout, dmitrykoval/vinum @ c36b351).  Pinning: tests/test_oracle_golden.py checks
this module against
  * the reference's own golden vectors (vinum_cpp/test/hash_agg_test.cpp fixtures,
    restated in tests/golden/gtest_vectors.py),
  * outputs of the reference's own C++ operators compiled here (oracle/_ref, built by
    oracle/build_ref.sh) committed as fixtures under tests/golden/*.arrow by
    oracle/gen_golden.py, and live against oracle/_ref when it is present.

Two aggregate implementations are kept on purpose: `hash_aggregate_rowwise` is a
literal row-at-a-time transcription of the C++ control flow (small inputs), and
`hash_aggregate` is the vectorised NumPy equivalent used at test sizes of 1e5-1e7.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa

# AggFuncType, vinum_cpp/src/operators/aggregate/agg_funcs.h:16-18 (public subset,
# vinum/core/vinum_lib.cpp:25-32)
COUNT_STAR, COUNT, MIN, MAX, SUM, AVG = "COUNT_STAR", "COUNT", "MIN", "MAX", "SUM", "AVG"
ASC, DESC = "ASC", "DESC"

_U64 = (1 << 64) - 1
_I64_MAX = (1 << 63) - 1


# =============================================================== L0: NumPy views
def np_view(arr: pa.Array) -> np.ndarray:
    """RecordBatch._arrow_array_to_numpy, vinum/arrow/record_batch.py:100-125:
    zero-copy for null-free numeric columns, otherwise a copy in which NULL becomes NaN
    (integers are promoted to float64)."""
    if isinstance(arr, pa.ChunkedArray):
        return arr.to_numpy()
    try:
        return arr.to_numpy(zero_copy_only=True)
    except pa.ArrowInvalid:
        return arr.to_numpy(zero_copy_only=False)


# ============================================================ a1: comparisons
_CMP = {
    "==": lambda x, y: x == y, "!=": lambda x, y: x != y, ">": lambda x, y: x > y,
    ">=": lambda x, y: x >= y, "<": lambda x, y: x < y, "<=": lambda x, y: x <= y,
}


def compare(x, op: str, y) -> np.ndarray:
    """Comparison lambdas, vinum/core/expressions.py:30-36 (operands are NumPy views
    or Python literals; result is a NumPy bool array)."""
    xv = np_view(x) if isinstance(x, (pa.Array, pa.ChunkedArray)) else x
    yv = np_view(y) if isinstance(y, (pa.Array, pa.ChunkedArray)) else y
    with np.errstate(invalid="ignore"):
        return np.asarray(_CMP[op](xv, yv))


def between(x, low, high, negate: bool = False) -> np.ndarray:
    """BETWEEN / NOT BETWEEN, vinum/core/expressions.py:43-48."""
    xv = np_view(x) if isinstance(x, (pa.Array, pa.ChunkedArray)) else x
    with np.errstate(invalid="ignore"):
        if negate:
            return np.logical_or(xv < low, xv > high)
        return np.logical_and(xv >= low, xv <= high)


def isin(x, values, negate: bool = False) -> np.ndarray:
    """IN / NOT IN, vinum/core/expressions.py:39-40 (np.isin)."""
    xv = np_view(x) if isinstance(x, (pa.Array, pa.ChunkedArray)) else x
    return np.isin(xv, list(values), invert=negate)


# ========================================================= a2: boolean algebra
def mask_and(a, b):
    """pc.and_, vinum/core/expressions.py:27 (null-free masks on the numeric path)."""
    return np.logical_and(np.asarray(a, dtype=bool), np.asarray(b, dtype=bool))


def mask_or(a, b):
    """pc.or_, vinum/core/expressions.py:28."""
    return np.logical_or(np.asarray(a, dtype=bool), np.asarray(b, dtype=bool))


def mask_not(a):
    """pc.invert, vinum/core/expressions.py:29."""
    return np.logical_not(np.asarray(a, dtype=bool))


def is_null(x: pa.Array) -> np.ndarray:
    """pc.is_null, vinum/core/expressions.py:37."""
    return ~is_valid(x)


def is_valid(x: pa.Array) -> np.ndarray:
    """pc.is_valid, vinum/core/expressions.py:38."""
    if x.null_count == 0:
        return np.ones(len(x), dtype=bool)
    return np.asarray(x.is_valid().to_numpy(zero_copy_only=False), dtype=bool)


# ================================================================== a4: filter
def filter_batch(batch: pa.RecordBatch, mask: np.ndarray) -> pa.RecordBatch:
    """FilterOperator._kernel -> RecordBatch.filter (vinum/core/algebra.py:119-123,
    vinum/arrow/record_batch.py:85-90): keep rows whose mask is true, input order
    preserved, every column compacted.  Restated as index selection (`take` of the
    selected row ids) so that it does not lean on the same Arrow kernel."""
    idx = np.flatnonzero(np.asarray(mask, dtype=bool))
    return batch.take(pa.array(idx, type=pa.int64()))


# ============================================================== a5: arithmetic
_ARITH = {
    "+": np.add, "-": np.subtract, "*": np.multiply, "/": np.divide, "%": np.mod,
    "&": np.bitwise_and, "|": np.bitwise_or, "#": np.bitwise_xor,
}


def arith(op: str, x, y=None) -> np.ndarray:
    """NumPy ufuncs of vinum/core/expressions.py:13-24 on NumPy views."""
    xv = np_view(x) if isinstance(x, (pa.Array, pa.ChunkedArray)) else x
    with np.errstate(all="ignore"):
        if op == "neg":
            return np.negative(xv)
        if op == "~":
            return ~xv
        yv = np_view(y) if isinstance(y, (pa.Array, pa.ChunkedArray)) else y
        return _ARITH[op](xv, yv)


# ======================================================== a9-a14: hash aggregate
def _physical(arr: pa.Array) -> Tuple[np.ndarray, np.ndarray]:
    """(values in the physical dtype, validity) of a fixed-width Arrow array --
    what NumericArrayIter walks (vinum_cpp/src/common/array_iterators.h:173-228)."""
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks()
    t = arr.type
    n = len(arr)
    valid = np.ones(n, dtype=bool) if arr.null_count == 0 else np.asarray(
        arr.is_valid().to_numpy(zero_copy_only=False), dtype=bool)
    if pa.types.is_boolean(t):
        vals = np.asarray(arr.fill_null(False).to_numpy(zero_copy_only=False), dtype=bool)
        return vals, valid
    if pa.types.is_date32(t) or pa.types.is_time32(t):
        phys = pa.int32()
    elif pa.types.is_date64(t) or pa.types.is_time64(t) or pa.types.is_timestamp(t) or pa.types.is_duration(t):
        phys = pa.int64()
    else:
        phys = t
    if phys != t:
        arr = arr.view(phys)
    npdt = phys.to_pandas_dtype()
    bufs = arr.buffers()
    vals = np.frombuffer(bufs[1], dtype=npdt, count=arr.offset + n)[arr.offset:] if n else np.empty(0, dtype=npdt)
    return vals, valid


def key_as_uint64(vals: np.ndarray) -> np.ndarray:
    """NextAsUInt64: integers `static_cast<uint64_t>` (sign-extending,
    array_iterators.h:215-217); floats are bit-cast (array_iterators.h:239-248)."""
    if vals.dtype == np.float64:
        return vals.view(np.uint64).copy()
    if vals.dtype == np.float32:
        return vals.view(np.uint32).astype(np.uint64)
    if vals.dtype.kind == "i":
        return vals.astype(np.int64).view(np.uint64)
    return vals.astype(np.uint64)


def _sum_output_type(t: pa.DataType) -> pa.DataType:
    """agg_func_factory.cpp:107-175."""
    if pa.types.is_signed_integer(t):
        return pa.int64()
    if pa.types.is_unsigned_integer(t):
        return pa.uint64()
    if pa.types.is_floating(t):
        return pa.float64()
    if pa.types.is_time(t) or pa.types.is_duration(t):
        return t
    raise RuntimeError("Column data type is not supported by sum().")


def _avg_output_type(t: pa.DataType) -> pa.DataType:
    """agg_func_factory.cpp:176-246."""
    if t in (pa.int8(), pa.int16(), pa.uint8(), pa.uint16()):
        return pa.float32()
    if pa.types.is_integer(t) or pa.types.is_floating(t) or pa.types.is_time(t) or pa.types.is_duration(t):
        return pa.float64()
    raise RuntimeError("Column data type is not supported by avg().")


def hugeint_try_cast_int64(v: int) -> Optional[int]:
    """Hugeint::TryCast<int64_t>, vinum_cpp/src/common/huge_int.cpp:341-361.  By the
    code, a value of exactly -2^63 does NOT fit (the negative branch needs
    lower > 2^63)."""
    lower, upper = v & _U64, v >> 64
    if upper == 0 and lower <= _I64_MAX:
        return lower
    if upper == -1 and lower > _U64 - _I64_MAX:
        return -(_U64 - lower + 1)
    return None


def hugeint_try_cast_uint64(v: int) -> Optional[int]:
    """Hugeint::TryCast<uint64_t>, huge_int.cpp:341-361,383-385."""
    lower, upper = v & _U64, v >> 64
    if upper == 0:
        return lower
    return None if not (upper == -1 and lower > 0) else (-(_U64 - lower + 1)) & _U64


def hugeint_to_double(v: int) -> float:
    """Hugeint::TryCast<double>, huge_int.cpp:394-406."""
    lower, upper = v & _U64, v >> 64
    if upper == -1:
        return -float(np.float64(_U64 - lower)) - 1.0
    return float(np.float64(lower) + np.float64(upper) * np.float64(_U64))


def _trunc_divmod(a: int, b: int) -> Tuple[int, int]:
    """Hugeint::DivMod, huge_int.cpp:218-265: truncating division, remainder takes the
    sign of the dividend."""
    q = abs(a) // abs(b)
    if (a < 0) != (b < 0):
        q = -q
    return q, a - q * b


def avg_hugeint(total: int, count: int) -> float:
    """AvgFunc::ComputeAvg<hugeint_t>, agg_funcs.h:524-540."""
    q, rem = _trunc_divmod(total, count)
    return float(np.float64(hugeint_to_double(q)) + np.float64(hugeint_to_double(rem)) / np.float64(count))


class _GroupState:
    __slots__ = ("key_vals", "count_star", "per_func")

    def __init__(self, key_vals, nfuncs):
        self.key_vals = key_vals
        self.count_star = 0
        self.per_func = [None] * nfuncs  # function-specific running state, None == "no valid value yet"


def hash_aggregate_rowwise(batches: Sequence[pa.RecordBatch], groupby_cols: Sequence[str], agg_cols: Sequence[str],
                           funcs: Sequence[Tuple[str, str, str]]) -> pa.RecordBatch:
    """Literal transcription of BaseAggregate::Next / Result
    (vinum_cpp/src/operators/aggregate/base_aggregate.cpp:23-68) with
    SingleNumerical/MultiNumerical key handling (single_numerical_hash_aggregate.cpp:15-46,
    multi_numerical_hash_aggregate.cpp:17-45) and the per-row Init/Update of
    agg_funcs.h.  `funcs` = [(type, column_name, out_col_name)].  Group order: first
    appearance, NULL single-key group last (single_...cpp:54-60)."""
    groups: Dict[tuple, _GroupState] = {}
    order: List[tuple] = []
    schema = None
    single = len(groupby_cols) == 1
    for batch in batches:
        schema = schema or batch.schema
        for name in list(groupby_cols) + list(agg_cols):
            if batch.schema.get_field_index(name) == -1:
                raise RuntimeError("Column not found: " + name)  # base_aggregate.cpp:121-131
        kcols = [_physical(batch.column(batch.schema.get_field_index(n))) for n in groupby_cols]
        kraw = [key_as_uint64(v) for v, _ in kcols]
        fcols = [(_physical(batch.column(batch.schema.get_field_index(c))) if c else None) for _, c, _ in funcs]
        if not groupby_cols and () not in groups:
            # OneGroupAggregate::Next creates its single group on the first batch, even an
            # empty one (InitBatch, one_group_aggregate.cpp:13-19)
            groups[()] = _GroupState((), len(funcs))
            order.append(())
        for r in range(batch.num_rows):
            key = tuple((None if not kcols[k][1][r] else int(kraw[k][r])) for k in range(len(groupby_cols)))
            st = groups.get(key)
            if st is None:
                key_vals = tuple((None if not kcols[k][1][r] else kcols[k][0][r]) for k in range(len(groupby_cols)))
                st = groups[key] = _GroupState(key_vals, len(funcs))
                order.append(key)
            st.count_star += 1
            for f, (ftype, col, _out) in enumerate(funcs):
                if ftype == COUNT_STAR:
                    continue
                vals, valid = fcols[f]
                if ftype == COUNT:  # CountFunc, agg_funcs.h:129-161
                    st.per_func[f] = (st.per_func[f] or 0) + (1 if valid[r] else 0)
                    continue
                if not valid[r]:  # NextIfNull(): NULL inputs are skipped
                    continue
                v = vals[r]
                cur = st.per_func[f]
                if ftype in (MIN, MAX):  # MinMaxFunc::Update, agg_funcs.h:187-200
                    if cur is None or bool(v < cur) ^ (ftype == MAX):
                        st.per_func[f] = v
                elif ftype == SUM:  # SumFunc / SumOverflowFunc, agg_funcs.h:280-356
                    if vals.dtype.kind == "f":
                        st.per_func[f] = np.float64(v) if cur is None else np.float64(cur + np.float64(v))
                    else:
                        st.per_func[f] = int(v) if cur is None else cur + int(v)
                else:  # AvgFunc, agg_funcs.h:439-542: (sum, count)
                    if vals.dtype.kind == "f":
                        s, c = (np.float64(0.0), 0) if cur is None else cur
                        st.per_func[f] = (np.float64(v) if cur is None else np.float64(s + np.float64(v)), c + 1)
                    else:
                        s, c = (0, 0) if cur is None else cur
                        st.per_func[f] = (s + int(v), c + 1)
    if schema is None:
        return pa.RecordBatch.from_arrays([], names=[])
    if single and (None,) in groups:  # NULL group is summarised last
        order = [k for k in order if k != (None,)] + [(None,)]
    return _build_result(schema, groupby_cols, agg_cols, funcs,
                         [groups[k].key_vals for k in order], [groups[k].count_star for k in order],
                         [[groups[k].per_func[f] for k in order] for f in range(len(funcs))])


def _build_result(schema, groupby_cols, agg_cols, funcs, key_vals, count_star, states) -> pa.RecordBatch:
    """BaseAggregate::Result + the Summarize() of every function, with the output type
    table of agg_func_factory.cpp."""
    arrays, names = [], []
    for name in agg_cols:  # GroupBuilder columns, base_aggregate.cpp:100-108
        k = list(groupby_cols).index(name)
        t = schema.field(name).type
        col = [kv[k] for kv in key_vals]
        arrays.append(_typed_array(col, t))
        names.append(name)
    for f, (ftype, col, out) in enumerate(funcs):
        t = schema.field(col).type if col else None
        st = states[f]
        if ftype == COUNT_STAR:
            arr = pa.array([int(c) for c in count_star], type=pa.uint64())
        elif ftype == COUNT:
            arr = pa.array([int(s or 0) for s in st], type=pa.uint64())
        elif ftype in (MIN, MAX):
            arr = _typed_array(st, t)
        elif ftype == SUM:
            out_t = _sum_output_type(t)
            if pa.types.is_floating(t):
                arr = pa.array([None if s is None else float(s) for s in st], type=pa.float64())
            elif t in (pa.int64(), pa.uint64()):
                cast = hugeint_try_cast_int64 if t == pa.int64() else hugeint_try_cast_uint64
                narrowed = [None if s is None else cast(s) for s in st]
                if any(s is not None and nv is None for s, nv in zip(st, narrowed)):
                    # any group overflows -> the whole column is decimal128(38, 0), agg_funcs.h:358-397
                    import decimal
                    arr = pa.array([None if s is None else decimal.Decimal(s) for s in st], type=pa.decimal128(38, 0))
                else:
                    arr = pa.array(narrowed, type=out_t)
            else:
                width = out_t.bit_width if hasattr(out_t, "bit_width") else 64
                signed = not pa.types.is_unsigned_integer(out_t)
                arr = _typed_array([None if s is None else _wrap(s, width, signed) for s in st], out_t)
        else:  # AVG
            out_t = _avg_output_type(t)
            vals = []
            for s in st:
                if s is None:
                    vals.append(None)
                    continue
                total, cnt = s
                if pa.types.is_floating(t):
                    v = np.float64(total) / np.float64(cnt)          # agg_funcs.h:519-522
                elif t in (pa.int64(), pa.uint64()):
                    v = avg_hugeint(int(total), int(cnt))             # agg_funcs.h:524-540
                else:
                    acc = _wrap(int(total), 64, not pa.types.is_unsigned_integer(t))
                    v = np.float64(acc) / np.float64(cnt)
                vals.append(float(np.float32(v)) if out_t == pa.float32() else float(v))
            arr = pa.array(vals, type=out_t)
        arrays.append(arr)
        names.append(out)
    return pa.RecordBatch.from_arrays(arrays, names=names)


def _wrap(v: int, bits: int, signed: bool) -> int:
    v &= (1 << bits) - 1
    if signed and v >> (bits - 1):
        v -= 1 << bits
    return v


def _typed_array(values: Sequence, t: pa.DataType) -> pa.Array:
    """Physical values (NumPy scalars or None) -> Arrow array of logical type `t`."""
    if pa.types.is_boolean(t):
        return pa.array([None if v is None else bool(v) for v in values], type=t)
    if pa.types.is_floating(t):
        return pa.array([None if v is None else float(v) for v in values], type=t)
    if pa.types.is_integer(t):
        return pa.array([None if v is None else int(v) for v in values], type=t)
    phys = pa.int32() if (pa.types.is_date32(t) or pa.types.is_time32(t)) else pa.int64()
    return pa.array([None if v is None else int(v) for v in values], type=phys).view(t)


def hash_aggregate(batches: Sequence[pa.RecordBatch], groupby_cols: Sequence[str], agg_cols: Sequence[str],
                   funcs: Sequence[Tuple[str, str, str]]) -> pa.RecordBatch:
    """Vectorised equivalent of `hash_aggregate_rowwise` for large inputs.  Float sums
    use np.bincount, which accumulates sequentially in row order exactly like
    `*last += row_val` (agg_funcs.h:298-305).  MIN/MAX over floats containing NaN are
    order dependent in the reference; this fast path requires NaN-free float inputs for
    MIN/MAX and falls back to the row-wise transcription otherwise."""
    batches = [b for b in batches]
    if not batches:
        return pa.RecordBatch.from_arrays([], names=[])
    schema = batches[0].schema
    table = pa.Table.from_batches(batches).combine_chunks()
    n = table.num_rows
    for _ftype, col, _ in funcs:
        if _ftype in (MIN, MAX) and col and pa.types.is_floating(schema.field(col).type):
            v, valid = _physical(table.column(col))
            if np.isnan(v[valid]).any():
                return hash_aggregate_rowwise(batches, groupby_cols, agg_cols, funcs)
    nk = len(groupby_cols)
    kphys = [_physical(table.column(c)) for c in groupby_cols]
    if nk:
        rec = np.zeros(n, dtype=[(f"k{i}", np.uint64) for i in range(nk)] + [(f"n{i}", np.uint8) for i in range(nk)])
        for i, (v, valid) in enumerate(kphys):
            rec[f"k{i}"] = np.where(valid, key_as_uint64(v), np.uint64(0))
            rec[f"n{i}"] = ~valid
        _uniq, first, inv = np.unique(rec, return_index=True, return_inverse=True)
        inv = inv.reshape(-1)
        # groups in order of first appearance (the reference's order is unspecified)
        order = np.argsort(first, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        gid = rank[inv]
        first = first[order]
        ng = len(first)
        if nk == 1 and (~kphys[0][1]).any():  # NULL single-key group last
            null_g = gid[np.flatnonzero(~kphys[0][1])[0]]
            perm = np.concatenate([np.delete(np.arange(ng), null_g), [null_g]])
            inv_perm = np.empty(ng, dtype=np.int64)
            inv_perm[perm] = np.arange(ng)
            gid = inv_perm[gid]
            first = first[perm]
    else:
        ng = 1
        gid = np.zeros(n, dtype=np.int64)
        first = np.zeros(1, dtype=np.int64)
    key_vals = [tuple((None if not kphys[k][1][r] else kphys[k][0][r]) for k in range(nk)) for r in first] if n or nk == 0 else []
    if nk and n == 0:
        ng = 0
    count_star = np.bincount(gid, minlength=ng) if n else np.zeros(ng, dtype=np.int64)
    states = []
    for ftype, col, _out in funcs:
        if ftype == COUNT_STAR:
            states.append([None] * ng)
            continue
        v, valid = _physical(table.column(col))
        g = gid[valid]
        vv = v[valid]
        nvalid = np.bincount(g, minlength=ng) if len(g) else np.zeros(ng, dtype=np.int64)
        if ftype == COUNT:
            states.append([int(c) for c in nvalid])
        elif ftype in (MIN, MAX):
            if vv.dtype == bool:
                vv = vv.astype(np.uint8)
            red = np.minimum if ftype == MIN else np.maximum
            if vv.dtype.kind == "f":
                init = np.inf if ftype == MIN else -np.inf
            else:
                init = np.iinfo(vv.dtype).max if ftype == MIN else np.iinfo(vv.dtype).min
            acc = np.full(ng, init, dtype=vv.dtype)
            red.at(acc, g, vv)
            if vv.dtype.kind == "f":
                # ties between -0.0 and +0.0 are order dependent (first for MIN, last for MAX)
                zero_groups = np.flatnonzero((acc == 0) & (nvalid > 0))
                for zg in zero_groups:
                    zs = vv[(g == zg) & (vv == 0)]
                    acc[zg] = zs[0] if ftype == MIN else zs[-1]
            states.append([acc[i] if nvalid[i] else None for i in range(ng)])
        elif ftype in (SUM, AVG):
            if vv.dtype.kind == "f":
                s = np.bincount(g, weights=vv.astype(np.float64), minlength=ng) if len(g) else np.zeros(ng)
                tot = [np.float64(x) for x in s]
            else:
                # exact integer sums via 32-bit limbs in float64-free arithmetic
                x = vv.astype(np.int64) if vv.dtype.kind == "i" else vv.astype(np.uint64)
                lo = (x & 0xFFFFFFFF).astype(np.int64) if vv.dtype.kind == "i" else (x & np.uint64(0xFFFFFFFF)).astype(np.int64)
                hi = (x >> 32).astype(np.int64) if vv.dtype.kind == "i" else (x >> np.uint64(32)).astype(np.int64)
                tot = [0] * ng
                if len(g):
                    # per-group limb sums stay below 2^63 for fewer than 2^31 rows per group
                    slo = np.zeros(ng, dtype=np.int64)
                    shi = np.zeros(ng, dtype=np.int64)
                    np.add.at(slo, g, lo)
                    np.add.at(shi, g, hi)
                    tot = [int(shi[i]) * (1 << 32) + int(slo[i]) for i in range(ng)]
            if ftype == SUM:
                states.append([tot[i] if nvalid[i] else None for i in range(ng)])
            else:
                states.append([(tot[i], int(nvalid[i])) if nvalid[i] else None for i in range(ng)])
        else:
            raise ValueError(ftype)
    return _build_result(schema, groupby_cols, agg_cols, funcs, key_vals, [int(c) for c in count_star], states)


def one_group_aggregate(batches: Sequence[pa.RecordBatch], funcs: Sequence[Tuple[str, str, str]]) -> pa.RecordBatch:
    """OneGroupAggregate::Next / SummarizeGroups (one_group_aggregate.cpp:9-37): the
    un-grouped reduction; a single output row even for empty input."""
    return hash_aggregate(batches, [], [], funcs)


# ================================================================ a15-a16: sort
def sort_indices(table: pa.Table, sort_cols: Sequence[str], sort_order: Sequence[str]) -> np.ndarray:
    """arrow::compute::SortIndices as called from Sort::Sorted
    (vinum_cpp/src/operators/sort/sort.cpp:22-38; Arrow is a third-party dependency,
    pinned ==3.0.0 in setup.py:33, absent from the reference tree).  Published
    semantics restated: stable; per-key direction; within a key, NaN sorts after every
    number and NULL after NaN in BOTH directions; -0.0 == +0.0."""
    n = table.num_rows
    keys = []
    for name, order in zip(sort_cols, sort_order):
        v, valid = _physical(table.column(name))
        if v.dtype == bool:
            raise RuntimeError("Failed to sort table.")
        desc = order == DESC
        if v.dtype.kind == "f":
            v = v.astype(np.float64)
            isnan = np.isnan(v) & valid
            cls = np.where(~valid, 2, np.where(isnan, 1, 0)).astype(np.int8)
            val = np.where(cls == 0, v, 0.0) + 0.0  # -0.0 -> +0.0
            val = -val if desc else val
        else:
            cls = np.where(valid, 0, 2).astype(np.int8)
            # exact for the full int64/uint64 range: compare through Python-int free ranks
            x = np.where(valid, v, v.dtype.type(0))
            _u, rank = np.unique(x, return_inverse=True)
            val = rank.reshape(-1).astype(np.int64)
            val = -val if desc else val
        keys.append((cls, val))
    # np.lexsort: last key is the primary one; stable
    cols = []
    for cls, val in reversed(keys):
        cols.append(val)
        cols.append(cls)
    return np.lexsort(tuple(cols)) if n else np.empty(0, dtype=np.int64)


def sort_table(table: pa.Table, sort_cols: Sequence[str], sort_order: Sequence[str]) -> pa.RecordBatch:
    """Sort::Sorted, sort.cpp:15-63: SortIndices + Take + CombineChunks -> one batch."""
    idx = sort_indices(table, sort_cols, sort_order)
    out = table.combine_chunks().take(pa.array(idx, type=pa.int64())).combine_chunks()
    batches = out.to_batches()
    if not batches:
        return pa.RecordBatch.from_arrays([pa.array([], type=f.type) for f in table.schema], schema=table.schema)
    return batches[0]


# =========================================== reference operator chain, restated
def filter_hash_aggregate(table: pa.Table, pred_col: str, op: str, scalar, groupby_cols: Sequence[str],
                          funcs: Sequence[Tuple[str, str, str]], batch_size: int = 10000) -> pa.RecordBatch:
    """The north-star plan as the reference executes it:
    TableReaderOperator (batch_size rows, vinum/core/algebra.py:250-265) ->
    FilterOperator (:108-123) -> AggregateOperator (vinum/core/aggregate.py:114-124)."""
    out_batches = []
    for batch in table.to_batches(max_chunksize=batch_size):
        mask = compare(batch.column(batch.schema.get_field_index(pred_col)), op, scalar)
        out_batches.append(filter_batch(batch, mask))
    return hash_aggregate(out_batches, groupby_cols, groupby_cols, funcs)
