"""Loader + drivers for the reference's OWN C++ operators compiled into oracle/_ref
(see oracle/build_ref.sh) -- TEST INFRASTRUCTURE ONLY, never on the product path.

`ref_lib()` returns the pybind11 module (or None when it has not been built); the
`ref_*` helpers drive it exactly like the reference's Python operators do
(vinum/core/aggregate.py:114-124, vinum/core/algebra.py:159-177, :108-123).
"""
from __future__ import annotations

import importlib.util
import sysconfig
from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa

_HERE = Path(__file__).resolve().parent
_MOD = None
_TRIED = False


def ref_lib():
    global _MOD, _TRIED
    if _TRIED:
        return _MOD
    _TRIED = True
    so = _HERE / "_ref" / ("ref_vinum_lib" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not so.exists():
        return None
    import pyarrow  # noqa: F401  (libarrow must be loaded before the extension)
    spec = importlib.util.spec_from_file_location("ref_vinum_lib", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.import_pyarrow() != 0:
        raise RuntimeError("reference oracle: import_pyarrow() failed")
    _MOD = mod
    return mod


def _funcdefs(lib, funcs):
    return [lib.AggFuncDef(getattr(lib.AggFuncType, t), c, o) for t, c, o in funcs]


def ref_aggregate(batches: Sequence[pa.RecordBatch], groupby_cols: Sequence[str], agg_cols: Sequence[str],
                  funcs: Sequence[Tuple[str, str, str]]) -> pa.RecordBatch:
    """AggregateOperator.next (vinum/core/aggregate.py:96-124): class chosen from key arity."""
    lib = ref_lib()
    fd = _funcdefs(lib, funcs)
    if len(groupby_cols) == 0:
        agg = lib.OneGroupAggregate(fd)
    elif len(groupby_cols) == 1:
        agg = lib.SingleNumericalHashAggregate(list(groupby_cols), list(agg_cols), fd)
    else:
        agg = lib.MultiNumericalHashAggregate(list(groupby_cols), list(agg_cols), fd)
    for b in batches:
        agg.next(b)
    return agg.result()


def ref_sort(batches: Sequence[pa.RecordBatch], cols: Sequence[str], orders: Sequence[str]) -> pa.RecordBatch:
    """SortOperator.next (vinum/core/algebra.py:159-177) -> Sort::Sorted."""
    lib = ref_lib()
    s = lib.Sort(list(cols), [getattr(lib.SortOrder, o) for o in orders])
    for b in batches:
        s.next(b)
    return s.sorted()


def ref_filter(batch: pa.RecordBatch, mask: np.ndarray) -> pa.RecordBatch:
    """RecordBatch.filter, vinum/arrow/record_batch.py:85-90 (the reference's own call)."""
    return batch.filter(pa.array(np.asarray(mask, dtype=bool)), null_selection_behavior="emit_null")


def ref_filter_hash_aggregate(table: pa.Table, pred_col: str, op: str, scalar, groupby_cols, funcs,
                              batch_size: int = 10000) -> pa.RecordBatch:
    """The reference operator chain for the north-star query: TableBatchReader ->
    NumPy comparison -> RecordBatch.filter -> SingleNumericalHashAggregate."""
    import operator
    ops = {"==": operator.eq, "!=": operator.ne, ">": operator.gt, ">=": operator.ge, "<": operator.lt, "<=": operator.le}
    lib = ref_lib()
    reader = lib.TableBatchReader(table)
    reader.set_batch_size(batch_size)
    fd = _funcdefs(lib, funcs)
    if len(groupby_cols) == 1:
        agg = lib.SingleNumericalHashAggregate(list(groupby_cols), list(groupby_cols), fd)
    else:
        agg = lib.MultiNumericalHashAggregate(list(groupby_cols), list(groupby_cols), fd)
    idx = table.schema.get_field_index(pred_col)
    while True:
        b = reader.next()
        if b is None:
            break
        x = b.column(idx).to_numpy(zero_copy_only=False)
        mask = ops[op](x, scalar)
        agg.next(b.filter(pa.array(mask), null_selection_behavior="emit_null"))
    return agg.result()
