"""MIN / MAX over string columns with the per-row work on the device.

Replaces StringMinMaxFunc (vinum_cpp/src/operators/aggregate/agg_funcs.h:219-261: one
`std::string` compare per row inside the group's hash-map entry).  Strings never enter HBM.
For every batch the host touches each string ONCE, to dictionary-encode it; the dictionary is
sorted, so that a row's code is the rank of its string among the batch's distinct values
(NULL stays NULL) and integer order == string order (binary, like `std::string::operator<`
on UTF-8 bytes).  The device then groups the batch by the operator's keys and takes
MIN / MAX of the int32 codes -- the same kernels as any other MIN / MAX -- and only one code
per group comes back.  Codes of different batches are not comparable, so the per-batch
winners (groups x batches strings, not rows) are merged on the host at result time.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

from . import _lib as L
from .aggregate import Aggregator
from .device import DeviceColumn, Stream


def rank_codes(arr) -> Tuple[pa.Array, pa.Array]:
    """(int32 codes, sorted distinct values): code i <=> i-th smallest distinct string; NULL stays NULL."""
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks()
    enc = pc.dictionary_encode(arr)
    dictionary = enc.dictionary
    order = pc.sort_indices(dictionary)                  # positions of the dictionary in sorted order
    rank = np.empty(len(dictionary), dtype=np.int32)
    rank[order.to_numpy()] = np.arange(len(dictionary), dtype=np.int32)
    idx = enc.indices
    if idx.null_count:
        mask = idx.is_null().to_numpy(zero_copy_only=False)
        codes = pa.array(rank[idx.fill_null(0).to_numpy(zero_copy_only=False)] if len(dictionary) else
                         np.zeros(len(arr), dtype=np.int32), type=pa.int32(), mask=mask)
    else:
        codes = pa.array(rank[idx.to_numpy()] if len(dictionary) else np.zeros(len(arr), dtype=np.int32), type=pa.int32())
    return codes, dictionary.take(order)


def key_matrix(key_arrays: Sequence[pa.Array], n_rows: int) -> np.ndarray:
    """Group keys as rows of int64 (value bits, validity) pairs: equal rows <=> the same group, with
    NULL == NULL and NaN payloads compared by their bits, as the device table does."""
    cols: List[np.ndarray] = []
    for a in key_arrays:
        if isinstance(a, pa.ChunkedArray):
            a = a.combine_chunks()
        valid = np.ones(len(a), dtype=bool) if not a.null_count else ~a.is_null().to_numpy(zero_copy_only=False)
        t = a.type
        if pa.types.is_floating(t):
            v = a.fill_null(0).to_numpy(zero_copy_only=False).astype(np.float64).view(np.int64)
        elif pa.types.is_boolean(t):
            v = a.fill_null(False).to_numpy(zero_copy_only=False).astype(np.int64)
        elif pa.types.is_temporal(t):
            v = a.cast(pa.int64() if t.bit_width == 64 else pa.int32()).fill_null(0).to_numpy(zero_copy_only=False).astype(np.int64)
        elif t == pa.uint64():
            v = a.fill_null(0).to_numpy(zero_copy_only=False).view(np.int64)
        else:
            v = a.fill_null(0).to_numpy(zero_copy_only=False).astype(np.int64)
        cols.append(np.where(valid, v, 0))
        cols.append(valid.astype(np.int64))
    if not cols:
        return np.zeros((n_rows, 0), dtype=np.int64)
    return np.stack(cols, axis=1)


class StringMinMax:
    """One MIN or MAX over a string column of a (possibly batched) GROUP BY."""

    def __init__(self, key_types: Sequence[pa.DataType], is_min: bool, value_type: pa.DataType):
        self._key_types = list(key_types)
        self._is_min = is_min
        self._value_type = value_type
        self._partials: List[Tuple[np.ndarray, pa.Array]] = []   # (key matrix, winning string per group) per batch

    def update(self, key_cols: Sequence[DeviceColumn], values, pred, stream: Stream) -> None:
        """One batch: `key_cols` are the operator's device key columns of this batch, `values` its host string
        column, `pred` the operator's predicate for the batch (or None)."""
        codes, sorted_values = rank_codes(values)
        agg = Aggregator(self._key_types, [(L.AGG_MIN if self._is_min else L.AGG_MAX, pa.int32())])
        try:
            agg.update(list(key_cols), [DeviceColumn.from_arrow(codes, stream)], pred, stream)
            keys, (best,) = agg.result_arrays(stream)
        finally:
            agg.close()
        n_groups = len(best)
        winners = sorted_values.take(best) if len(sorted_values) else pa.nulls(n_groups, self._value_type)
        self._partials.append((key_matrix(keys, n_groups), winners))

    def result(self, final_keys: Sequence[pa.Array], n_groups: int) -> pa.Array:
        """The winning string of every group, in the row order of `final_keys` (the operator's own result)."""
        want = key_matrix(final_keys, n_groups)
        mats = [m for m, _ in self._partials] + [want]
        if want.shape[1] == 0:
            gid = np.zeros(sum(len(m) for m in mats), dtype=np.int64)
        else:
            _, gid = np.unique(np.concatenate(mats, axis=0), axis=0, return_inverse=True)
            gid = gid.reshape(-1).astype(np.int64)
        n_partial = len(gid) - n_groups
        if n_partial == 0:
            return pa.nulls(n_groups, self._value_type)
        strings = pa.concat_arrays([w.cast(self._value_type) for _, w in self._partials])
        merged = pa.table({"g": pa.array(gid[:n_partial]), "s": strings}).group_by("g", use_threads=False).aggregate(
            [("s", "min" if self._is_min else "max")])
        pos = pc.index_in(pa.array(gid[n_partial:]), value_set=merged.column("g").combine_chunks())
        return merged.column(1).combine_chunks().take(pos)
