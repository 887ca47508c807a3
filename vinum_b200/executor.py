"""Host -> device streaming executor for the hot path.

The reference's executor pulls 10 000-row RecordBatches through the operator chain on
one CPU thread (vinum/executor/executor.py:24-31, vinum/core/base.py:254-260).  Here
the RecordBatch stream is cut into large chunks, each chunk's referenced columns are
copied host -> device on a copy stream (plain DMA when the Arrow buffers are pinned, through
the library's pinned bounce-buffer pool when they are pageable -- vk_memcpy_h2d_auto) into
one of two device staging slots, and the fused
filter -> aggregate kernel consumes slot k on the compute stream while slot k+1 is in
flight -- the WHERE mask and the filtered rows are never materialised
(FilterOperator + AggregateOperator fused; vinum/core/algebra.py:108-123,
vinum/core/aggregate.py:114-124).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import pyarrow as pa

from . import _lib as L
from ._lib import lib
from .aggregate import Aggregator
from .device import (DeviceBuffer, DeviceColumn, Stream, VK_SIZE, default_stream, vk_dtype_of)
from .ops import Predicate

DEFAULT_CHUNK_ROWS = 1 << 24  # the reference's 10 000-row batches would be launch-latency bound

_FUNC_CODES = {"COUNT_STAR": L.AGG_COUNT_STAR, "COUNT": L.AGG_COUNT, "MIN": L.AGG_MIN, "MAX": L.AGG_MAX,
               "SUM": L.AGG_SUM, "AVG": L.AGG_AVG}


class _Event:
    def __init__(self):
        h = C.c_void_p()
        lib.vk_event_create(C.byref(h))
        self.h = h

    def record(self, stream: Stream):
        lib.vk_event_record(self.h, stream.ptr)

    def sync(self):
        lib.vk_event_sync(self.h)

    def __del__(self):
        try:
            L._lib.vk_event_destroy(self.h)
        except Exception:
            pass


class _StagingSlot:
    """Device buffers for one chunk of the referenced columns."""

    def __init__(self, names: Sequence[str], dtypes: Sequence[int], rows: int, stream: Stream):
        self.bufs = {n: DeviceBuffer(max(rows, 1) * VK_SIZE[dt], stream) for n, dt in zip(names, dtypes)}
        self.copied = _Event()
        self.consumed = _Event()
        self.used = False


def filter_aggregate(table: pa.Table, groupby: Sequence[str], funcs: Sequence[Tuple[str, str, str]],
                     where: Optional[Tuple[str, str, object]] = None, chunk_rows: int = DEFAULT_CHUNK_ROWS,
                     expected_groups: int = 0, stats: Optional[dict] = None, exchange=None) -> pa.RecordBatch:
    """`SELECT <groupby>, <funcs> FROM table [WHERE col <op> literal] GROUP BY <groupby>` over a
    HOST pyarrow.Table, end to end on the GPU.

    funcs = [(type, column, out_name)], type in COUNT_STAR/COUNT/MIN/MAX/SUM/AVG.
    Columns must be null-free fixed-width numerics (the fused fast path); use
    vinum_b200.vinum_lib for the general case.  Result column order and types are the
    reference's: group-by columns, then one column per function (base_aggregate.cpp:47-68).

    `exchange(aggregator, stream) -> raw groups` (vinum_b200.sharded): `table` is this rank's row-range
    shard and the partial groups of every rank are merged before they are finalised; the ranks that
    do not own the result return zero groups."""
    schema = table.schema
    used: List[str] = []
    for name in list(groupby) + [c for _, c, _ in funcs if c] + ([where[0]] if where else []):
        if schema.get_field_index(name) == -1:
            raise RuntimeError("Column not found: " + name)
        if name not in used:
            used.append(name)
    dtypes = []
    for name in used:
        dt = vk_dtype_of(schema.field(name).type)
        if dt is None or dt == L.BOOL8:
            raise TypeError(f"column {name}: type {schema.field(name).type} is not supported by the streaming path")
        dtypes.append(dt)
    n = table.num_rows
    compute = default_stream()
    copy = Stream()
    key_types = [schema.field(k).type for k in groupby]
    specs = [(_FUNC_CODES[t], schema.field(c).type if c else None) for t, c, _ in funcs]
    agg = Aggregator(key_types, specs, expected_groups)
    chunk_rows = max(1024, min(chunk_rows, max(n, 1)))
    slots = [_StagingSlot(used, dtypes, chunk_rows, compute) for _ in range(2)]
    compute.sync()

    # contiguous host buffers per column (one Arrow chunk each, no nulls)
    host_ptrs = {}
    keep = []
    for name, dt in zip(used, dtypes):
        col = table.column(name)
        if col.null_count:
            raise TypeError(f"column {name} has NULLs: use vinum_b200.vinum_lib (general path)")
        arr = col.chunk(0) if col.num_chunks == 1 else col.combine_chunks()
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.chunk(0) if arr.num_chunks else pa.array([], type=col.type)
        keep.append(arr)
        host_ptrs[name] = arr.buffers()[1].address + arr.offset * VK_SIZE[dt] if len(arr) else 0

    h2d_bytes = 0
    pos = 0
    k = 0
    while pos < n:
        rows = min(chunk_rows, n - pos)
        slot = slots[k & 1]
        if slot.used:
            # the kernel that last read this slot must be done before it is overwritten
            lib.vk_stream_wait_event(copy.ptr, slot.consumed.h)
        for name, dt in zip(used, dtypes):
            nbytes = rows * VK_SIZE[dt]
            # pinned Arrow buffers are DMA'd in place; pageable ones go through the pinned bounce pool
            lib.vk_memcpy_h2d_auto(C.c_void_p(slot.bufs[name].ptr), C.c_void_p(host_ptrs[name] + pos * VK_SIZE[dt]),
                                   nbytes, copy.ptr)
            h2d_bytes += nbytes
        slot.copied.record(copy)
        # the compute stream waits for this slot's copies; the host runs ahead and queues the
        # next chunk's copies while this chunk's kernel executes
        lib.vk_stream_wait_event(compute.ptr, slot.copied.h)
        cols = {name: DeviceColumn(slot.bufs[name], None, 0, rows, dt, schema.field(name).type)
                for name, dt in zip(used, dtypes)}
        pred = Predicate.compare(cols[where[0]], where[1], where[2]) if where else None
        agg.update([cols[g] for g in groupby], [cols[c] if c else None for _, c, _ in funcs], pred, compute)
        slot.consumed.record(compute)
        slot.used = True
        pos += rows
        k += 1
    if exchange is not None:
        agg, raw = exchange(agg, compute)
        key_arrays, agg_arrays = agg.result_arrays(compute, raw=raw)
    else:
        key_arrays, agg_arrays = agg.result_arrays(compute)
    arrays = list(key_arrays) + list(agg_arrays)
    names = list(groupby) + [o for _, _, o in funcs]
    out = pa.RecordBatch.from_arrays(arrays, names=names)
    if stats is not None:
        stats["h2d_bytes"] = h2d_bytes
        stats["d2h_bytes"] = int(sum(a.nbytes for a in arrays))
        stats["chunks"] = k
        stats["agg_path"] = agg.last_path
    agg.close()
    return out
