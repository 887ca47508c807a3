"""Drop-in replacement of the reference's `vinum_lib` extension module.

Same Python-visible names, constructor signatures, method names and error behaviour
as the pybind11 module of vinum/core/vinum_lib.cpp:20-167, but every operator runs on
the GPU through the C ABI (include/vinum_b200.h):

    AggFuncType, SortOrder, AggFuncDef,
    SingleNumericalHashAggregate, MultiNumericalHashAggregate, OneGroupAggregate,
    GenericHashAggregate (string keys: dictionary codes on the device, strings stay on the host),
    Sort, TableBatchReader, import_pyarrow

`next(batch)` takes a host `pyarrow.RecordBatch` (as the reference does), copies only
the referenced columns to the device asynchronously and launches the update kernel;
`result()` / `sorted()` return a host `pyarrow.RecordBatch`.  A `DeviceBatch`
(vinum_b200.device) is accepted wherever a RecordBatch is, which skips the copy.

To use it in the reference: `sys.modules['vinum_lib'] = vinum_b200.vinum_lib` before
`import vinum`.
"""
from __future__ import annotations

import enum
from typing import List, Optional, Sequence

import pyarrow as pa
import pyarrow.compute as pc

from . import _lib as L
from .aggregate import Aggregator
from .strings import StringMinMax
from .device import DeviceBatch, DeviceColumn, Stream, default_stream, vk_dtype_of
from . import ops


def import_pyarrow() -> int:
    """vinum_lib.cpp:22-23.  The shim has no libarrow_python dependency; 0 == success."""
    return 0


class AggFuncType(enum.IntEnum):
    """vinum_lib.cpp:25-32 (values are the C ABI's VkAggFunc codes)."""
    COUNT_STAR = L.AGG_COUNT_STAR
    COUNT = L.AGG_COUNT
    MIN = L.AGG_MIN
    MAX = L.AGG_MAX
    SUM = L.AGG_SUM
    AVG = L.AGG_AVG


class SortOrder(enum.IntEnum):
    """vinum_lib.cpp:34-37."""
    ASC = L.ASC
    DESC = L.DESC


# pybind11's export_values()
COUNT_STAR, COUNT, MIN, MAX, SUM, AVG = (AggFuncType.COUNT_STAR, AggFuncType.COUNT, AggFuncType.MIN,
                                         AggFuncType.MAX, AggFuncType.SUM, AggFuncType.AVG)
ASC, DESC = SortOrder.ASC, SortOrder.DESC


class AggFuncDef:
    """vinum_lib.cpp:39-51 / agg_funcs.h:20-24.  `column_name == ''` for COUNT(*)."""

    def __init__(self, func: AggFuncType, column_name: str, out_col_name: str):
        self.func = AggFuncType(func)
        self._column_name = str(column_name)
        self._out_col_name = str(out_col_name)

    @property
    def column_name(self) -> str:
        return self._column_name

    @property
    def out_col_name(self) -> str:
        return self._out_col_name

    def __repr__(self) -> str:
        return f"<AggFuncDef col_name: {self._column_name}, out_col_name: {self._out_col_name}>"


def _as_batch_like(batch):
    if isinstance(batch, (pa.RecordBatch, pa.Table, DeviceBatch)):
        return batch
    # the reference aborts the process here (`.ValueOrDie()`, vinum_lib.cpp:62-63)
    raise TypeError(f"expected pyarrow.RecordBatch, got {type(batch).__name__}")


def _rejoin(col: pa.ChunkedArray) -> pa.Array:
    """The chunks of `col` as one array: without a copy when they are consecutive slices of one buffer
    (fixed-width, no NULLs -- what zero-copy row-range slices of a table are), else concatenated."""
    chunks = [c for c in col.chunks if len(c)]
    if not chunks:
        return pa.array([], type=col.type)
    if len(chunks) == 1:
        return chunks[0]
    t = col.type
    try:
        width = t.bit_width // 8 if t.bit_width % 8 == 0 else 0
    except ValueError:
        width = 0
    if width and all(c.null_count == 0 for c in chunks):
        bufs = [c.buffers()[1] for c in chunks]
        start = bufs[0].address + chunks[0].offset * width
        pos, ok = start, True
        for c, b in zip(chunks, bufs):
            if b is None or b.address + c.offset * width != pos:
                ok = False
                break
            pos += len(c) * width
        if ok:
            total = sum(len(c) for c in chunks)
            whole = pa.foreign_buffer(start, total * width, base=(chunks, bufs))
            return pa.Array.from_buffers(t, total, [None, whole])
    return pa.concat_arrays(chunks)


def _schema_of(batch) -> pa.Schema:
    return batch.schema() if isinstance(batch, DeviceBatch) else batch.schema


def _column_to_device(batch, name: str, stream: Stream) -> DeviceColumn:
    if isinstance(batch, DeviceBatch):
        return batch.column(name)
    idx = batch.schema.get_field_index(name)
    return DeviceColumn.from_arrow(batch.column(idx), stream)


class _BaseAggregate:
    """BaseAggregate (base_aggregate.h:15-45): streaming state of one aggregate operator."""

    _min_keys = 1
    _max_keys = L.AGG_MAX_KEYS

    def __init__(self, groupby_cols: Sequence[str], agg_cols: Sequence[str], agg_funcs: Sequence[AggFuncDef]):
        self._groupby_cols = [str(c) for c in groupby_cols]
        self._agg_cols = [str(c) for c in agg_cols]
        self._funcs = list(agg_funcs)
        for f in self._funcs:
            if not isinstance(f, AggFuncDef):
                raise TypeError("agg_funcs must be a list of AggFuncDef")
        if not (self._min_keys <= len(self._groupby_cols) <= self._max_keys):
            raise RuntimeError(f"{type(self).__name__}: unsupported number of group-by columns "
                               f"({len(self._groupby_cols)})")
        self._agg: Optional[Aggregator] = None
        self._key_types: List[pa.DataType] = []
        self._stream = default_stream()
        self._str_funcs: List[int] = []             # indices of MIN / MAX functions over string columns
        self._str_state: dict = {}                  # function index -> StringMinMax

    @staticmethod
    def _is_str(t) -> bool:
        return pa.types.is_string(t) or pa.types.is_large_string(t)

    @staticmethod
    def _validity_column(arr) -> pa.Array:
        """What COUNT(string column) needs on the device: a uint8 column with the strings' validity."""
        import numpy as np
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks()
        mask = arr.is_null().to_numpy(zero_copy_only=False) if arr.null_count else None
        return pa.array(np.zeros(len(arr), dtype=np.uint8), mask=mask)

    def _string_spec(self, i: int, f: "AggFuncDef"):
        """Device spec of function i over a STRING column (AggFuncFactory's string cases,
        agg_func_factory.cpp: COUNT and MIN / MAX only)."""
        if int(f.func) == L.AGG_COUNT:
            return (L.AGG_COUNT, pa.uint8())             # validity only
        if int(f.func) in (L.AGG_MIN, L.AGG_MAX):
            self._str_funcs.append(i)
            return (L.AGG_COUNT, pa.uint8())             # placeholder slot, replaced at result()
        raise RuntimeError("Column data type is not supported by sum()/avg().")

    def _update_strings(self, keys, batch, schema, st) -> None:
        """String MIN / MAX of one host batch: the device groups it by the same keys and reduces rank codes."""
        for i in self._str_funcs:
            f = self._funcs[i]
            if i not in self._str_state:
                self._str_state[i] = StringMinMax(self._key_types, int(f.func) == L.AGG_MIN, schema.field(f.column_name).type)
            self._str_state[i].update(keys, batch.column(schema.get_field_index(f.column_name)), None, st)

    # -- BaseAggregate::EnsureInitAggFuncs, base_aggregate.cpp:91-119
    def _ensure_init(self, schema: pa.Schema) -> None:
        if self._agg is not None:
            return
        for name in self._groupby_cols + self._agg_cols:
            if schema.get_field_index(name) == -1:
                raise RuntimeError("Column not found: " + name)  # base_aggregate.cpp:121-131
        self._key_types = [schema.field(n).type for n in self._groupby_cols]
        specs = []
        for i, f in enumerate(self._funcs):
            if f.column_name:
                if schema.get_field_index(f.column_name) == -1:
                    raise RuntimeError("Column not found: " + f.column_name)
                t = schema.field(f.column_name).type
                specs.append(self._string_spec(i, f) if self._is_str(t) else (int(f.func), t))
            else:
                specs.append((int(f.func), None))
        self._check_key_types(self._key_types)
        self._agg = Aggregator(self._key_types, specs)

    def _check_key_types(self, types) -> None:
        for t in types:
            if not (pa.types.is_integer(t) or pa.types.is_floating(t) or pa.types.is_temporal(t)):
                raise RuntimeError("Unsupported data type for aggregation column.")  # array_iterators.cpp:79

    # The reference feeds 10 000-row batches (vinum/__init__.py:52).  One device launch and one stream
    # synchronisation per such batch is launch-latency bound, so host batches are collected until
    # COALESCE_ROWS rows are pending and go to the device as ONE chunk: consecutive zero-copy slices of
    # the same table (what TableBatchReader yields) are re-joined without touching the data, anything
    # else (FilterOperator's fresh arrays) is concatenated on the host first.
    COALESCE_ROWS = 1 << 22

    def next(self, batch) -> None:
        batch = _as_batch_like(batch)
        self._ensure_init(_schema_of(batch))
        if isinstance(batch, DeviceBatch):
            self._flush()
            self._update(batch)
            return
        if not hasattr(self, "_pending"):
            self._pending, self._pending_rows = [], 0
        if batch.num_rows == 0 and self._pending:
            return
        self._pending.append(batch)
        self._pending_rows += batch.num_rows
        if self._pending_rows >= self.COALESCE_ROWS:
            self._flush()

    def _flush(self) -> None:
        pending = getattr(self, "_pending", None)
        if not pending:
            return
        self._pending, self._pending_rows = [], 0
        if len(pending) == 1:
            chunk = pending[0]
        else:
            names = self._needed_columns(pending[0].schema)
            tables = [b if isinstance(b, pa.Table) else pa.Table.from_batches([b]) for b in pending]
            chunk = pa.table({n: _rejoin(pa.chunked_array([c for t in tables for c in t.column(n).chunks],
                                                          type=tables[0].schema.field(n).type)) for n in names}) \
                if names else pa.table({"__rows": pa.nulls(sum(t.num_rows for t in tables))})
        self._update(chunk)

    def _needed_columns(self, schema: pa.Schema) -> List[str]:
        out = []
        for n in self._groupby_cols + [f.column_name for f in self._funcs if f.column_name]:
            if n not in out:
                out.append(n)
        return out

    def _update(self, batch) -> None:
        st = self._stream
        cache = {}

        def col(name):
            if name not in cache:
                cache[name] = _column_to_device(batch, name, st)
            return cache[name]

        def value(f):
            if not f.column_name:
                return None
            if not isinstance(batch, DeviceBatch) and self._is_str(batch.schema.field(f.column_name).type):
                key = "\0validity:" + f.column_name
                if key not in cache:
                    keep.append(self._validity_column(batch.column(batch.schema.get_field_index(f.column_name))))
                    cache[key] = DeviceColumn.from_arrow(keep[-1], st)
                return cache[key]
            return col(f.column_name)

        keep = []
        keys = [col(n) for n in self._groupby_cols]
        values = [value(f) for f in self._funcs]
        if not keys and all(v is None for v in values):
            self._agg.update_count_rows(batch.num_rows, None, st)
        else:
            self._agg.update(keys, values, None, st)
        if self._str_funcs:
            if isinstance(batch, DeviceBatch):
                raise TypeError("string columns do not live on the device: MIN / MAX over them takes host batches")
            self._update_strings(keys, batch, batch.schema, st)
        if not isinstance(batch, DeviceBatch):
            st.sync()  # the async copies read the caller's Arrow buffers

    def result(self) -> pa.RecordBatch:
        self._flush()
        if self._agg is None:
            # result() before any batch: the reference builds an empty batch with no
            # schema information (base_aggregate.cpp:47-68)
            return pa.RecordBatch.from_arrays([], names=[])
        key_arrays, agg_arrays = self._agg.result_arrays(self._stream)
        agg_arrays = list(agg_arrays)
        for i in self._str_funcs:
            agg_arrays[i] = self._str_state[i].result(key_arrays, len(agg_arrays[i]))
        arrays, names = [], []
        for name in self._agg_cols:  # GROUP_BUILDER columns first, base_aggregate.cpp:100-108
            arrays.append(key_arrays[self._groupby_cols.index(name)])
            names.append(name)
        for f, arr in zip(self._funcs, agg_arrays):
            arrays.append(arr)
            names.append(f.out_col_name)
        return pa.RecordBatch.from_arrays(arrays, names=names)


class SingleNumericalHashAggregate(_BaseAggregate):
    """vinum_lib.cpp:54-71; single_numerical_hash_aggregate.cpp:15-68."""
    _min_keys = 1
    _max_keys = 1


class MultiNumericalHashAggregate(_BaseAggregate):
    """vinum_lib.cpp:73-90; multi_numerical_hash_aggregate.cpp:17-58."""
    _min_keys = 1


class GenericHashAggregate(_BaseAggregate):
    """vinum_lib.cpp:92-109; generic_hash_aggregate.{h:9-43,cpp:6-51}: keys of ANY type.

    Numeric / temporal keys go to the device as they are.  String keys never enter HBM: every
    batch is dictionary-encoded on the host against one dictionary per key column that lives
    as long as the operator, and the device groups by the int32 codes (NULL stays NULL);
    boolean keys are grouped as uint8.  Aggregates over numeric columns run on the device as
    usual.  COUNT over a string column only needs its validity; MIN / MAX over strings
    (StringMinMaxFunc, agg_funcs.h:219-261) reduce order-preserving rank codes on the device,
    batch by batch (vinum_b200.strings.StringMinMax); only groups x batches strings are merged
    on the host at result()."""
    _min_keys = 1

    def __init__(self, groupby_cols, agg_cols, agg_funcs):
        super().__init__(groupby_cols, agg_cols, agg_funcs)
        self._key_kind: List[str] = []
        self._dicts: List[Optional[dict]] = []      # string key -> {value: code}
        self._dict_values: List[Optional[list]] = []
        self._user_key_types: List[pa.DataType] = []

    def _check_key_types(self, types) -> None:
        return

    def _ensure_init(self, schema: pa.Schema) -> None:
        if self._agg is not None:
            return
        for name in self._groupby_cols + self._agg_cols:
            if schema.get_field_index(name) == -1:
                raise RuntimeError("Column not found: " + name)
        self._user_key_types = [schema.field(n).type for n in self._groupby_cols]
        dev_types = []
        for t in self._user_key_types:
            if self._is_str(t):
                self._key_kind.append("str")
                self._dicts.append({})
                self._dict_values.append([])
                dev_types.append(pa.int32())
            elif pa.types.is_boolean(t):
                self._key_kind.append("bool")
                self._dicts.append(None)
                self._dict_values.append(None)
                dev_types.append(pa.uint8())
            elif pa.types.is_integer(t) or pa.types.is_floating(t) or pa.types.is_temporal(t):
                self._key_kind.append("num")
                self._dicts.append(None)
                self._dict_values.append(None)
                dev_types.append(t)
            else:
                raise NotImplementedError(f"GenericHashAggregate over key type {t} is not implemented")
        specs = []
        for i, f in enumerate(self._funcs):
            if not f.column_name:
                specs.append((int(f.func), None))
                continue
            if schema.get_field_index(f.column_name) == -1:
                raise RuntimeError("Column not found: " + f.column_name)
            t = schema.field(f.column_name).type
            specs.append(self._string_spec(i, f) if self._is_str(t) else (int(f.func), t))
        self._key_types = dev_types
        self._agg = Aggregator(dev_types, specs)

    def _encode_key(self, k: int, arr) -> pa.Array:
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks()
        kind = self._key_kind[k]
        if kind == "bool":
            return arr.cast(pa.uint8())
        if kind != "str":
            return arr
        import numpy as np
        import pyarrow.compute as pc
        enc = pc.dictionary_encode(arr)
        local = enc.dictionary.to_pylist()
        table, values = self._dicts[k], self._dict_values[k]
        mapping = np.empty(max(len(local), 1), dtype=np.int32)
        for j, v in enumerate(local):
            code = table.get(v)
            if code is None:
                code = len(values)
                table[v] = code
                values.append(v)
            mapping[j] = code
        idx = enc.indices
        null_mask = idx.is_null().to_numpy(zero_copy_only=False) if idx.null_count else None
        local_codes = idx.fill_null(0).to_numpy(zero_copy_only=False) if idx.null_count else idx.to_numpy()
        return pa.array(mapping[local_codes] if len(local) else np.zeros(len(arr), dtype=np.int32), type=pa.int32(),
                        mask=null_mask)

    def next(self, batch) -> None:
        batch = _as_batch_like(batch)
        if isinstance(batch, DeviceBatch):
            raise TypeError("GenericHashAggregate takes host batches (string keys are encoded on the host)")
        self._ensure_init(batch.schema)
        st = self._stream
        schema = batch.schema
        keep = []
        keys = []
        for k, name in enumerate(self._groupby_cols):
            arr = self._encode_key(k, batch.column(schema.get_field_index(name)))
            keep.append(arr)
            keys.append(DeviceColumn.from_arrow(arr, st))
        values = []
        for f in self._funcs:
            if not f.column_name:
                values.append(None)
                continue
            arr = batch.column(schema.get_field_index(f.column_name))
            if self._is_str(arr.type):
                arr = self._validity_column(arr)
            keep.append(arr)
            values.append(DeviceColumn.from_arrow(arr, st))
        self._agg.update(keys, values, None, st)
        self._update_strings(keys, batch, schema, st)
        st.sync()

    def result(self) -> pa.RecordBatch:
        if self._agg is None:
            return pa.RecordBatch.from_arrays([], names=[])
        key_arrays, agg_arrays = self._agg.result_arrays(self._stream)
        user_keys = []
        for k, arr in enumerate(key_arrays):
            kind = self._key_kind[k]
            if kind == "str":
                values = pa.array(self._dict_values[k], type=self._user_key_types[k])
                arr = values.take(arr) if len(values) else pa.nulls(len(arr), self._user_key_types[k])
            elif kind == "bool":
                arr = arr.cast(pa.bool_())
            user_keys.append(arr)
        agg_arrays = list(agg_arrays)
        for i in self._str_funcs:
            # joined by the DEVICE keys (dictionary codes, uint8 booleans), which mean the same in every batch
            agg_arrays[i] = self._str_state[i].result(key_arrays, len(agg_arrays[i]))
        arrays, names = [], []
        for name in self._agg_cols:
            arrays.append(user_keys[self._groupby_cols.index(name)])
            names.append(name)
        for f, arr in zip(self._funcs, agg_arrays):
            arrays.append(arr)
            names.append(f.out_col_name)
        return pa.RecordBatch.from_arrays(arrays, names=names)


class OneGroupAggregate(_BaseAggregate):
    """vinum_lib.cpp:111-124; one_group_aggregate.cpp:9-37."""
    _min_keys = 0
    _max_keys = 0

    def __init__(self, agg_funcs: Sequence[AggFuncDef]):
        super().__init__([], [], agg_funcs)


class Sort:
    """vinum_lib.cpp:126-142; Sort::Next buffers (sort.cpp:11-13), Sort::Sorted sorts
    everything at once (sort.cpp:15-63)."""

    def __init__(self, sort_cols: Sequence[str], sort_order: Sequence[SortOrder]):
        self._cols = [str(c) for c in sort_cols]
        self._order = [int(SortOrder(o)) for o in sort_order]
        if len(self._cols) != len(self._order):
            raise RuntimeError("Sort: sort_cols and sort_order differ in length")
        self._batches: List = []
        self._stream = default_stream()

    def next(self, batch) -> None:
        self._batches.append(_as_batch_like(batch))

    def sorted(self) -> pa.RecordBatch:
        if not self._batches:
            raise RuntimeError("Failed to create table from record batches.")  # sort.cpp:17-19
        st = self._stream
        if all(isinstance(b, DeviceBatch) for b in self._batches) and len(self._batches) == 1:
            dev = self._batches[0]
            for name in self._cols:
                if name not in dev.column_names:
                    raise RuntimeError("Failed to sort table.")
                if pa.types.is_boolean(dev.column(name).arrow_type):
                    raise RuntimeError("Failed to sort table.")  # Arrow 3.0 could not sort booleans (algebra.py:191-201)
            return ops.sort_batch(dev, self._cols, self._order, st).to_arrow(st)
        host = [b.to_arrow(st) if isinstance(b, DeviceBatch) else b for b in self._batches]
        table = pa.Table.from_batches([b for b in host]) if not isinstance(host[0], pa.Table) else pa.concat_tables(host)
        table = table.combine_chunks()
        schema = table.schema
        for name in self._cols:
            if schema.get_field_index(name) == -1:
                raise RuntimeError("Failed to sort table.")  # sort.cpp:34-36
            if pa.types.is_boolean(schema.field(name).type):
                raise RuntimeError("Failed to sort table.")  # Arrow 3.0 could not sort booleans (algebra.py:191-201)
        # The reference hands the WHOLE batch to Sort (algebra.py:160-175): string payload columns and
        # string sort keys included.  Only fixed-width columns live on the device: string keys sort by
        # order-preserving rank codes (NULL stays NULL -> last), string payloads are taken on the host
        # with the device permutation.
        dev_names = [f.name for f in schema if vk_dtype_of(f.type) is not None]
        host_names = [f.name for f in schema if vk_dtype_of(f.type) is None]
        keys = []
        for name in self._cols:
            col = table.column(name)
            if name in host_names:
                arr = col.chunk(0) if col.num_chunks == 1 else col.combine_chunks()
                uniq = pc.unique(arr).drop_null()
                ranked = uniq.take(pc.sort_indices(uniq))
                keys.append(DeviceColumn.from_arrow(pc.index_in(arr, value_set=ranked), st))
            else:
                keys.append(None)
        dev = DeviceBatch.from_arrow(table.select(dev_names), st) if dev_names else None
        keys = [k if k is not None else dev.column(name) for k, name in zip(keys, self._cols)]
        idx, sorted0 = ops.sort_indices_keys(keys, self._order, st)
        arrays = {}
        for name in dev_names:
            c = dev.column(name)
            arrays[name] = (sorted0 if (sorted0 is not None and c is keys[0]) else ops.take(c, idx, st)).to_arrow(st)
        if host_names:
            hidx = pa.array(idx.to_numpy(st))
            for name in host_names:
                arrays[name] = table.column(name).combine_chunks().take(hidx)
        return pa.RecordBatch.from_arrays([_as_array(arrays[f.name]) for f in schema], schema=schema)


def _as_array(a):
    if isinstance(a, pa.ChunkedArray):
        return a.chunk(0) if a.num_chunks == 1 else pa.concat_arrays(a.chunks) if a.num_chunks else pa.array([], type=a.type)
    return a


class TableBatchReader:
    """vinum_lib.cpp:144-165; table_batch_reader.cpp:5-16 -- zero-copy row-range slices
    of the table, one RecordBatch per call, None at the end."""

    def __init__(self, table: pa.Table):
        if not isinstance(table, pa.Table):
            raise TypeError(f"expected pyarrow.Table, got {type(table).__name__}")
        self._table = table
        self._batch_size: Optional[int] = None
        self._iter = None

    def set_batch_size(self, batch_size: int) -> None:
        self._batch_size = int(batch_size)

    def next(self) -> Optional[pa.RecordBatch]:
        if self._iter is None:
            self._iter = iter(self._table.to_batches(max_chunksize=self._batch_size))
        return next(self._iter, None)
