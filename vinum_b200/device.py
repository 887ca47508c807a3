"""Arrow column chunks resident in HBM.

`DeviceColumn` is the device-side twin of a `pyarrow.Array` of a fixed-width type: a
values buffer, an optional validity bitmap, an element offset and a length -- the
layout the reference's iterators read on the host
(vinum_cpp/src/common/array_iterators.h:180-187).  `DeviceBatch` is the twin of the
reference's `RecordBatch` wrapper (vinum/arrow/record_batch.py:13-142).

Host -> device copies are `cudaMemcpyAsync` straight from the Arrow buffers (true DMA
when the buffers are pinned, see `pinned_array`); nothing here computes on the CPU.
"""
from __future__ import annotations

import atexit
import sys as _sys

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import pyarrow as pa

from . import _lib as L
from ._lib import lib


# ----------------------------------------------------------------- dtypes ----
_ARROW_TO_VK = {
    pa.int8(): L.I8, pa.int16(): L.I16, pa.int32(): L.I32, pa.int64(): L.I64,
    pa.uint8(): L.U8, pa.uint16(): L.U16, pa.uint32(): L.U32, pa.uint64(): L.U64,
    pa.float32(): L.F32, pa.float64(): L.F64,
}
VK_TO_NUMPY = {
    L.I8: np.int8, L.I16: np.int16, L.I32: np.int32, L.I64: np.int64,
    L.U8: np.uint8, L.U16: np.uint16, L.U32: np.uint32, L.U64: np.uint64,
    L.F32: np.float32, L.F64: np.float64, L.BOOL8: np.bool_,
}
VK_SIZE = {L.I8: 1, L.I16: 2, L.I32: 4, L.I64: 8, L.U8: 1, L.U16: 2, L.U32: 4, L.U64: 8,
           L.F32: 4, L.F64: 8, L.BOOL8: 1}
_NUMPY_TO_VK = {np.dtype(v): k for k, v in VK_TO_NUMPY.items()}


def vk_dtype_of(t: pa.DataType) -> Optional[int]:
    """Physical device dtype of an Arrow type, or None when the device path does not
    handle it (strings, decimals, half floats, nested types ...)."""
    if t in _ARROW_TO_VK:
        return _ARROW_TO_VK[t]
    if pa.types.is_date32(t) or pa.types.is_time32(t):
        return L.I32
    if pa.types.is_date64(t) or pa.types.is_time64(t) or pa.types.is_timestamp(t) or pa.types.is_duration(t):
        return L.I64
    if pa.types.is_boolean(t):
        return L.BOOL8
    return None


def vk_dtype_of_numpy(dt) -> int:
    return _NUMPY_TO_VK[np.dtype(dt)]


def arrow_type_of_vk(dt: int) -> pa.DataType:
    if dt == L.BOOL8:
        return pa.bool_()
    return pa.from_numpy_dtype(VK_TO_NUMPY[dt])


# ---------------------------------------------------------------- streams ----

# Device objects that are still alive when the interpreter exits are NOT released one by
# one: by then the stream they were allocated on (and possibly the CUDA context) is gone,
# and the process is about to give everything back anyway.
_finalizing = False


def _mark_finalizing() -> None:
    global _finalizing
    _finalizing = True


atexit.register(_mark_finalizing)


def _shutting_down() -> bool:
    return _finalizing or _sys.is_finalizing()


class Stream:
    """Thin owner of a cudaStream_t (or a borrowed handle, e.g. torch's current stream)."""

    def __init__(self, handle: Optional[int] = None):
        self._owned = handle is None
        if handle is None:
            h = C.c_void_p()
            lib.vk_stream_create(C.byref(h))
            handle = h.value
        self.handle = handle or 0

    def sync(self) -> None:
        lib.vk_stream_sync(C.c_void_p(self.handle))

    @property
    def ptr(self) -> C.c_void_p:
        return C.c_void_p(self.handle)

    def __del__(self):
        try:
            if _shutting_down():
                return
            if getattr(self, "_owned", False) and getattr(self, "handle", 0):
                L._lib.vk_stream_destroy(C.c_void_p(self.handle))
        except Exception:   # module globals are already gone at interpreter teardown
            pass
            self.handle = 0


_default_stream: Optional[Stream] = None


def default_stream() -> Stream:
    """The library-wide stream operators use when the caller does not pass one."""
    global _default_stream
    if _default_stream is None:
        _default_stream = Stream()
    return _default_stream


def _sp(stream: Optional[Stream]) -> C.c_void_p:
    return (stream or default_stream()).ptr


# ---------------------------------------------------------------- buffers ----
class DeviceBuffer:
    """Owning handle of a device allocation from the stream-ordered pool."""

    __slots__ = ("ptr", "nbytes", "_stream_handle", "_stream", "__weakref__")

    def __init__(self, nbytes: int, stream: Optional[Stream] = None, zero: bool = False):
        st = stream or default_stream()
        p = C.c_void_p()
        lib.vk_malloc(C.byref(p), int(nbytes), st.ptr)
        self.ptr = p.value or 0
        self.nbytes = int(nbytes)
        self._stream_handle = st.handle
        # the allocation is freed in stream order on the stream it came from: keep that stream alive
        # (a Stream collected before its buffers would leave cudaFreeAsync with a destroyed handle)
        self._stream = st
        if zero and nbytes:
            lib.vk_memset(C.c_void_p(self.ptr), 0, int(nbytes), st.ptr)

    def free(self) -> None:
        if self.ptr:
            L._lib.vk_free(C.c_void_p(self.ptr), C.c_void_p(self._stream_handle))
            self.ptr = 0

    def __del__(self):
        if _shutting_down():
            return
        try:
            self.free()
        except Exception:
            pass

    def to_numpy(self, dtype, count: int, stream: Optional[Stream] = None, byte_offset: int = 0) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        if out.nbytes:
            st = stream or default_stream()
            lib.vk_memcpy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr + byte_offset), out.nbytes, st.ptr)
            st.sync()
        return out

    @staticmethod
    def from_host(address: int, nbytes: int, stream: Optional[Stream] = None) -> "DeviceBuffer":
        buf = DeviceBuffer(nbytes, stream)
        if nbytes:
            # pinned / registered source: plain DMA; pageable: the pinned bounce-buffer pool (vk_ingest.cu)
            lib.vk_memcpy_h2d_auto(C.c_void_p(buf.ptr), C.c_void_p(address), int(nbytes), _sp(stream))
        return buf

    @staticmethod
    def from_numpy(arr: np.ndarray, stream: Optional[Stream] = None) -> "DeviceBuffer":
        arr = np.ascontiguousarray(arr)
        buf = DeviceBuffer.from_host(arr.ctypes.data, arr.nbytes, stream)
        # the source must outlive the (possibly still queued) copy
        (stream or default_stream()).sync()
        return buf


class PinnedBuffer:
    """Page-locked host memory (cudaHostAlloc) exposed to NumPy / Arrow without copies."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        lib.vk_host_alloc(C.byref(p), int(max(nbytes, 1)))
        self.ptr = p.value
        self.nbytes = int(nbytes)

    def as_numpy(self, dtype, count: Optional[int] = None) -> np.ndarray:
        dt = np.dtype(dtype)
        n = self.nbytes // dt.itemsize if count is None else count
        arr = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr)).view(dt)[:n]
        # keep the allocation alive as long as any view of it
        return _keepalive_view(arr, self)

    def __del__(self):
        if _shutting_down():
            return
        try:
            if self.ptr:
                L._lib.vk_host_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


class _KeepAliveArray(np.ndarray):
    pass


def _keepalive_view(arr: np.ndarray, owner) -> np.ndarray:
    v = arr.view(_KeepAliveArray)
    v._vk_owner = owner  # type: ignore[attr-defined]
    return v


def pinned_array(values: np.ndarray) -> np.ndarray:
    """Copy `values` into page-locked memory; `pa.array(result)` is zero-copy, so the
    Arrow column built on it is DMA-able by `DeviceColumn.from_arrow`."""
    values = np.ascontiguousarray(values)
    buf = PinnedBuffer(values.nbytes)
    out = buf.as_numpy(values.dtype, values.size)
    out[...] = values.reshape(-1)
    return out


# ---------------------------------------------------------------- columns ----
class DeviceColumn:
    """A fixed-width Arrow array resident in device memory."""

    def __init__(self, data: Optional[DeviceBuffer], validity: Optional[DeviceBuffer], offset: int, length: int,
                 dtype: int, arrow_type: Optional[pa.DataType] = None, null_count: int = 0,
                 data_ptr: Optional[int] = None):
        self.data = data
        self.validity = validity
        self.offset = int(offset)
        self.length = int(length)
        self.dtype = int(dtype)
        self.arrow_type = arrow_type if arrow_type is not None else arrow_type_of_vk(dtype)
        self.null_count = int(null_count) if validity is not None else 0
        self._data_ptr = data_ptr if data_ptr is not None else (data.ptr if data is not None else 0)

    # -- views ---------------------------------------------------------------
    @property
    def data_ptr(self) -> int:
        return self._data_ptr

    @property
    def has_nulls(self) -> bool:
        return self.validity is not None and self.null_count != 0

    def __len__(self) -> int:
        return self.length

    def vk(self, nulls_as_nan: bool = False) -> L.VkColumn:
        c = L.VkColumn()
        c.data = self.data_ptr
        c.validity = self.validity.ptr if self.has_nulls else None
        c.offset = self.offset
        c.length = self.length
        c.dtype = self.dtype
        c.nulls_as_nan = 1 if (nulls_as_nan and self.has_nulls) else 0
        return c

    def slice(self, offset: int, length: int) -> "DeviceColumn":
        c = DeviceColumn(self.data, self.validity, self.offset + offset, length, self.dtype, self.arrow_type,
                         self.null_count, self._data_ptr)
        return c

    # -- construction --------------------------------------------------------
    @staticmethod
    def empty(length: int, dtype: int, arrow_type: Optional[pa.DataType] = None,
              stream: Optional[Stream] = None) -> "DeviceColumn":
        buf = DeviceBuffer(max(length, 1) * VK_SIZE[dtype], stream)
        return DeviceColumn(buf, None, 0, length, dtype, arrow_type)

    @staticmethod
    def from_device_ptr(ptr: int, length: int, dtype: int, owner=None) -> "DeviceColumn":
        """Wrap memory owned by someone else (e.g. a torch tensor's data_ptr())."""
        c = DeviceColumn(None, None, 0, length, dtype, None, 0, data_ptr=ptr)
        c._owner = owner  # keep the owner alive
        return c

    @staticmethod
    def from_numpy(arr: np.ndarray, stream: Optional[Stream] = None) -> "DeviceColumn":
        arr = np.ascontiguousarray(arr)
        dt = vk_dtype_of_numpy(arr.dtype)
        buf = DeviceBuffer.from_numpy(arr, stream)
        return DeviceColumn(buf, None, 0, arr.size, dt)

    @staticmethod
    def from_arrow(arr: pa.Array, stream: Optional[Stream] = None) -> "DeviceColumn":
        """Async H2D of an Arrow array's buffers (the caller keeps `arr` alive until the
        stream is synchronised)."""
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
        dt = vk_dtype_of(arr.type)
        if dt is None:
            raise TypeError(f"column type {arr.type} is not supported on the device path")
        n = len(arr)
        bufs = arr.buffers()
        vbuf, dbuf = bufs[0], bufs[1]
        off = arr.offset
        null_count = arr.null_count
        if dt == L.BOOL8:
            # bit-packed values -> one byte per row on the device
            nbytes = (off % 8 + n + 7) // 8
            staged = DeviceBuffer.from_host(dbuf.address + off // 8, nbytes, stream) if n else DeviceBuffer(1, stream)
            data = DeviceBuffer(max(n, 1), stream)
            if n:
                lib.vk_bits_to_mask(C.c_void_p(staged.ptr), off % 8, n, C.c_void_p(data.ptr), _sp(stream))
            doff = 0
        else:
            es = VK_SIZE[dt]
            base = off - off % 8  # keep data and validity on one shared element offset
            doff = off % 8
            nbytes = (doff + n) * es
            data = DeviceBuffer.from_host(dbuf.address + base * es, nbytes, stream) if n else DeviceBuffer(1, stream)
        validity = None
        if vbuf is not None and null_count != 0 and n:
            vbytes = (off % 8 + n + 7) // 8
            validity = DeviceBuffer.from_host(vbuf.address + off // 8, vbytes, stream)
        col = DeviceColumn(data, validity, off % 8 if dt != L.BOOL8 or validity is not None else 0, n, dt, arr.type,
                           null_count)
        if dt == L.BOOL8:
            # data is unsliced (offset 0) but the validity bitmap keeps its bit offset:
            # re-base the data pointer so one offset serves both
            col._data_ptr = data.ptr - col.offset
        col._host_ref = arr  # the async copy reads these buffers
        return col

    # -- back to the host ----------------------------------------------------
    def to_numpy(self, stream: Optional[Stream] = None) -> np.ndarray:
        """Values as a NumPy array of the physical dtype (NULL slots hold garbage)."""
        npdt = np.dtype(VK_TO_NUMPY[self.dtype])
        out = np.empty(self.length, dtype=npdt)
        if out.nbytes:
            st = stream or default_stream()
            lib.vk_memcpy_d2h(out.ctypes.data_as(C.c_void_p),
                              C.c_void_p(self.data_ptr + self.offset * npdt.itemsize), out.nbytes, st.ptr)
            st.sync()
        return out

    def validity_to_numpy(self, stream: Optional[Stream] = None) -> Optional[np.ndarray]:
        """Boolean validity per row, or None when there are no nulls."""
        if not self.has_nulls:
            return None
        st = stream or default_stream()
        nbytes = (self.offset + self.length + 7) // 8
        raw = self.validity.to_numpy(np.uint8, nbytes, st)
        bits = np.unpackbits(raw, bitorder="little")[self.offset:self.offset + self.length]
        return bits.astype(bool)

    def to_arrow(self, stream: Optional[Stream] = None) -> pa.Array:
        values = self.to_numpy(stream)
        valid = self.validity_to_numpy(stream)
        return arrow_from_numpy(values, valid, self.arrow_type)


def arrow_from_numpy(values: np.ndarray, valid: Optional[np.ndarray], arrow_type: pa.DataType) -> pa.Array:
    """Build an Arrow array of `arrow_type` from physical values + boolean validity
    without reinterpreting (exact for temporal types)."""
    n = len(values)
    if pa.types.is_boolean(arrow_type):
        return pa.array(values.astype(bool), type=pa.bool_(), mask=None if valid is None else ~valid)
    null_count = 0
    vbuf = None
    if valid is not None:
        null_count = int(n - np.count_nonzero(valid))
        if null_count:
            vbuf = pa.py_buffer(np.packbits(valid, bitorder="little").tobytes())
    values = np.ascontiguousarray(values)
    return pa.Array.from_buffers(arrow_type, n, [vbuf, pa.py_buffer(values)], null_count=null_count)


# ----------------------------------------------------------------- batches ----
class DeviceBatch:
    """Named device columns of equal length (twin of vinum.arrow.record_batch.RecordBatch)."""

    def __init__(self, columns: Sequence[DeviceColumn], names: Sequence[str], num_rows: Optional[int] = None):
        self.columns: List[DeviceColumn] = list(columns)
        self.column_names: List[str] = list(names)
        assert len(self.columns) == len(self.column_names)
        self.num_rows = num_rows if num_rows is not None else (self.columns[0].length if self.columns else 0)

    @staticmethod
    def from_arrow(batch, stream: Optional[Stream] = None, columns: Optional[Iterable[str]] = None) -> "DeviceBatch":
        """`batch` is a pyarrow RecordBatch or Table; only `columns` are copied when given."""
        names = list(batch.schema.names)
        want = names if columns is None else [n for n in names if n in set(columns)]
        cols = [DeviceColumn.from_arrow(batch.column(names.index(n)), stream) for n in want]
        return DeviceBatch(cols, want, batch.num_rows)

    def has_column(self, name: str) -> bool:
        return name in self.column_names

    def column(self, name: str) -> DeviceColumn:
        if name not in self.column_names:
            raise ValueError(f'Column "{name}" is not found.')  # record_batch.py:75-76
        return self.columns[self.column_names.index(name)]

    def schema(self) -> pa.Schema:
        return pa.schema([pa.field(n, c.arrow_type) for n, c in zip(self.column_names, self.columns)])

    def to_arrow(self, stream: Optional[Stream] = None) -> pa.RecordBatch:
        arrays = [c.to_arrow(stream) for c in self.columns]
        return pa.RecordBatch.from_arrays(arrays, names=self.column_names)

    def slice(self, offset: int, length: int) -> "DeviceBatch":
        return DeviceBatch([c.slice(offset, length) for c in self.columns], self.column_names, length)


def device_info(device: int = 0) -> Dict[str, int]:
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    tot, free = C.c_uint64(), C.c_uint64()
    lib.vk_device_info(device, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(tot), C.byref(free))
    return {"sm_count": sm.value, "cc_major": maj.value, "cc_minor": mnr.value,
            "total_bytes": tot.value, "free_bytes": free.value}


def device_count() -> int:
    n = C.c_int()
    lib.vk_device_count(C.byref(n))
    return n.value
