"""Streaming hash aggregate on the device + Arrow result assembly.

`Aggregator` owns one VkAgg (include/vinum_b200.h) and mirrors the life cycle of the
reference's C++ aggregate objects: created lazily on the first batch, fed every batch
(`BaseAggregate::Next`, vinum_cpp/src/operators/aggregate/base_aggregate.cpp:23-45),
asked once for the result (`BaseAggregate::Result`, :47-68).  The output type table is
the reference's (`agg_func_factory.cpp:13-329`).
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa

from . import _lib as L
from ._lib import lib
from .device import (_shutting_down, DeviceBuffer, DeviceColumn, Stream, VK_SIZE, VK_TO_NUMPY, default_stream, arrow_from_numpy,
                     vk_dtype_of)
from .ops import Predicate

FUNC_NAMES = {L.AGG_COUNT_STAR: "count_star", L.AGG_COUNT: "count", L.AGG_MIN: "min", L.AGG_MAX: "max",
              L.AGG_SUM: "sum", L.AGG_AVG: "avg"}

_INT64_MAX = 0x7FFFFFFFFFFFFFFF

_PINNED_SCRATCH = {}


def _pinned_scratch(nbytes: int):
    """One page-locked read-back block per thread, grown on demand (cudaHostAlloc costs far more than a query)."""
    import threading
    from .device import PinnedBuffer
    key = threading.get_ident()
    buf = _PINNED_SCRATCH.get(key)
    if buf is None or buf.nbytes < nbytes:
        buf = PinnedBuffer(max(nbytes, 1 << 16))
        _PINNED_SCRATCH[key] = buf
    return buf


# Large results (1e6 groups = 48 MB) are handed out as VIEWS of the page-locked block they were copied into: one more
# host copy of the block costs more than the PCIe transfer.  The block goes back to this free list when the last view
# of it dies (results of a later query then overwrite it), so a loop of queries pins memory once.
_PINNED_POOL: list = []
_PINNED_POOL_MAX = 8
_PINNED_POOL_BYTES = 2 << 30
_LARGE_RESULT = 4 << 20


def _pinned_take(nbytes: int):
    from .device import PinnedBuffer
    best = None
    for i, b in enumerate(_PINNED_POOL):
        if b.nbytes >= nbytes and (best is None or b.nbytes < _PINNED_POOL[best].nbytes):
            best = i
    if best is not None:
        try:
            return _PINNED_POOL.pop(best)
        except IndexError:      # another thread took it
            pass
    return PinnedBuffer(nbytes)


def _pinned_give_back(buf) -> None:
    # keep the large blocks (pinning 400 MB costs ~0.1 s, as does unpinning it): when the list is over its budget the
    # SMALLEST block goes, never the one a loop of large queries is about to ask for again
    _PINNED_POOL.append(buf)
    while len(_PINNED_POOL) > _PINNED_POOL_MAX or (len(_PINNED_POOL) > 1 and sum(b.nbytes for b in _PINNED_POOL) > _PINNED_POOL_BYTES):
        _PINNED_POOL.remove(min(_PINNED_POOL, key=lambda b: b.nbytes))


def check_supported(func: int, t: Optional[pa.DataType]) -> None:
    """Type errors of agg_func_factory.cpp (same messages)."""
    if func in (L.AGG_COUNT_STAR, L.AGG_COUNT):
        return
    bad_sum_avg = (pa.types.is_boolean(t) or pa.types.is_date(t) or pa.types.is_timestamp(t))
    if func == L.AGG_SUM and bad_sum_avg:
        raise RuntimeError("Column data type is not supported by sum().")  # agg_func_factory.cpp:150-175
    if func == L.AGG_AVG and bad_sum_avg:
        raise RuntimeError("Column data type is not supported by avg().")  # agg_func_factory.cpp:221-246


def output_type(func: int, t: Optional[pa.DataType]) -> pa.DataType:
    """Result Arrow type per aggregate function and input type."""
    if func in (L.AGG_COUNT_STAR, L.AGG_COUNT):
        return pa.uint64()                                             # agg_funcs.h:97-100,129-133
    if func in (L.AGG_MIN, L.AGG_MAX):
        return t                                                       # agg_func_factory.cpp:35-93
    if func == L.AGG_SUM:
        if pa.types.is_signed_integer(t):
            return pa.int64()                                          # :110-117
        if pa.types.is_unsigned_integer(t):
            return pa.uint64()                                         # :118-125
        if pa.types.is_floating(t):
            return pa.float64()                                        # :126-131
        return t                                                       # time32/time64/duration keep their unit (:132-149)
    if func == L.AGG_AVG:
        if t in (pa.int8(), pa.int16(), pa.uint8(), pa.uint16()):
            return pa.float32()                                        # :179-184,191-196
        return pa.float64()
    raise ValueError(func)


class Aggregator:
    """Group-by (n_keys >= 1) or un-grouped (n_keys == 0) streaming aggregate."""

    def __init__(self, key_types: Sequence[pa.DataType], funcs: Sequence[Tuple[int, Optional[pa.DataType]]],
                 expected_groups: int = 0):
        self.key_types = list(key_types)
        self.funcs = list(funcs)
        self.key_vk = []
        for t in self.key_types:
            dt = vk_dtype_of(t)
            if dt is None or dt == L.BOOL8:
                raise TypeError(f"group-by key type {t} is not supported on the device path")
            self.key_vk.append(dt)
        self.func_vk = []
        for f, t in self.funcs:
            check_supported(f, t)
            if f == L.AGG_COUNT_STAR:
                self.func_vk.append(L.I64)
                continue
            dt = vk_dtype_of(t)
            if dt is None:
                raise TypeError(f"aggregate input type {t} is not supported on the device path")
            if dt == L.BOOL8 and f not in (L.AGG_COUNT, L.AGG_MIN, L.AGG_MAX):
                raise RuntimeError(f"Column data type is not supported by {FUNC_NAMES[f]}().")
            self.func_vk.append(dt)
        nk, nf = len(self.key_vk), len(self.funcs)
        kd = (C.c_int32 * max(nk, 1))(*self.key_vk)
        fc = (C.c_int32 * max(nf, 1))(*[f for f, _ in self.funcs])
        fd = (C.c_int32 * max(nf, 1))(*self.func_vk)
        h = C.c_void_p()
        lib.vk_agg_create(C.byref(h), nk, kd, nf, fc, fd, int(expected_groups))
        self._h = h
        self.rows_in = 0

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            L._lib.vk_agg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        if _shutting_down():
            return
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ update
    def update(self, keys: Sequence[DeviceColumn], values: Sequence[Optional[DeviceColumn]],
               pred: Optional[Predicate] = None, stream: Optional[Stream] = None) -> None:
        st = stream or default_stream()
        nk, nf = len(self.key_vk), len(self.funcs)
        if len(keys) != nk or len(values) != nf:
            raise ValueError("key / value column count mismatch")
        n = keys[0].length if nk else (next((v.length for v in values if v is not None), None))
        if n is None:
            n = pred.length if pred is not None else 0
        kcols = (L.VkColumn * max(nk, 1))()
        for i, k in enumerate(keys):
            kcols[i] = k.vk()
        vcols = (L.VkColumn * max(nf, 1))()
        for i, v in enumerate(values):
            if v is not None:
                vcols[i] = v.vk()
        vp = pred.vk() if pred is not None else None
        lib.vk_agg_update(self._h, C.byref(vp) if vp is not None else None, int(n), kcols, vcols, st.ptr)
        self.rows_in += int(n)

    def update_count_rows(self, n_rows: int, pred: Optional[Predicate] = None, stream: Optional[Stream] = None) -> None:
        """Un-grouped COUNT(*) over a batch that has no usable column."""
        st = stream or default_stream()
        vp = pred.vk() if pred is not None else None
        kcols = (L.VkColumn * 1)()
        vcols = (L.VkColumn * max(len(self.funcs), 1))()
        lib.vk_agg_update(self._h, C.byref(vp) if vp is not None else None, int(n_rows), kcols, vcols, st.ptr)

    @property
    def last_path(self) -> int:
        return int(L._lib.vk_agg_last_path(self._h))

    def profile(self, enable: bool = True) -> None:
        """Per-launch CUDA-event timing of the update kernels (no extra syncs)."""
        lib.vk_agg_profile(self._h, int(enable))

    def profile_read(self, path: int = 1):
        """(kernel ms, launches, rows) of kernel path 1 (shared-memory) or 2 (global)."""
        ms, ln, rows = C.c_double(), C.c_int64(), C.c_int64()
        lib.vk_agg_profile_read(self._h, path, C.byref(ms), C.byref(ln), C.byref(rows))
        return ms.value, ln.value, rows.value

    def num_groups(self, stream: Optional[Stream] = None) -> int:
        st = stream or default_stream()
        g = C.c_int64()
        lib.vk_agg_num_groups(self._h, C.byref(g), st.ptr)
        return g.value

    # ------------------------------------------------------------------ result
    def result_raw(self, stream: Optional[Stream] = None, extra_d2h=None):
        """Finalised groups as host NumPy arrays:
        (keys u64[nk][g], key_valid bool[nk][g], count_star u64[g], lo u64[nf][g], hi u64[nf][g], valid bool[nf][g]).

        Group-by aggregates finalise into one packed device block that is copied back with ONE
        memcpy and ONE stream synchronisation (vk_agg_result_packed); the block is sized from the
        last result and doubled on the rare overflow.  `extra_d2h` = (host address, device address,
        bytes) rides on the same synchronisation (the peer exchange's decision word)."""
        st = stream or default_stream()
        nk, nf = len(self.key_vk), len(self.funcs)
        if nk == 0:
            return self._result_raw_unpacked(st)
        cap = getattr(self, "_result_cap", 1024)
        while True:
            nbytes = int(L._lib.vk_agg_result_packed_bytes(nk, nf, cap))
            large = nbytes >= _LARGE_RESULT
            host = _pinned_take(nbytes) if large else _pinned_scratch(nbytes)
            dev = DeviceBuffer(nbytes, st, zero=True)   # the whole block is copied back: no uninitialised padding
            lib.vk_agg_result_packed(self._h, cap, C.c_void_p(dev.ptr), st.ptr)
            lib.vk_memcpy_d2h(C.c_void_p(host.ptr), C.c_void_p(dev.ptr), nbytes, st.ptr)
            if extra_d2h is not None:
                lib.vk_memcpy_d2h(C.c_void_p(extra_d2h[0]), C.c_void_p(extra_d2h[1]), int(extra_d2h[2]), st.ptr)
            st.sync()
            raw = host.as_numpy(np.uint8, nbytes)
            if large:
                weakref.finalize(raw, _pinned_give_back, host)   # every view derived below keeps `raw` alive
            g = int(raw[:8].view(np.uint64)[0])
            if g <= cap:
                break
            cap = 1 << (g - 1).bit_length()
        self._result_cap = max(1024, 1 << max(g - 1, 1).bit_length())
        words = 2 + (nk + 1 + 2 * nf) * cap
        h64 = raw[:words * 8].view(np.uint64)[2:].reshape(nk + 1 + 2 * nf, cap)[:, :g]
        if not large:
            h64 = h64.copy()    # the small scratch block is reused by the next call
        h8 = raw[words * 8:words * 8 + (nk + nf) * cap].reshape(nk + nf, cap)[:, :g].astype(bool)
        return h64[:nk], h8[:nk], h64[nk], h64[nk + 1:nk + 1 + nf], h64[nk + 1 + nf:], h8[nk:]

    def _result_raw_unpacked(self, st: Stream):
        """vk_agg_num_groups + vk_agg_result into separate arrays (the un-grouped aggregate's path)."""
        g = self.num_groups(st)
        nk, nf = len(self.key_vk), len(self.funcs)
        gg = max(g, 1)
        # one allocation, carved: [keys | count | lo | hi] u64 then [key_valid | valid] u8
        words = (nk + 1 + 2 * nf) * gg
        buf64 = DeviceBuffer(words * 8, st)
        buf8 = DeviceBuffer(max((nk + nf) * gg, 1), st)
        p64 = lambda i: buf64.ptr + i * gg * 8
        p8 = lambda i: buf8.ptr + i * gg
        okeys = (C.c_void_p * max(nk, 1))(*[p64(i) for i in range(nk)])
        okv = (C.c_void_p * max(nk, 1))(*[p8(i) for i in range(nk)])
        ocount = C.c_void_p(p64(nk))
        olo = (C.c_void_p * max(nf, 1))(*[p64(nk + 1 + f) for f in range(nf)])
        ohi = (C.c_void_p * max(nf, 1))(*[p64(nk + 1 + nf + f) for f in range(nf)])
        oval = (C.c_void_p * max(nf, 1))(*[p8(nk + f) for f in range(nf)])
        lib.vk_agg_result(self._h, g, okeys, okv, ocount, olo, ohi, oval, st.ptr)
        h64 = buf64.to_numpy(np.uint64, words, st).reshape(nk + 1 + 2 * nf, gg)[:, :g]
        h8 = buf8.to_numpy(np.uint8, max((nk + nf) * gg, 1), st)[:(nk + nf) * gg].reshape(nk + nf, gg)[:, :g]
        keys = h64[:nk]
        count = h64[nk]
        lo = h64[nk + 1:nk + 1 + nf]
        hi = h64[nk + 1 + nf:]
        return keys, h8[:nk].astype(bool), count, lo, hi, h8[nk:].astype(bool)

    def empty_raw(self):
        """`result_raw` of zero groups (what a rank that is not the owner of a sharded query holds)."""
        nk, nf = len(self.key_vk), len(self.funcs)
        z64 = lambda r: np.zeros((r, 0), dtype=np.uint64)
        zb = lambda r: np.zeros((r, 0), dtype=bool)
        return z64(nk), zb(nk), np.zeros(0, dtype=np.uint64), z64(nf), z64(nf), zb(nf)

    def result_arrays(self, stream: Optional[Stream] = None, raw=None) -> Tuple[List[pa.Array], List[pa.Array]]:
        """(key arrays, aggregate arrays) with the reference's output types; `raw` = an already
        finalised `result_raw` tuple (the merged groups of a sharded query) instead of this object's."""
        keys, key_valid, _count, lo, hi, valid = raw if raw is not None else self.result_raw(stream)
        key_arrays = [_key_array(keys[k], key_valid[k], self.key_types[k], self.key_vk[k])
                      for k in range(len(self.key_vk))]
        agg_arrays = []
        for f, (func, t) in enumerate(self.funcs):
            agg_arrays.append(_agg_array(func, t, self.func_vk[f], lo[f], hi[f], valid[f]))
        return key_arrays, agg_arrays


def _narrow(u64: np.ndarray, vk_dtype: int) -> np.ndarray:
    """64-bit lane -> physical values of `vk_dtype` (inverse of NextAsUInt64)."""
    npdt = np.dtype(VK_TO_NUMPY[vk_dtype])
    if vk_dtype == L.F64:
        return u64.view(np.float64)
    if vk_dtype == L.F32:
        return u64.astype(np.uint32).view(np.float32)
    if vk_dtype == L.BOOL8:
        return u64.astype(np.uint8).astype(bool)
    return u64.astype(npdt)  # modular truncation == two's complement narrowing


def _key_array(u64: np.ndarray, valid: np.ndarray, t: pa.DataType, vk_dtype: int) -> pa.Array:
    return arrow_from_numpy(_narrow(np.ascontiguousarray(u64), vk_dtype), valid, t)


def _agg_array(func: int, t: Optional[pa.DataType], vk_dtype: int, lo: np.ndarray, hi: np.ndarray,
               valid: np.ndarray) -> pa.Array:
    lo = np.ascontiguousarray(lo)
    hi = np.ascontiguousarray(hi)
    out_t = output_type(func, t)
    if func in (L.AGG_COUNT_STAR, L.AGG_COUNT):
        return pa.array(lo, type=pa.uint64())
    if func in (L.AGG_MIN, L.AGG_MAX):
        if vk_dtype == L.F32:  # computed in f64 (exact widening), narrowed back
            vals = lo.view(np.float64).astype(np.float32)
        else:
            vals = _narrow(lo, vk_dtype)
        return arrow_from_numpy(vals, valid, out_t)
    if func == L.AGG_AVG:
        vals = lo.view(np.float64)
        if out_t == pa.float32():
            vals = vals.astype(np.float32)
        return arrow_from_numpy(vals, valid, out_t)
    # ---- SUM ----
    if pa.types.is_floating(t):
        return arrow_from_numpy(lo.view(np.float64), valid, pa.float64())
    if VK_SIZE[vk_dtype] == 8 and (pa.types.is_integer(t)):
        # 128-bit accumulate; int64/uint64 unless ANY group overflows, then the whole
        # column is decimal128(38, 0) (SumOverflowFunc::Summarize, agg_funcs.h:358-397).
        hi_s = hi.view(np.int64)
        if pa.types.is_signed_integer(t):
            # Hugeint::TryCast<int64_t>, huge_int.cpp:341-361: note -2^63 itself fails the cast
            fits = ((hi_s == 0) & (lo <= np.uint64(_INT64_MAX))) | ((hi_s == -1) & (lo > np.uint64(1 << 63)))
            narrow = lo.view(np.int64)
        else:
            fits = (hi_s == 0) | ((hi_s == -1) & (lo > np.uint64(0)))
            narrow = lo
        if bool(np.all(fits | ~valid)):
            return arrow_from_numpy(narrow, valid, out_t)
        pairs = np.empty((len(lo), 2), dtype=np.uint64)
        pairs[:, 0] = lo
        pairs[:, 1] = hi
        return _decimal_array(pairs, valid)
    # small integers and time/duration types: 64-bit wrapping sum narrowed to the output width
    out_vk = vk_dtype_of(out_t)
    return arrow_from_numpy(_narrow(lo, out_vk), valid, out_t)


def _decimal_array(pairs: np.ndarray, valid: np.ndarray) -> pa.Array:
    n = pairs.shape[0]
    null_count = int(n - np.count_nonzero(valid))
    vbuf = pa.py_buffer(np.packbits(valid, bitorder="little").tobytes()) if null_count else None
    return pa.Array.from_buffers(pa.decimal128(38, 0), n, [vbuf, pa.py_buffer(np.ascontiguousarray(pairs))],
                                 null_count=null_count)
