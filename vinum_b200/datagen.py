"""Deterministic synthetic table: host (NumPy) and device generators that
produce bit-identical columns for any row range.

value(row r, column c) = f_c(splitmix64(r * 16 + c + seed * 0x9E3779B97F4A7C15)).
The device side is `vk_datagen` (vinum_b200/csrc/vk_runtime.cu); the formulas below
are the same single IEEE operations, so results match bit for bit.

(SURVEY 8d sketches `r * 8 + c`; 16 is used so the ninth column, the int32 key `k32`,
gets its own stream instead of aliasing column 0.)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional

import numpy as np
import pyarrow as pa

from . import _lib as L
from ._lib import lib
from .device import DeviceBuffer, DeviceColumn, DeviceBatch, Stream, default_stream

SEED = 42

KINDS = {
    "i0": L.GEN_I0, "i1": L.GEN_I1, "i2": L.GEN_I2, "i3": L.GEN_I3,
    "f0": L.GEN_F0, "f1": L.GEN_F1, "f2": L.GEN_F2, "f3": L.GEN_F3, "k32": L.GEN_K32,
}
DTYPES = {
    "i0": L.I64, "i1": L.I64, "i2": L.I64, "i3": L.I64,
    "f0": L.F64, "f1": L.F64, "f2": L.F64, "f3": L.F64, "k32": L.I32,
}
T8_COLUMNS = ("i0", "i1", "i2", "i3", "f0", "f1", "f2", "f3")

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def host_column(name: str, row0: int, nrows: int, seed: int = SEED) -> np.ndarray:
    """NumPy restatement of the device generator."""
    kind = KINDS[name]
    r = np.arange(row0, row0 + nrows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        u = _splitmix64(r * np.uint64(16) + np.uint64(kind) + np.uint64((seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF))
    if name == "i0":
        return (u % np.uint64(1000)).astype(np.int64)
    if name == "i1":
        return (u >> np.uint64(23)).astype(np.int64) - np.int64(1 << 40)
    if name == "i2":
        return r.astype(np.int64)
    if name == "i3":
        return (u % np.uint64(1000000)).astype(np.int64)
    if name == "k32":
        return (u % np.uint64(1000)).astype(np.int32)
    x = (u >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    if name == "f0":
        return x
    if name == "f1":
        return (x - 0.5) * 2000.0
    if name == "f2":
        m = np.uint64(0xFFFF)
        s = ((u & m).astype(np.float64) + ((u >> np.uint64(16)) & m).astype(np.float64)
             + ((u >> np.uint64(32)) & m).astype(np.float64) + ((u >> np.uint64(48)) & m).astype(np.float64))
        return s * (1.0 / 65536.0) - 2.0
    if name == "f3":
        return x * 1000000.0
    raise KeyError(name)


def host_table(names: Iterable[str], row0: int, nrows: int, seed: int = SEED) -> pa.Table:
    return pa.table({n: host_column(n, row0, nrows, seed) for n in names})


def device_column(name: str, row0: int, nrows: int, seed: int = SEED, stream: Optional[Stream] = None) -> DeviceColumn:
    col = DeviceColumn.empty(nrows, DTYPES[name], stream=stream)
    st = stream or default_stream()
    lib.vk_datagen(KINDS[name], seed, row0, nrows, C.c_void_p(col.data_ptr), st.ptr)
    return col


def device_table(names: Iterable[str], row0: int, nrows: int, seed: int = SEED,
                 stream: Optional[Stream] = None) -> DeviceBatch:
    names = list(names)
    return DeviceBatch([device_column(n, row0, nrows, seed, stream) for n in names], names, nrows)
