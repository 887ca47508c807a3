"""vinum_b200 -- B200-native (sm_100a) physical operators for Vinum's hot path:
comparison/filter -> projection/arithmetic -> hash group-by aggregate -> sort, over
Arrow column chunks resident in HBM.  See DESIGN.md and include/vinum_b200.h.

Importing the package loads the CUDA library (vinum_b200/_C/libvinum_b200.so); there
is no CPU fallback -- a missing library is an ImportError.
"""
from ._lib import VinumB200Error, lib, LIB_PATH  # noqa: F401  (loads the shared object)
from .device import (DeviceBatch, DeviceBuffer, DeviceColumn, PinnedBuffer, Stream, default_stream,  # noqa: F401
                     device_count, device_info, pinned_array)
from . import ops, datagen, vinum_lib  # noqa: F401
from .aggregate import Aggregator  # noqa: F401
from .ops import Predicate  # noqa: F401
from .table import StreamReader, Table, read_csv, read_json, read_parquet, stream_csv  # noqa: F401
from .sql.functions import register_numpy, register_python  # noqa: F401
from .sharded import ShardedTable  # noqa: F401

__version__ = "0.1.0"


# ------------------------------------------------------------------ configuration ----
import contextlib as _contextlib
import ctypes as _C

_batch_size = 10000


def get_batch_size() -> int:
    """vinum.get_batch_size (vinum/__init__.py:52-57): rows per batch of `vinum_lib.TableBatchReader` when
    the reference's Python layers drive this package.  The engine behind `Table.sql()` ignores it: it cuts
    a table into device-sized chunks (executor.DEFAULT_CHUNK_ROWS) -- 10 000-row batches would be
    launch-latency bound on a GPU."""
    return _batch_size


def set_batch_size(batch_size: int) -> None:
    """vinum.set_batch_size (vinum/__init__.py:60-62)."""
    global _batch_size
    _batch_size = int(batch_size)


def set_device(device: int) -> None:
    """Bind this process to one GPU (one process per GPU; `LOCAL_RANK` under torchrun)."""
    lib.vk_set_device(int(device))


def set_option(name: str, value: int) -> None:
    """Kernel-selection knob; same names as the VINUM_B200_<NAME> environment variables."""
    lib.vk_set_option(name.encode(), int(value))


def get_option(name: str) -> int:
    v = _C.c_int64()
    lib.vk_get_option(name.encode(), _C.byref(v))
    return int(v.value)


@_contextlib.contextmanager
def options(**kwargs):
    """`with vinum_b200.options(AGG_DIRECT=0, SORT_FUSE_LAST=0): ...` -- set knobs, restore them after."""
    old = {k: get_option(k) for k in kwargs}
    try:
        for k, v in kwargs.items():
            set_option(k, v)
        yield
    finally:
        for k, v in old.items():
            set_option(k, v)
