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

__version__ = "0.1.0"
