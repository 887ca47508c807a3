"""Built-in scalar functions, applied on the HOST.

In the reference these are NumPy callables or small NumPy/pyarrow classes
(`_default_functions_registry`, vinum/core/functions.py:341-367; `np.*` names are looked up
in the numpy namespace, vinum/core/udf.py:28-64).  Arbitrary Python/NumPy callables cannot
run on the device, so they stay host functions here too: the engine
hands them NumPy arrays (NULL -> NaN, like RecordBatch.get_np_column) or pyarrow string arrays
and uploads the result.  User-defined functions registered with `register_numpy` /
`register_python` (udf.py:67-218) go through the same table.
"""
from __future__ import annotations

from typing import Callable, Dict, List

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc


class FunctionError(Exception):
    """vinum/errors/__init__.py: FunctionError."""


def _cast(dtype: str) -> Callable:
    def f(*args):
        arr = args if len(args) > 1 else args[0]          # functions.py:150-156
        if isinstance(arr, (pa.Array, pa.ChunkedArray)):
            arr = arr.to_numpy(zero_copy_only=False)
        return np.array(arr, dtype=dtype)
    return f


def _to_str_array(a, has_arrays: bool):
    if isinstance(a, np.ndarray):
        return pa.array(np.array(a, dtype="U"))
    if isinstance(a, (pa.Array, pa.ChunkedArray)):
        return a if pa.types.is_string(a.type) else a.cast(pa.string())
    return pa.array((str(a),), type=pa.string()) if has_arrays else str(a)


def _upper(x):
    r = pc.utf8_upper(_to_str_array(x, True))
    return r


def _lower(x):
    return pc.utf8_lower(_to_str_array(x, True))


def _concat(*args):
    """ConcatFunction (functions.py:243-271): np.char.add folded over the arguments as unicode arrays."""
    out = None
    for a in args:
        if isinstance(a, (pa.Array, pa.ChunkedArray)):
            a = a.to_numpy(zero_copy_only=False)
        a = np.array(a, dtype="U")
        out = a if out is None else np.char.add(out, a)
    return pa.array(out) if isinstance(out, np.ndarray) and out.shape != () else str(out)


# ---- datetime constructors (DatetimeFunction and subclasses, functions.py:25-147) ----
_NUMPY_UNITS = ["Y", "M", "W", "D", "h", "m", "s", "ms", "us", "ns"]


def _as_numpy(a):
    if isinstance(a, (pa.Array, pa.ChunkedArray)):
        return a.to_numpy(zero_copy_only=False)
    return a


def _datetime_function(name: str, units, default_unit):
    def f(*args):
        if not args:
            raise FunctionError(f"No arguments provided for {name} operator.")
        arg = _as_numpy(args[0])
        if not isinstance(arg, (np.ndarray, list, tuple)):
            arg = (arg,)                                     # ensure_is_array
        unit = default_unit
        if len(args) > 1:
            unit = args[1]
            if unit not in units:
                raise FunctionError(f"Unsupported {name} unit: '{unit}'. Supported units are: [{', '.join(units)}]")
        arr = np.array(arg, dtype=f"datetime64[{unit}]" if unit else "datetime64")
        got = np.datetime_data(arr.dtype)[0]
        if got not in units:                                 # _ensure_unit_correctness: next finer supported unit
            finer = units[-1]
            if got in _NUMPY_UNITS:
                for u in _NUMPY_UNITS[_NUMPY_UNITS.index(got) + 1:]:
                    if u in units:
                        finer = u
                        break
            arr = arr.astype(f"datetime64[{finer}]")
        return arr
    return f


_datetime = _datetime_function("datetime", ["D", "s", "ms", "us", "ns"], None)
_date = _datetime_function("date", ["D"], "D")
_from_timestamp = _datetime_function("from_timestamp", ["s", "ms", "us", "ns"], "s")


def _now(*_args):
    return _datetime("now")


_REGISTRY: Dict[str, Callable] = {
    "now": _now, "date": _date, "datetime": _datetime, "from_timestamp": _from_timestamp,
    "timedelta": np.timedelta64, "is_busday": np.is_busday,
    "to_bool": _cast("bool"), "to_float": _cast("float"), "to_int": _cast("int"), "to_str": _cast("str"),
    "abs": np.absolute, "sqrt": np.sqrt, "cos": np.cos, "sin": np.sin, "tan": np.tan, "power": np.power,
    "log": np.log, "log2": np.log2, "log10": np.log10, "pi": lambda: np.pi, "e": lambda: np.e,
    "concat": _concat, "upper": _upper, "lower": _lower,
}
_USER: Dict[str, Callable] = {}


def register_numpy(name: str, function: Callable) -> None:
    """vn.register_numpy (vinum/core/udf.py:138-218): `function` receives whole NumPy arrays."""
    _USER[name.lower()] = function


def register_python(name: str, function: Callable) -> None:
    """vn.register_python (udf.py:67-135): `function` receives one row at a time (np.vectorize)."""
    _USER[name.lower()] = np.vectorize(function)


def lookup(name: str) -> Callable:
    key = name.lower()
    if key in _USER:
        return _USER[key]
    if key in _REGISTRY:
        return _REGISTRY[key]
    if key.startswith("np.") or key.startswith("numpy."):
        obj = np
        for part in name.split(".")[1:]:
            obj = getattr(obj, part, None)
            if obj is None:
                break
        if callable(obj):
            return obj
    raise FunctionError(f"Function '{name}' is not found.")


def call_host_function(name: str, args: List):
    fn = lookup(name)
    with np.errstate(all="ignore"):
        return fn(*args)
