"""The reference's OWN Python layers, runnable where /root/reference does not exist (the GPU box).

TEST INFRASTRUCTURE ONLY -- imported by tests/ and by bench.py's reference / cpu_baseline legs, never
by vinum_b200/.

`reference_vinum()` returns the reference's `vinum` package, imported from oracle/_ref/vinum_pyref.zip
(packed from /root/reference/vinum by oracle/build_ref.sh, unmodified), with
  * `vinum_lib` = the reference's own C++ operators compiled into oracle/_ref (oracle/ref.py), or -- for
    the drop-in tests -- `vinum_b200.vinum_lib`;
  * the SQL parser = this repo's stand-in (vinum_b200.compat): the reference's parser needs the pglast
    1.17 C extension, which cannot be installed here.  Everything after the syntax tree -- binder,
    QueryPlanner, RecursiveExecutor, every operator -- is the reference's stock code path
    (vinum/api/table.py:266-274).
"""
from __future__ import annotations

import sys
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_ZIP = _HERE / "_ref" / "vinum_pyref.zip"
_LOADED = None


def available() -> bool:
    from . import ref
    return _ZIP.exists() and ref.ref_lib() is not None


def reference_vinum(gpu_operators: bool = False):
    """`import vinum` on the reference's operators (default) or on vinum_b200's (gpu_operators=True).
    One flavour per process: the reference binds `vinum_lib` at import time."""
    global _LOADED
    if _LOADED is not None:
        if _LOADED[0] != gpu_operators:
            raise RuntimeError("the reference stack is already loaded with the other vinum_lib")
        return _LOADED[1]
    if not _ZIP.exists():
        raise RuntimeError(f"{_ZIP} is missing: run oracle/build_ref.sh where /root/reference exists")
    root = str(_HERE.parent)
    if root not in sys.path:
        sys.path.insert(0, root)
    sys.path.insert(0, str(_ZIP))
    import vinum_b200.compat as compat
    if gpu_operators:
        compat.install(use_gpu_operators=True, use_parser=True)
    else:
        from . import ref
        lib = ref.ref_lib()
        if lib is None:
            raise RuntimeError("oracle/_ref (compiled reference operators) is not built")
        sys.modules["vinum_lib"] = lib
        compat.install(use_gpu_operators=False, use_parser=True)
    import vinum
    _LOADED = (gpu_operators, vinum)
    return vinum
