"""The drop-in under the REFERENCE's own Python layers, on the GPU (-m gpu).

`vinum_b200.compat.install()` + the reference's unmodified `vinum` package (oracle/_ref/vinum_pyref.zip,
test infrastructure): `vn.Table.from_arrow(t).sql(...)` -- the reference's binder, planner, executor and
Python operators -- runs on `vinum_b200.vinum_lib` (aggregate / sort / batch reader classes) and, for
scan -> filter -> aggregate plans, on the fused device path the patched AggregateOperator dispatches to.
Every result is compared with the same query run by the reference on its OWN compiled operators
(oracle/_ref), in a second process (the reference binds `vinum_lib` at import time)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

QUERIES = [
    ("northstar", "SELECT i0, COUNT(*), SUM(f1) FROM t WHERE f0 > 0.5 GROUP BY i0 ORDER BY i0"),
    ("fused_min_max_avg", "SELECT i0, MIN(i1) AS mn, MAX(f1) AS mx, AVG(f1) AS av, COUNT(f0) AS c FROM t WHERE i1 <= 0 GROUP BY i0 ORDER BY i0"),
    ("no_where", "SELECT i0, SUM(i1) AS s FROM t GROUP BY i0 ORDER BY i0"),
    ("expr_where_general_path", "SELECT i0, COUNT(*) AS c FROM t WHERE f0 * 2 > 1 GROUP BY i0 ORDER BY i0"),
    ("string_key_generic", "SELECT s, COUNT(*) AS c, SUM(f1) AS sf FROM t GROUP BY s ORDER BY s"),
    ("order_by_with_string_payload", "SELECT s, i1 FROM t WHERE f0 > 0.99 ORDER BY i1 DESC, s LIMIT 25"),
    ("order_by_string_key", "SELECT s, f1 FROM t WHERE f0 > 0.995 ORDER BY s DESC, f1"),
    ("config0_plumbing", "SELECT f1 FROM t WHERE s = 'k3' LIMIT 10"),
]

_SCRIPT = r"""
import json, sys, time
sys.path.insert(0, {root!r})
import numpy as np, pyarrow as pa
from oracle import ref_stack
from vinum_b200 import datagen
gpu = {gpu!r}
vn = ref_stack.reference_vinum(gpu_operators=gpu)
n = {rows}
t = datagen.host_table(["i0", "i1", "f0", "f1"], 0, n)
t = t.append_column("s", pa.array(np.char.add("k", (datagen.host_column("i3", 0, n) % 7).astype(str))))
tbl = vn.Table.from_arrow(t)
out = {{}}
for name, q in {queries!r}:
    t0 = time.perf_counter()
    res = tbl.sql(q).to_arrow() if hasattr(tbl.sql("select i0 from t limit 1"), "to_arrow") else None
    out[name] = {{"seconds": time.perf_counter() - t0, "columns": res.column_names,
                 "data": {{c: res.column(c).to_pylist() for c in res.column_names}}}}
if gpu:
    import vinum_lib, vinum.core.aggregate as ra
    out["_meta"] = {{"vinum_lib": vinum_lib.__name__, "dispatch": bool(getattr(ra.AggregateOperator, "_vinum_b200_dispatch", False))}}
print("RESULT " + json.dumps(out))
"""


def _run(gpu: bool, rows: int):
    code = _SCRIPT.format(root=str(ROOT), gpu=gpu, rows=rows, queries=QUERIES)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_reference_python_layers_on_gpu_operators_match_reference_operators():
    from oracle import ref_stack
    if not ref_stack.available():
        pytest.skip("oracle/_ref (compiled reference operators + vinum_pyref.zip) is not built")
    rows = int(os.environ.get("VK_TEST_DROPIN_ROWS", 1_000_003))
    got = _run(True, rows)
    want = _run(False, rows)
    assert got["_meta"] == {"vinum_lib": "vinum_b200.vinum_lib", "dispatch": True}
    for name, _ in QUERIES:
        g, w = got[name], want[name]
        assert g["columns"] == w["columns"], name
        for c in g["columns"]:
            a, b = g["data"][c], w["data"][c]
            assert len(a) == len(b), (name, c)
            if a and isinstance(a[0], float):
                import numpy as np
                assert np.allclose(np.array(a, dtype=float), np.array(b, dtype=float), rtol=1e-6, atol=0, equal_nan=True), (name, c)
            else:
                assert a == b, (name, c)
