"""Harvest (input table, SQL, result) triples from the REFERENCE's own test suite.

Runs vinum/tests/test_query_results.py + test_table_api.py exactly like
oracle/run_reference_tests.py (reference planner / executor / compiled operators; this repo's
parser in place of pglast) with `Table.sql` wrapped so that every query the reference's tests
execute is recorded together with its input table and the reference's result (or the exception
type it raised).  The triples are written to tests/golden/ref_suite/ and replayed on the GPU by
tests/test_sql_gpu.py::test_reference_suite_queries.  TEST INFRASTRUCTURE ONLY.
    python oracle/harvest_reference_suite.py
"""
import hashlib
import json
import sys
from pathlib import Path

import pyarrow as pa

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")
OUT = ROOT / "tests" / "golden" / "ref_suite"

sys.argv = [sys.argv[0], "--collect-only", "-q", str(REFERENCE / "vinum" / "tests" / "test_table_api.py")]
try:
    exec(compile((ROOT / "oracle" / "run_reference_tests.py").read_text(), "run_reference_tests", "exec"))
except SystemExit:
    pass

import pytest  # noqa: E402
import vinum  # noqa: E402
from vinum.api.table import Table  # noqa: E402

records = []
_orig_sql = Table.sql


def _recording_sql(self, query):
    src = self._arrow_table.get_table()
    try:
        res = _orig_sql(self, query)
    except Exception as e:  # noqa: BLE001
        records.append((src, query, None, type(e).__name__))
        raise
    records.append((src, query, res._arrow_table.get_table(), None))
    return res


Table.sql = _recording_sql
rc = pytest.main(["-p", "no:cacheprovider", "-q", "--rootdir", "/tmp",
                  str(REFERENCE / "vinum" / "tests" / "test_query_results.py"),
                  str(REFERENCE / "vinum" / "tests" / "test_table_api.py")])
assert rc == 0, rc


def write(table: pa.Table, path: Path):
    table = table.combine_chunks()
    with pa.OSFile(str(path), "wb") as f, pa.ipc.new_file(f, table.schema) as w:
        w.write_table(table)


def digest(table: pa.Table) -> str:
    sink = pa.BufferOutputStream()
    with pa.ipc.new_stream(sink, table.schema) as w:
        w.write_table(table.combine_chunks())
    return hashlib.sha1(sink.getvalue().to_pybytes()).hexdigest()[:12]


OUT.mkdir(parents=True, exist_ok=True)
for old in OUT.glob("*.arrow"):
    old.unlink()
tables = {}
manifest = []
seen = set()
for src, query, res, err in records:
    d = digest(src)
    if d not in tables:
        tables[d] = f"table_{len(tables):02d}.arrow"
        write(src, OUT / tables[d])
    key = (d, query)
    if key in seen:
        continue
    seen.add(key)
    entry = {"id": len(manifest), "table": tables[d], "sql": query}
    if err is not None:
        entry["error"] = err
    else:
        entry["result"] = f"res_{len(manifest):03d}.arrow"
        write(res, OUT / entry["result"])
    manifest.append(entry)
(OUT / "manifest.json").write_text(json.dumps(manifest, indent=1))
print(f"harvested {len(manifest)} distinct queries over {len(tables)} tables "
      f"({sum(1 for m in manifest if 'error' in m)} expected errors)")
