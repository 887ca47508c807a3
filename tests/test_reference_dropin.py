"""The drop-in on the REFERENCE itself (CPU, this container only): `import vinum` with
`vinum_b200.compat.install()` -- this package's `vinum_lib` module in place of the pybind11
extension and this package's parser in place of the pglast one -- runs BASELINE.json configs[0]
through the reference's unmodified Table / planner / executor.  The query has no aggregate and no
sort, so the only `vinum_lib` class on its path is TableBatchReader (host slicing): no GPU needed.
Skipped where /root/reference does not exist (the GPU box)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")

_SCRIPT = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {ref!r})
import vinum_b200.compat as compat
compat.install()
import vinum as vn
import vinum_lib
assert vinum_lib.__name__ == "vinum_b200.vinum_lib", vinum_lib.__name__
tbl = vn.Table.from_pydict({{'value': [300.1, 2.8, 880], 'mode': ['air', 'bus', 'air']}})
pdf = tbl.sql_pd("SELECT value FROM t WHERE mode='air'")
assert list(pdf.columns) == ['value'] and pdf['value'].tolist() == [300.1, 880.0], pdf
out = tbl.sql("select value * 2 as dbl, mode from t where value between 2 and 500 and mode in ('air', 'bus') limit 5").to_pandas()
assert out['dbl'].tolist() == [600.2, 5.6] and out['mode'].tolist() == ['air', 'bus'], out
try:
    tbl.sql("select nope from t")
except Exception as e:
    assert type(e).__name__ == "ParserError" and "not found" in str(e), (type(e), e)
else:
    raise AssertionError("missing column accepted")
print("dropin ok")
"""


@pytest.mark.skipif(not (REFERENCE / "vinum" / "__init__.py").exists(), reason="the reference checkout is not present")
def test_reference_runs_config0_on_our_vinum_lib_and_parser():
    env = dict(os.environ)
    r = subprocess.run([sys.executable, "-c", _SCRIPT.format(root=str(ROOT), ref=str(REFERENCE))],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "dropin ok" in r.stdout


@pytest.mark.skipif(not (REFERENCE / "vinum" / "tests" / "test_query_results.py").exists(),
                    reason="the reference checkout is not present")
def test_reference_own_suite_passes_with_our_parser():
    """The reference's OWN tests (vinum/tests/test_sql_syntax_tree.py, test_query_results.py,
    test_table_api.py: ~265 tests) with this repo's parser in place of the pglast one and the
    reference's compiled operators (oracle/_ref) for everything else: the stand-in parser yields
    trees the reference's planner, executor and expected results accept."""
    from oracle import ref
    if ref.ref_lib() is None:
        pytest.skip("oracle/_ref is not built")
    tests = [str(REFERENCE / "vinum" / "tests" / f) for f in
             ("test_sql_syntax_tree.py", "test_query_results.py", "test_table_api.py")]
    r = subprocess.run([sys.executable, str(ROOT / "oracle" / "run_reference_tests.py")] + tests,
                       capture_output=True, text=True, timeout=900, cwd="/tmp")
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    import re
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 250, tail
    assert "failed" not in r.stdout.splitlines()[-1], tail
