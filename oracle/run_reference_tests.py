"""Run the REFERENCE's own pytest suite (vinum/tests/test_sql_syntax_tree.py and
test_query_results.py) with this repo's SQL parser standing in for the pglast one
(vinum_b200.compat) and the reference's own compiled C++ operators (oracle/_ref) as `vinum_lib`.
Everything except the parser is the reference: a pass means the stand-in parser produces trees the
reference's planner / executor / expected results accept.  TEST INFRASTRUCTURE ONLY; this
container only (needs /root/reference).
    python oracle/run_reference_tests.py [pytest args]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REFERENCE))

from oracle import ref  # noqa: E402

lib = ref.ref_lib()
assert lib is not None, "build oracle/_ref first (bash oracle/build_ref.sh)"
sys.modules["vinum_lib"] = lib

import importlib.util  # noqa: E402
import types  # noqa: E402

# load vinum_b200.compat and vinum_b200.sql without importing the CUDA library
pkg = types.ModuleType("vinum_b200")
pkg.__path__ = [str(ROOT / "vinum_b200")]
sys.modules["vinum_b200"] = pkg
sqlpkg = types.ModuleType("vinum_b200.sql")
sqlpkg.__path__ = [str(ROOT / "vinum_b200" / "sql")]
sys.modules["vinum_b200.sql"] = sqlpkg
for name, rel in (("vinum_b200.sql.ast", "sql/ast.py"), ("vinum_b200.sql.parser", "sql/parser.py"),
                  ("vinum_b200.compat", "compat.py")):
    spec = importlib.util.spec_from_file_location(name, ROOT / "vinum_b200" / rel)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
sys.modules["vinum_b200.compat"].install(use_gpu_operators=False, use_parser=True)

import pytest  # noqa: E402

args = sys.argv[1:] or [str(REFERENCE / "vinum" / "tests" / "test_sql_syntax_tree.py"),
                        str(REFERENCE / "vinum" / "tests" / "test_query_results.py")]
sys.exit(pytest.main(["-p", "no:cacheprovider", "-q", "--rootdir", "/tmp"] + args))
