"""Golden results for the SQL parity cases (tests/sql_cases.py), produced by the REFERENCE:
each query is parsed by this repo's parser (the reference's own parser needs the pglast C
extension, which is not installable here), the tree is mapped node for node onto the
reference's AST classes (vinum/parser/query.py) and executed by the reference's own
QueryPlanner + RecursiveExecutor over its compiled C++ operators (oracle/_ref).

TEST INFRASTRUCTURE ONLY.  Needs /root/reference; run here, commit tests/golden/sql/.
    python oracle/gen_sql_golden.py
"""
import json
import sys
from pathlib import Path

import pyarrow as pa

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "oracle" / "stubs"))
sys.path.insert(0, str(REFERENCE))

from oracle import ref  # noqa: E402

sys.modules["vinum_lib"] = ref.ref_lib()
import vinum  # noqa: E402,F401
from vinum.arrow.arrow_table import ArrowTable  # noqa: E402
from vinum.executor.executor import RecursiveExecutor  # noqa: E402
from vinum.parser import query as rq  # noqa: E402
from vinum.planner.planner import QueryPlanner  # noqa: E402

import importlib.util  # noqa: E402

# the parser is pure Python: load it without importing the CUDA library
def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod

import types  # noqa: E402
pkg = types.ModuleType("vb_sql")
pkg.__path__ = [str(ROOT / "vinum_b200" / "sql")]
sys.modules["vb_sql"] = pkg
ast_mod = _load("vb_sql.ast", ROOT / "vinum_b200" / "sql" / "ast.py")
parser_mod = _load("vb_sql.parser", ROOT / "vinum_b200" / "sql" / "parser.py")

import sql_cases  # noqa: E402


def to_ref(node):
    if node is None:
        return None
    if isinstance(node, ast_mod.Literal):
        return rq.Literal(node.value, node.alias)
    if isinstance(node, ast_mod.Column):
        return rq.Column(node.name, node.alias)
    return rq.Expression(rq.SQLExpression[node.op.name], tuple(to_ref(a) for a in node.args),
                         function_name=node.function_name, alias=node.alias)


def run_reference(sql: str, table: pa.Table) -> pa.Table:
    q = parser_mod.parse_sql(sql, table.schema.names)
    rquery = rq.Query(table.schema, tuple(to_ref(e) for e in q.select), bool(q.distinct or q.has_group_clause),
                      q.distinct, to_ref(q.where), tuple(to_ref(g) for g in q.group_by), to_ref(q.having),
                      tuple(to_ref(o) for o in q.order_by),
                      tuple(rq.SortOrder.DESC if s.name == "DESC" else rq.SortOrder.ASC for s in q.sort_order),
                      q.limit, q.offset)
    at = ArrowTable(table)
    dag = QueryPlanner(rquery, at).plan_query()
    return RecursiveExecutor().execute(dag).get_table()


def main():
    out_dir = ROOT / "tests" / "golden" / "sql"
    out_dir.mkdir(parents=True, exist_ok=True)
    tables = {k: f() for k, f in sql_cases.TABLES.items()}
    manifest = []
    for i, (tname, sql, ordered) in enumerate(sql_cases.CASES):
        entry = {"id": i, "table": tname, "sql": sql, "ordered": ordered}
        try:
            res = run_reference(sql, tables[tname]).combine_chunks()
            path = out_dir / f"case_{i:03d}.arrow"
            with pa.OSFile(str(path), "wb") as f, pa.ipc.new_file(f, res.schema) as w:
                w.write_table(res)
            entry["file"] = path.name
            entry["rows"] = res.num_rows
            entry["columns"] = res.schema.names
        except Exception as e:  # noqa: BLE001
            entry["reference_error"] = f"{type(e).__name__}: {e}"[:300]
        manifest.append(entry)
        print(i, entry.get("rows"), entry.get("reference_error", ""), sql[:70])
    (out_dir / "manifest.json").write_text(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
