"""The reference-side binding: make an importable copy of the reference (`import vinum`) run on
this package without touching its sources.

    import vinum_b200.compat as compat
    compat.install()          # before the first `import vinum`
    import vinum as vn
    vn.Table.from_pydict(...).sql_pd("SELECT ...")

`install()` does the two substitutions a maintainer would make:

* `vinum_lib` (the pybind11 extension, vinum/core/vinum_lib.cpp:20-167) -> `vinum_b200.vinum_lib`,
  looked up by module name, so `AggregateOperator` / `SortOperator` / `TableReaderOperator`
  (vinum/core/aggregate.py:96-124, vinum/core/algebra.py:150-177,250-265) drive the GPU operators;
* `parser_factory` (vinum/parser/parser.py:295-311) -> a parser object whose `.parse()` returns the
  reference's own `Query` tree, built by this package's recursive-descent parser -- the reference's
  parser needs the pglast C extension (pinned ==1.17); when pglast is not importable a stub
  module with the two names the reference imports from it is registered first.
"""
from __future__ import annotations

import sys
import types


def _ensure_pglast() -> None:
    try:
        import pglast  # noqa: F401
        import pglast.enums  # noqa: F401
        return
    except Exception:
        pass
    import enum
    pg = types.ModuleType("pglast")
    en = types.ModuleType("pglast.enums")

    class Node:  # placeholder type: the reference only names it
        pass

    def parse_sql(sql):
        raise RuntimeError("pglast is not installed: vinum_b200.compat routes parsing to vinum_b200.sql.parser")

    class A_Expr_Kind(enum.IntEnum):
        AEXPR_OP = 0
        AEXPR_IN = 7
        AEXPR_LIKE = 8
        AEXPR_BETWEEN = 11
        AEXPR_NOT_BETWEEN = 12

    class BoolExprType(enum.IntEnum):
        AND_EXPR = 0
        OR_EXPR = 1
        NOT_EXPR = 2

    pg.Node, pg.parse_sql, pg.enums = Node, parse_sql, en
    en.A_Expr_Kind, en.BoolExprType = A_Expr_Kind, BoolExprType
    sys.modules["pglast"] = pg
    sys.modules["pglast.enums"] = en


def to_reference_tree(node, rq):
    """This package's AST node -> the reference's (vinum/parser/query.py), one to one."""
    from .sql import ast as A
    if node is None:
        return None
    if isinstance(node, A.Literal):
        return rq.Literal(node.value, node.alias)
    if isinstance(node, A.Column):
        return rq.Column(node.name, node.alias)
    return rq.Expression(rq.SQLExpression[node.op.name], tuple(to_reference_tree(a, rq) for a in node.args),
                         function_name=node.function_name, alias=node.alias)


class _Parser:
    """Stands in for PglastParser (vinum/parser/parser.py:55-289): same constructor, same `.parse()`."""

    def __init__(self, sql: str, schema):
        self._sql = sql
        self._schema = schema

    def parse(self):
        from vinum.parser import query as rq
        from vinum.errors import ParserError as RefParserError
        from .sql.parser import ParserError, parse_sql
        try:
            q = parse_sql(self._sql, self._schema.names)
        except ParserError as e:
            raise RefParserError(str(e))
        conv = lambda n: to_reference_tree(n, rq)  # noqa: E731
        return rq.Query(self._schema, tuple(conv(e) for e in q.select), bool(q.distinct or q.has_group_clause),
                        q.distinct, conv(q.where), tuple(conv(g) for g in q.group_by), conv(q.having),
                        tuple(conv(o) for o in q.order_by),
                        tuple(rq.SortOrder.DESC if s.name == "DESC" else rq.SortOrder.ASC for s in q.sort_order),
                        q.limit, q.offset)


def install(use_gpu_operators: bool = True, use_parser: bool = True) -> None:
    """Register the substitutions.  Call before the first `import vinum`."""
    if use_gpu_operators:
        from . import vinum_lib
        sys.modules["vinum_lib"] = vinum_lib
    if use_parser:
        _ensure_pglast()
        import vinum.parser.parser as rp
        factory = lambda sql, schema: _Parser(sql, schema)  # noqa: E731
        rp.parser_factory = factory
        # modules that bound the name at import time (table.py:6, stream_reader.py:5)
        for name in ("vinum.api.table", "vinum.api.stream_reader"):
            mod = sys.modules.get(name)
            if mod is None:
                __import__(name)
                mod = sys.modules[name]
            mod.parser_factory = factory
