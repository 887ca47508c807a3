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
* `AggregateOperator.next` (the operator dispatch of vinum/core) -> the fused, streaming
  filter -> hash-aggregate path when the plan below it is scan -> [prune] -> [filter] -> aggregate over
  plain numeric columns (`_install_operator_dispatch`); every other plan is untouched;
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


def _install_operator_dispatch() -> None:
    """The operator dispatch of vinum/core routed to the fused device path (BASELINE.json north_star).

    `AggregateOperator.next` (vinum/core/aggregate.py:114-124) pulls 10 000-row batches through
    TableReaderOperator -> [ProjectOperator: column pruning] -> FilterOperator -> AggregateOperator, one
    NumPy comparison, one Arrow filter of every column and one `vinum_lib` call per batch.  When that
    chain is exactly this shape -- the WHERE is `column <cmp> numeric literal`, keys and aggregate
    arguments are plain null-free numeric columns of the table -- the patched `next` hands the WHOLE table
    to `vinum_b200.executor.filter_aggregate` (chunks copied host -> device on a copy stream while the
    fused filter -> hash-aggregate kernel consumes the previous one; no mask, no filtered batch) and yields
    the one RecordBatch `BaseAggregate::Result` would have produced (base_aggregate.cpp:47-68).  Any other
    plan runs the reference's own loop unchanged, on `vinum_b200.vinum_lib`'s aggregate classes."""
    import pyarrow as pa
    import vinum.core.aggregate as ra
    import vinum.core.algebra as alg
    import vinum.core.expressions as rex
    from vinum.arrow.record_batch import RecordBatch
    from vinum.core.base import VectorizedExpression
    from vinum.parser.query import Column, Literal, SQLExpression
    from . import executor

    if getattr(ra.AggregateOperator, "_vinum_b200_dispatch", False):
        return
    cmp_of = {id(rex.EXPRESSION_FUNCTIONS[k][0]): sym for k, sym in (
        (SQLExpression.EQUALS, "=="), (SQLExpression.NOT_EQUALS, "!="), (SQLExpression.GREATER_THAN, ">"),
        (SQLExpression.GREATER_THAN_OR_EQUAL, ">="), (SQLExpression.LESS_THAN, "<"), (SQLExpression.LESS_THAN_OR_EQUAL, "<="))}
    flip = {"==": "==", "!=": "!=", ">": "<", ">=": "<=", "<": ">", "<=": ">="}
    original_next = ra.AggregateOperator.next

    def match(op):
        """(table, where, group-by names, [(FUNC, column, out name)]) of a fusable plan, else None."""
        p = op._parent_operator
        where = None
        if isinstance(p, alg.FilterOperator):
            pred = p._arguments[0]
            if type(pred) is not VectorizedExpression or id(pred._function) not in cmp_of or len(pred._arguments) != 2:
                return None
            a, b = pred._arguments
            sym = cmp_of[id(pred._function)]
            if isinstance(a, Literal) and isinstance(b, Column):
                a, b, sym = b, a, flip[sym]
            if not (isinstance(a, Column) and isinstance(b, Literal)) or isinstance(b.value, bool) or \
                    not isinstance(b.value, (int, float)):
                return None
            where = (a.get_column_name(), sym, b.value)
            p = p._parent_operator
        if isinstance(p, alg.ProjectOperator):
            if p._keep_input_table or p._col_names or not all(isinstance(x, Column) for x in p._arguments):
                return None
            p = p._parent_operator
        if type(p) is not alg.TableReaderOperator:
            return None
        table = getattr(p._reader, "_table", None)
        if not isinstance(table, pa.Table):
            return None
        keys = [c.get_column_name() for c in op._group_by_columns]
        if not keys or len(set(keys)) != len(keys):
            return None
        funcs = []
        for f in op._agg_funcs:
            name = f.get_agg_func_name()
            if name not in ("COUNT_STAR", "COUNT", "MIN", "MAX", "SUM", "AVG"):
                return None
            funcs.append((name, f.get_input_column_name(), f.get_column_name()))
        if not funcs:
            return None
        for name in keys + [c for _, c, _ in funcs if c] + ([where[0]] if where else []):
            i = table.schema.get_field_index(name)
            if i < 0:
                return None
            col = table.column(i)
            t = col.type
            if col.null_count or not (pa.types.is_integer(t) or pa.types.is_floating(t)) or pa.types.is_float16(t):
                return None
        return table, where, keys, funcs

    def next(self):
        plan = match(self)
        if plan is None:
            yield from original_next(self)
            return
        table, where, keys, funcs = plan
        rb = executor.filter_aggregate(table, keys, funcs, where)
        names = [c.get_column_name() for c in self._agg_cols] + [o for _, _, o in funcs]
        arrays = [rb.column(rb.schema.get_field_index(n)) for n in names]   # agg_cols first (base_aggregate.cpp:100-118)
        self.fused_device_path = True
        yield RecordBatch(pa.RecordBatch.from_arrays(arrays, names=names))

    ra.AggregateOperator.next = next
    ra.AggregateOperator._vinum_b200_dispatch = True


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
    if use_gpu_operators:
        _install_operator_dispatch()
