"""Query syntax tree of the SQL front end.

Same shape as the reference's AST (vinum/parser/query.py:13-432): a tree of `Literal`,
`Column` and `Expression(op, arguments, function_name)` nodes under a `Query` with SELECT /
WHERE / GROUP BY / HAVING / ORDER BY / LIMIT blocks, and the same operator vocabulary
(`Op` mirrors `SQLExpression`, query.py:13-56) -- so a maintainer can map one onto the
other node for node (oracle/gen_sql_golden.py does exactly that to produce the golden
results with the reference's own planner and executor).
"""
from __future__ import annotations

import enum
from typing import Any, Optional, Tuple, Union


class Op(enum.Enum):
    # arithmetic
    ADDITION = "+"
    SUBTRACTION = "-"
    MULTIPLICATION = "*"
    DIVISION = "/"
    MODULUS = "%"
    NEGATION = "neg"
    BINARY_NOT = "~"
    BINARY_AND = "&"
    BINARY_OR = "|"
    BINARY_XOR = "#"
    CONCAT = "||"
    # comparison
    EQUALS = "=="
    NOT_EQUALS = "!="
    GREATER_THAN = ">"
    GREATER_THAN_OR_EQUAL = ">="
    LESS_THAN = "<"
    LESS_THAN_OR_EQUAL = "<="
    # logical
    AND = "and"
    OR = "or"
    NOT = "not"
    BETWEEN = "between"
    NOT_BETWEEN = "not between"
    IN = "in"
    NOT_IN = "not in"
    LIKE = "like"
    NOT_LIKE = "not like"
    IS_NULL = "is null"
    IS_NOT_NULL = "is not null"
    # functions
    FUNCTION = "function"


class SortOrder(enum.Enum):
    ASC = 0
    DESC = 1


class Literal:
    __slots__ = ("value", "alias")

    def __init__(self, value: Any, alias: Optional[str] = None):
        self.value = value
        self.alias = alias

    def key(self):
        v = self.value
        return ("lit", type(v).__name__, tuple(v) if isinstance(v, list) else v)

    def output_name(self) -> Optional[str]:  # Literal.get_alias, query.py:152-153
        return self.alias

    def __repr__(self):
        return f"Literal({self.value!r})"


class Column:
    __slots__ = ("name", "alias")

    def __init__(self, name: str, alias: Optional[str] = None):
        assert name, "Column name is required."
        self.name = name
        self.alias = alias

    def key(self):
        return ("col", self.name)

    def output_name(self) -> Optional[str]:  # Column.get_alias, query.py:194-198
        return self.alias or self.name

    def __repr__(self):
        return f"Column({self.name!r})"


class Expression:
    __slots__ = ("op", "args", "function_name", "alias")

    def __init__(self, op: Op, args: Tuple["Node", ...], function_name: Optional[str] = None,
                 alias: Optional[str] = None):
        self.op = op
        self.args = tuple(args)
        self.function_name = function_name
        self.alias = alias

    def key(self):
        """Structural identity (alias excluded), Expression.__eq__ query.py:327-338."""
        return ("expr", self.op.name, (self.function_name or "").lower(), tuple(a.key() for a in self.args))

    def output_name(self) -> Optional[str]:  # Expression.get_alias, query.py:263-269
        if self.alias:
            return self.alias
        if self.op == Op.FUNCTION:
            return str(self.function_name)
        return None

    def __repr__(self):
        f = f":{self.function_name}" if self.function_name else ""
        return f"Expression({self.op.name}{f}, {list(self.args)!r})"


Node = Union[Literal, Column, Expression]

AGG_FUNCS = {"count_star", "count", "min", "max", "sum", "avg", "np.min", "np.max", "np.sum"}  # functions.py:390-400
NUMPY_AGG_MAPPING = {"np.min": "min", "np.max": "max", "np.sum": "sum"}                          # functions.py:402-406


def is_aggregate_call(node) -> bool:
    return (isinstance(node, Expression) and node.op == Op.FUNCTION and node.function_name is not None
            and node.function_name.lower() in AGG_FUNCS)


def contains_aggregate(node) -> bool:
    if is_aggregate_call(node):
        return True
    if isinstance(node, Expression):
        return any(contains_aggregate(a) for a in node.args)
    return False


def walk(node):
    yield node
    if isinstance(node, Expression):
        for a in node.args:
            yield from walk(a)


class Query:
    """SELECT statement (Query, vinum/parser/query.py:341-432)."""

    def __init__(self, select: Tuple[Node, ...], distinct: bool = False, where: Optional[Node] = None,
                 group_by: Tuple[Node, ...] = (), having: Optional[Node] = None, order_by: Tuple[Node, ...] = (),
                 sort_order: Tuple[SortOrder, ...] = (), limit: Optional[int] = None, offset: int = 0,
                 has_group_clause: bool = False):
        self.select = tuple(select)
        self.distinct = distinct
        self.where = where
        self.group_by = tuple(group_by)
        self.having = having
        self.order_by = tuple(order_by)
        self.sort_order = tuple(sort_order)
        self.limit = limit
        self.offset = offset
        self.has_group_clause = has_group_clause

    def __repr__(self):
        return (f"Query(select={list(self.select)!r}, distinct={self.distinct}, where={self.where!r}, "
                f"group_by={list(self.group_by)!r}, having={self.having!r}, order_by={list(self.order_by)!r}, "
                f"sort_order={[s.name for s in self.sort_order]}, limit={self.limit}, offset={self.offset})")
