"""Recursive-descent parser for the SELECT dialect the reference accepts.

Stand-in for `PglastParser` (vinum/parser/parser.py:55-289), which wraps the PostgreSQL
grammar through pglast -- a C extension that is not installable in this environment.  The
accepted statement is the one documented in doc/source/select.rst:7-13:

    SELECT [DISTINCT] expr [[AS] alias], ... FROM name
        [WHERE expr] [GROUP BY expr, ...] [HAVING expr]
        [ORDER BY expr [ASC|DESC], ...] [LIMIT n [OFFSET m]]

with PostgreSQL operator precedence and the operator table of parser.py:61-88.  The tree
that comes out has the reference's shape: n-ary AND / OR (`BoolExpr`), `x = NULL` rewritten
to IS NULL (parser.py:141-146), `count(*)` named `count_star` (parser.py:206-207), a negative
numeric literal folded into the constant, `*` expanded against the schema (parser.py:126-130).
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence

from .ast import Column, Expression, Literal, Node, Op, Query, SortOrder


class ParserError(Exception):
    """vinum/errors/__init__.py: ParserError."""


_KEYWORDS = {"select", "distinct", "from", "where", "group", "by", "having", "order", "asc", "desc", "limit",
             "offset", "and", "or", "not", "is", "null", "in", "between", "like", "as", "true", "false",
             "nulls", "first", "last"}

_NON_RESERVED = {"by", "first", "last", "nulls"}

_TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|--[^\n]*)
  | (?P<number>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?)
  | (?P<string>'(?:[^']|'')*')
  | (?P<qident>"(?:[^"]|"")*")
  | (?P<ident>[A-Za-z_][A-Za-z_0-9]*(?:\.[A-Za-z_][A-Za-z_0-9]*)*)
  | (?P<op><>|!=|>=|<=|==|\|\||[-+*/%=<>|&\#~(),;])
""", re.VERBOSE)


class _Tok:
    __slots__ = ("kind", "text", "pos")

    def __init__(self, kind, text, pos):
        self.kind, self.text, self.pos = kind, text, pos

    def __repr__(self):
        return f"{self.kind}:{self.text}"


def tokenize(sql: str) -> List[_Tok]:
    out, pos = [], 0
    while pos < len(sql):
        m = _TOKEN_RE.match(sql, pos)
        if not m:
            raise ParserError(f"Failed to parse the query: unexpected character {sql[pos]!r} at {pos}.")
        kind = m.lastgroup
        text = m.group(kind)
        if kind == "ident" and text.lower() in _KEYWORDS:
            out.append(_Tok("kw", text.lower(), pos))
        elif kind != "ws":
            out.append(_Tok(kind, text, pos))
        pos = m.end()
    out.append(_Tok("eof", "", len(sql)))
    return out


_CMP = {"=": Op.EQUALS, "==": Op.EQUALS, "!=": Op.NOT_EQUALS, "<>": Op.NOT_EQUALS, ">": Op.GREATER_THAN,
        ">=": Op.GREATER_THAN_OR_EQUAL, "<": Op.LESS_THAN, "<=": Op.LESS_THAN_OR_EQUAL}
_OTHER = {"|": Op.BINARY_OR, "&": Op.BINARY_AND, "#": Op.BINARY_XOR, "||": Op.CONCAT}
_ADD = {"+": Op.ADDITION, "-": Op.SUBTRACTION}
_MUL = {"*": Op.MULTIPLICATION, "/": Op.DIVISION, "%": Op.MODULUS}


class Parser:
    def __init__(self, sql: str, column_names: Optional[Sequence[str]] = None):
        self.sql = sql
        self.columns = list(column_names) if column_names is not None else None
        self.toks = tokenize(sql)
        self.i = 0

    # ------------------------------------------------------------ helpers
    @property
    def tok(self) -> _Tok:
        return self.toks[self.i]

    def _advance(self) -> _Tok:
        t = self.toks[self.i]
        self.i += 1
        return t

    def _is_kw(self, *words) -> bool:
        return self.tok.kind == "kw" and self.tok.text in words

    def _is_op(self, *ops) -> bool:
        return self.tok.kind == "op" and self.tok.text in ops

    def _accept_kw(self, word) -> bool:
        if self._is_kw(word):
            self.i += 1
            return True
        return False

    def _accept_op(self, op) -> bool:
        if self._is_op(op):
            self.i += 1
            return True
        return False

    def _expect_kw(self, word):
        if not self._accept_kw(word):
            raise ParserError(f"Failed to parse the query: expected {word.upper()} near position {self.tok.pos}.")

    def _expect_op(self, op):
        if not self._accept_op(op):
            raise ParserError(f"Failed to parse the query: expected '{op}' near position {self.tok.pos}.")

    # ---------------------------------------------------------- statement
    def parse(self) -> Query:
        if not self._is_kw("select"):
            raise ParserError("Only SELECT statements are supported.")
        self._advance()
        distinct = self._accept_kw("distinct")
        select: List[Node] = []
        while True:
            select.extend(self._select_item())
            if not self._accept_op(","):
                break
        if self._accept_kw("from"):
            if self.tok.kind not in ("ident", "qident"):
                raise ParserError("Failed to parse the query: table name expected after FROM.")
            self._advance()
            if self.tok.kind == "ident":  # table alias
                self._advance()
        where = self._expr() if self._accept_kw("where") else None
        group_by: List[Node] = []
        has_group = False
        if self._accept_kw("group"):
            self._expect_kw("by")
            has_group = True
            while True:
                group_by.append(self._expr())
                if not self._accept_op(","):
                    break
        having = self._expr() if self._accept_kw("having") else None
        order_by: List[Node] = []
        sort_order: List[SortOrder] = []
        if self._accept_kw("order"):
            self._expect_kw("by")
            while True:
                order_by.append(self._expr())
                if self._accept_kw("desc"):
                    sort_order.append(SortOrder.DESC)
                else:
                    self._accept_kw("asc")
                    sort_order.append(SortOrder.ASC)
                if self._accept_kw("nulls"):
                    if not (self._accept_kw("first") or self._accept_kw("last")):
                        raise ParserError("Failed to parse the query: NULLS FIRST / NULLS LAST expected.")
                if not self._accept_op(","):
                    break
        limit, offset = None, 0
        seen_offset = None
        while self._is_kw("limit", "offset"):
            if self._accept_kw("limit"):
                limit = self._int_literal("LIMIT")
            else:
                self._advance()
                seen_offset = self._int_literal("OFFSET")
        if limit is not None and seen_offset is not None:  # OFFSET is only read next to a LIMIT (parser.py:268-272)
            offset = seen_offset
        self._accept_op(";")
        if self.tok.kind != "eof":
            raise ParserError(f"Failed to parse the query: unexpected '{self.tok.text}' at position {self.tok.pos}.")
        return Query(tuple(select), distinct, where, tuple(group_by), having, tuple(order_by), tuple(sort_order),
                     limit, offset, has_group_clause=has_group)

    def _int_literal(self, what: str) -> int:
        t = self._advance()
        if t.kind != "number" or not t.text.isdigit():
            raise ParserError(f"Failed to parse the query: {what} needs an integer.")
        return int(t.text)

    def _select_item(self) -> List[Node]:
        if self._is_op("*"):
            self._advance()
            if self.columns is None:
                raise ParserError("SELECT * needs the table schema.")
            return [Column(c) for c in self.columns]
        node = self._expr()
        alias = None
        if self._accept_kw("as"):
            t = self._advance()
            if t.kind not in ("ident", "qident", "kw"):
                raise ParserError("Failed to parse the query: alias expected after AS.")
            alias = _unquote(t)
        elif self.tok.kind in ("ident", "qident"):
            alias = _unquote(self._advance())
        if alias:
            node.alias = alias
        return [node]

    # --------------------------------------------------------- expressions
    def _expr(self) -> Node:
        return self._or()

    def _or(self) -> Node:
        args = [self._and()]
        while self._accept_kw("or"):
            args.append(self._and())
        return args[0] if len(args) == 1 else Expression(Op.OR, tuple(args))

    def _and(self) -> Node:
        args = [self._not()]
        while self._accept_kw("and"):
            args.append(self._not())
        return args[0] if len(args) == 1 else Expression(Op.AND, tuple(args))

    def _not(self) -> Node:
        if self._accept_kw("not"):
            return Expression(Op.NOT, (self._not(),))
        return self._is()

    def _is(self) -> Node:
        node = self._comparison()
        while self._is_kw("is"):
            self._advance()
            neg = self._accept_kw("not")
            self._expect_kw("null")
            node = Expression(Op.IS_NOT_NULL if neg else Op.IS_NULL, (node,))
        return node

    def _comparison(self) -> Node:
        left = self._range()
        if self.tok.kind == "op" and self.tok.text in _CMP:
            name = self._advance().text
            right = self._range()
            # `x = NULL` / `x != NULL` become null tests (parser.py:141-146)
            args = [a for a in (left, right) if not (isinstance(a, Literal) and a.value is None)]
            if len(args) < 2:
                if name in ("=", "=="):
                    return Expression(Op.IS_NULL, tuple(args))
                if name in ("!=", "<>"):
                    return Expression(Op.IS_NOT_NULL, tuple(args))
            return Expression(_CMP[name], (left, right))
        return left

    def _range(self) -> Node:
        left = self._other()
        neg = False
        save = self.i
        if self._is_kw("not"):
            self._advance()
            neg = True
        if self._accept_kw("between"):
            lo = self._other()
            self._expect_kw("and")
            hi = self._other()
            return Expression(Op.NOT_BETWEEN if neg else Op.BETWEEN, (left, lo, hi))
        if self._accept_kw("in"):
            self._expect_op("(")
            values = []
            while True:
                item = self._expr()
                if not isinstance(item, Literal):
                    raise ParserError("Failed to parse the query: IN needs a list of literals.")
                values.append(item.value)
                if not self._accept_op(","):
                    break
            self._expect_op(")")
            return Expression(Op.NOT_IN if neg else Op.IN, (left, Literal(values)))
        if self._accept_kw("like"):
            pattern = self._other()
            return Expression(Op.NOT_LIKE if neg else Op.LIKE, (left, pattern))
        self.i = save
        return left

    def _other(self) -> Node:
        """"Any other operator" level of PostgreSQL (| & # || and prefix ~), left associative:
        `~a | b` is `(~a) | b`, while `~a + b` is `~(a + b)` because + binds tighter."""
        node = self._other_operand()
        while self.tok.kind == "op" and self.tok.text in _OTHER:
            op = _OTHER[self._advance().text]
            node = Expression(op, (node, self._other_operand()))
        return node

    def _other_operand(self) -> Node:
        if self._accept_op("~"):
            return Expression(Op.BINARY_NOT, (self._other_operand(),))
        return self._additive()

    def _additive(self) -> Node:
        node = self._multiplicative()
        while self.tok.kind == "op" and self.tok.text in _ADD:
            op = _ADD[self._advance().text]
            node = Expression(op, (node, self._multiplicative()))
        return node

    def _multiplicative(self) -> Node:
        node = self._unary()
        while self.tok.kind == "op" and self.tok.text in _MUL:
            op = _MUL[self._advance().text]
            node = Expression(op, (node, self._unary()))
        return node

    def _unary(self) -> Node:
        if self._accept_op("-"):
            arg = self._unary()
            if isinstance(arg, Literal) and isinstance(arg.value, (int, float)) and not isinstance(arg.value, bool):
                return Literal(-arg.value)  # the PostgreSQL grammar folds the sign into the constant
            return Expression(Op.NEGATION, (arg,))
        if self._accept_op("+"):
            return self._unary()
        if self._accept_op("~"):
            return Expression(Op.BINARY_NOT, (self._unary(),))
        return self._atom()

    def _atom(self) -> Node:
        t = self.tok
        if t.kind == "number":
            self._advance()
            if re.fullmatch(r"\d+", t.text):
                return Literal(int(t.text))
            return Literal(float(t.text))
        if t.kind == "string":
            self._advance()
            return Literal(t.text[1:-1].replace("''", "'"))
        if t.kind == "kw" and t.text in ("true", "false"):
            self._advance()
            return Literal(t.text == "true")
        if t.kind == "kw" and t.text == "null":
            self._advance()
            return Literal(None)
        if t.kind == "op" and t.text == "(":
            self._advance()
            node = self._expr()
            self._expect_op(")")
            return node
        if t.kind == "qident":
            self._advance()
            return Column(_unquote(t))
        if t.kind == "kw" and t.text in _NON_RESERVED:
            # BY / FIRST / LAST / NULLS are non-reserved in PostgreSQL (the grammar pglast gives the reference,
            # vinum/parser/parser.py): where an operand is expected they are ordinary column names
            self._advance()
            return Column(t.text)
        if t.kind == "ident":
            self._advance()
            if self._accept_op("("):
                return self._call(t.text)
            name = t.text
            if "." in name and (self.columns is None or name not in self.columns):
                name = name.split(".")[-1]  # table-qualified column
            return Column(name)
        raise ParserError(f"Failed to parse the query: unexpected '{t.text or 'end of input'}' at position {t.pos}.")

    def _call(self, name: str) -> Node:
        if self._is_op("*"):
            self._advance()
            self._expect_op(")")
            if name.lower() == "count":
                return Expression(Op.FUNCTION, (), function_name="count_star")
            return Expression(Op.FUNCTION, (), function_name=name)
        args = []
        if not self._is_op(")"):
            self._accept_kw("distinct")
            while True:
                args.append(self._expr())
                if not self._accept_op(","):
                    break
        self._expect_op(")")
        return Expression(Op.FUNCTION, tuple(args), function_name=name)


def _unquote(t: _Tok) -> str:
    if t.kind == "qident":
        return t.text[1:-1].replace('""', '"')
    return t.text


def parse_sql(sql: str, column_names: Optional[Sequence[str]] = None) -> Query:
    """SQL text -> Query.  `column_names` expands `SELECT *` (the schema's column order)."""
    return Parser(sql, column_names).parse()
