"""CPU tests of the SQL front end: the parser (stand-in for the reference's pglast parser,
vinum/parser/parser.py) yields the reference's tree shapes (modelled on
vinum/tests/test_sql_syntax_tree.py) and the output column names of every parity case equal
the names the REFERENCE produced (tests/golden/sql/manifest.json)."""
import importlib.util
import json
import sys
import types
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(Path(__file__).resolve().parent))

# the parser / AST are pure Python: load them without importing the CUDA library
pkg = types.ModuleType("vb_sql_t")
pkg.__path__ = [str(ROOT / "vinum_b200" / "sql")]
sys.modules.setdefault("vb_sql_t", pkg)


def _load(name):
    full = f"vb_sql_t.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, ROOT / "vinum_b200" / "sql" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


A = _load("ast")
P = _load("parser")
import sql_cases  # noqa: E402

COLS = ["id", "tax", "tip", "total", "lat", "vendor_id", "city"]


def parse(sql):
    return P.parse_sql(sql, COLS)


def test_select_star_expands_schema():
    q = parse("select * from tbl")
    assert [c.name for c in q.select] == COLS and q.where is None


def test_aliases_and_expression():
    q = parse("select tip as total, tax + tip as with_tip from t")
    assert isinstance(q.select[0], A.Column) and q.select[0].name == "tip" and q.select[0].alias == "total"
    e = q.select[1]
    assert e.op == A.Op.ADDITION and e.alias == "with_tip" and [a.name for a in e.args] == ["tax", "tip"]


def test_functions_and_count_star():
    q = parse("select np.power(10, np.min(total)) as exp, count(*) as cnt from t")
    e = q.select[0]
    assert e.op == A.Op.FUNCTION and e.function_name == "np.power" and e.alias == "exp"
    assert e.args[0].value == 10 and e.args[1].function_name == "np.min" and e.args[1].args[0].name == "total"
    c = q.select[1]
    assert c.function_name == "count_star" and c.args == () and c.alias == "cnt"


def test_where_and_is_nary_with_precedence():
    q = parse("select * from tbl where vendor_id > 1 and lat < 4.5 and tip = 0 or tax <> 1")
    w = q.where
    assert w.op == A.Op.OR and len(w.args) == 2
    assert w.args[0].op == A.Op.AND and len(w.args[0].args) == 3          # BoolExpr is n-ary
    assert w.args[0].args[1].op == A.Op.LESS_THAN and w.args[0].args[1].args[1].value == 4.5
    assert w.args[1].op == A.Op.NOT_EQUALS


def test_parentheses_and_arithmetic_precedence():
    q = parse("select (tax + tip) * 2 - total / 4 % 3, -tip, -5, ~id & 3 | 1, ~id + 1, tax | ~tip + 1 from t")
    e = q.select[0]
    assert e.op == A.Op.SUBTRACTION and e.args[0].op == A.Op.MULTIPLICATION and e.args[0].args[0].op == A.Op.ADDITION
    assert e.args[1].op == A.Op.MODULUS and e.args[1].args[0].op == A.Op.DIVISION
    assert q.select[1].op == A.Op.NEGATION
    assert isinstance(q.select[2], A.Literal) and q.select[2].value == -5   # sign folded into the constant
    # prefix ~ is an "other" operator (PostgreSQL): left associative among | & #, looser than + - * /
    e = q.select[3]
    assert e.op == A.Op.BINARY_OR and e.args[0].op == A.Op.BINARY_AND and e.args[0].args[0].op == A.Op.BINARY_NOT
    assert q.select[4].op == A.Op.BINARY_NOT and q.select[4].args[0].op == A.Op.ADDITION
    e = q.select[5]
    assert e.op == A.Op.BINARY_OR and e.args[1].op == A.Op.BINARY_NOT and e.args[1].args[0].op == A.Op.ADDITION


def test_null_tests_in_between_like():
    q = parse("select id from t where tip = null and tax != NULL and city is not null and id in (1,2,3) "
              "and tax not between 1 and 2 and city like 'a%' and city not in ('x','y')")
    ops = [a.op for a in q.where.args]
    assert ops == [A.Op.IS_NULL, A.Op.IS_NOT_NULL, A.Op.IS_NOT_NULL, A.Op.IN, A.Op.NOT_BETWEEN, A.Op.LIKE, A.Op.NOT_IN]
    assert q.where.args[3].args[1].value == [1, 2, 3]
    assert [a.value for a in q.where.args[4].args[1:]] == [1, 2]
    assert q.where.args[6].args[1].value == ["x", "y"]


def test_group_having_order_limit():
    q = parse("select city, sum(tip) s from t where tax > 0 group by city having sum(tip) > 10 "
              "order by s desc, city limit 5 offset 2")
    assert q.has_group_clause and [g.name for g in q.group_by] == ["city"]
    assert q.having.op == A.Op.GREATER_THAN
    assert [o.name for o in q.order_by] == ["s", "city"] and [s.name for s in q.sort_order] == ["DESC", "ASC"]
    assert (q.limit, q.offset) == (5, 2)
    assert q.select[1].alias == "s"
    q = parse("select distinct city from t offset 3")
    assert q.distinct and q.limit is None and q.offset == 0     # OFFSET is only read next to LIMIT


def test_non_reserved_keywords_are_column_names():
    """BY, FIRST, LAST, NULLS are non-reserved in the PostgreSQL grammar the reference parses with (pglast):
    `select first from t` names a column there; the clause keywords keep their meaning in clause position."""
    q = P.parse_sql("select first, last + 1 as l, nulls from t where by > 2 group by first order by last desc nulls first",
                    ["first", "last", "nulls", "by"])
    assert [getattr(c, "name", None) for c in q.select] == ["first", None, "nulls"]
    assert q.select[1].op == A.Op.ADDITION and q.select[1].args[0].name == "last" and q.select[1].alias == "l"
    assert q.where.op == A.Op.GREATER_THAN and q.where.args[0].name == "by"
    assert [g.name for g in q.group_by] == ["first"]
    assert [o.name for o in q.order_by] == ["last"] and [s.name for s in q.sort_order] == ["DESC"]


def test_string_literal_quotes_and_concat():
    q = parse("select city || '_' || 'it''s' from t")
    e = q.select[0]
    assert e.op == A.Op.CONCAT and e.args[1].value == "it's" and e.args[0].op == A.Op.CONCAT


@pytest.mark.parametrize("sql", ["update t set a = 1", "select from", "select a from t where", "select a, from t",
                                 "select a from t limit x", "select a from t order a", "select (a from t"])
def test_errors(sql):
    with pytest.raises(P.ParserError):
        parse(sql)


def test_every_parity_case_parses_and_names_match_reference():
    manifest = json.loads((ROOT / "tests" / "golden" / "sql" / "manifest.json").read_text())
    eng_names = None
    for entry in manifest:
        table = sql_cases.TABLES[entry["table"]]()
        q = P.parse_sql(entry["sql"], table.schema.names)
        assert entry["sql"] == sql_cases.CASES[entry["id"]][1], "manifest is stale: rerun oracle/gen_sql_golden.py"
        # QueryPlanner._column_names (planner.py:290-323), restated in engine.output_names
        out, index, unnamed = [], {}, 0
        for e in q.select:
            name = e.output_name()
            if not name:
                name = f"col_{unnamed}"
                unnamed += 1
            if name in index:
                index[name] += 1
                name = f"{name}_{index[name]}"
            else:
                index[name] = 0
            out.append(name)
        assert out == entry["columns"], entry["sql"]


# ---------------------------------------------------------------------------------------------
# Property: printing a random expression tree with the MINIMAL parentheses PostgreSQL's precedence
# table requires and parsing it back yields the same tree (precedence and associativity of every
# operator level, sign folding, n-ary AND / OR).
from hypothesis import given, settings, strategies as st  # noqa: E402

OR_, AND_, NOT_, IS_, CMP_, RANGE_, OTHER_, ADD_, MUL_, UNARY_, ATOM_ = range(1, 12)
_BIN_LEVEL = {A.Op.MULTIPLICATION: MUL_, A.Op.DIVISION: MUL_, A.Op.MODULUS: MUL_, A.Op.ADDITION: ADD_,
              A.Op.SUBTRACTION: ADD_, A.Op.BINARY_OR: OTHER_, A.Op.BINARY_AND: OTHER_, A.Op.BINARY_XOR: OTHER_,
              A.Op.CONCAT: OTHER_}
_CMP_OPS = [A.Op.EQUALS, A.Op.NOT_EQUALS, A.Op.GREATER_THAN, A.Op.GREATER_THAN_OR_EQUAL, A.Op.LESS_THAN,
            A.Op.LESS_THAN_OR_EQUAL]
_SYM = {A.Op.EQUALS: "=", A.Op.NOT_EQUALS: "<>"}


def _lit(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, str):
        return "'" + v.replace("'", "''") + "'"
    return repr(v)


def _level(n):
    if isinstance(n, A.Literal):
        return UNARY_ if isinstance(n.value, (int, float)) and not isinstance(n.value, bool) and n.value < 0 else ATOM_
    if isinstance(n, A.Column) or n.op == A.Op.FUNCTION:
        return ATOM_
    if n.op in _BIN_LEVEL:
        return _BIN_LEVEL[n.op]
    return {A.Op.NEGATION: UNARY_, A.Op.BINARY_NOT: OTHER_, A.Op.OR: OR_, A.Op.AND: AND_, A.Op.NOT: NOT_,
            A.Op.IS_NULL: IS_, A.Op.IS_NOT_NULL: IS_, A.Op.BETWEEN: RANGE_, A.Op.NOT_BETWEEN: RANGE_, A.Op.IN: RANGE_,
            A.Op.NOT_IN: RANGE_, A.Op.LIKE: RANGE_, A.Op.NOT_LIKE: RANGE_}.get(n.op, CMP_)


def _p(n, need):
    text = _print(n)
    return f"({text})" if _level(n) < need else text


def _print(n):
    if isinstance(n, A.Literal):
        return _lit(n.value)
    if isinstance(n, A.Column):
        return n.name
    op, a = n.op, n.args
    if op == A.Op.FUNCTION:
        return f"{n.function_name}({', '.join(_print(x) for x in a)})"
    if op in _BIN_LEVEL:
        lv = _BIN_LEVEL[op]
        return f"{_p(a[0], lv)} {op.value} {_p(a[1], lv + 1)}"
    if op == A.Op.NEGATION:
        return f"- {_p(a[0], UNARY_)}"
    if op == A.Op.BINARY_NOT:
        return f"~ {_p(a[0], ADD_)}"
    if op in (A.Op.OR, A.Op.AND):
        lv = OR_ if op == A.Op.OR else AND_
        return f" {op.value} ".join(_p(x, lv + 1) for x in a)
    if op == A.Op.NOT:
        return f"not {_p(a[0], NOT_)}"
    if op in (A.Op.IS_NULL, A.Op.IS_NOT_NULL):
        return f"{_p(a[0], CMP_)} {op.value}"
    if op in (A.Op.BETWEEN, A.Op.NOT_BETWEEN):
        return f"{_p(a[0], OTHER_)} {op.value} {_p(a[1], OTHER_)} and {_p(a[2], OTHER_)}"
    if op in (A.Op.IN, A.Op.NOT_IN):
        return f"{_p(a[0], OTHER_)} {op.value} ({', '.join(_lit(v) for v in a[1].value)})"
    if op in (A.Op.LIKE, A.Op.NOT_LIKE):
        return f"{_p(a[0], OTHER_)} {op.value} {_p(a[1], OTHER_)}"
    return f"{_p(a[0], RANGE_)} {_SYM.get(op, op.value)} {_p(a[1], RANGE_)}"   # comparisons: non-associative


_numbers = st.one_of(st.integers(-10**12, 10**12), st.floats(-1e9, 1e9, allow_nan=False, allow_infinity=False))
_strings = st.text(alphabet="abc xyz'%_-", max_size=6)
_leaves = st.one_of(_numbers.map(A.Literal), _strings.map(A.Literal), st.booleans().map(A.Literal),
                    st.sampled_from(COLS).map(A.Column))


def _not_numeric_literal(n):
    return not (isinstance(n, A.Literal) and isinstance(n.value, (int, float)) and not isinstance(n.value, bool))


def _extend(children):
    two = st.tuples(children, children)
    return st.one_of(
        st.tuples(st.sampled_from(list(_BIN_LEVEL)), two).map(lambda t: A.Expression(t[0], t[1])),
        st.tuples(st.sampled_from(_CMP_OPS), two).map(lambda t: A.Expression(t[0], t[1])),
        children.filter(_not_numeric_literal).map(lambda c: A.Expression(A.Op.NEGATION, (c,))),   # -<number> folds
        children.map(lambda c: A.Expression(A.Op.BINARY_NOT, (c,))),
        children.map(lambda c: A.Expression(A.Op.NOT, (c,))),
        st.tuples(st.sampled_from([A.Op.IS_NULL, A.Op.IS_NOT_NULL]), children).map(lambda t: A.Expression(t[0], (t[1],))),
        st.tuples(st.sampled_from([A.Op.AND, A.Op.OR]), st.lists(children, min_size=2, max_size=4)).map(
            lambda t: A.Expression(t[0], tuple(t[1]))),
        st.tuples(st.sampled_from([A.Op.BETWEEN, A.Op.NOT_BETWEEN]), children, children, children).map(
            lambda t: A.Expression(t[0], t[1:])),
        st.tuples(st.sampled_from([A.Op.IN, A.Op.NOT_IN]), children,
                  st.lists(st.one_of(st.integers(-99, 99), _strings), min_size=1, max_size=4)).map(
            lambda t: A.Expression(t[0], (t[1], A.Literal(t[2])))),
        st.tuples(st.sampled_from([A.Op.LIKE, A.Op.NOT_LIKE]), children, _strings.map(A.Literal)).map(
            lambda t: A.Expression(t[0], t[1:])),
        st.tuples(st.sampled_from(["f", "np.sin", "to_int", "sum"]), st.lists(children, min_size=1, max_size=3)).map(
            lambda t: A.Expression(A.Op.FUNCTION, tuple(t[1]), function_name=t[0])),
    )


@settings(max_examples=400, deadline=None)
@given(st.recursive(_leaves, _extend, max_leaves=12))
def test_print_parse_round_trip(tree):
    sql = "select " + _print(tree) + " from t"
    got = P.parse_sql(sql, COLS).select[0]
    assert got.key() == tree.key(), sql
