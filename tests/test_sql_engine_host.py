"""CPU tests of the query engine's host logic (no kernel is launched): binding and its error
messages (Binder, vinum/planner/binder.py:41-265), output column naming
(QueryPlanner._column_names, planner.py:290-323), the choice of the streaming aggregate path,
dictionary coding of string keys across batches, and the host function registry
(vinum/core/functions.py:341-367)."""
import numpy as np
import pyarrow as pa
import pytest

from vinum_b200.sql import engine as E
from vinum_b200.sql import functions as F
from vinum_b200.sql.parser import ParserError, parse_sql


def _table():
    return pa.table({"k": pa.array([1, 2, 1], type=pa.int64()), "v": [1.0, 2.0, 3.0], "s": ["a", "b", None],
                     "n": pa.array([1, None, 3], type=pa.int64()), "b": [True, False, True]})


def _bind(sql):
    t = _table()
    eng = E.Engine(t)
    return eng, eng._bind(parse_sql(sql, t.schema.names))


def test_output_names_follow_the_reference_rules():
    q = parse_sql("select k, k, v + 1, v * 2, sum(v), sum(v) as total, count(*), 3 as three, np.sin(v) from t", ["k", "v"])
    assert E.output_names(q.select) == ["k", "k_1", "col_0", "col_1", "sum", "total", "count_star", "three", "np.sin"]


def test_alias_substitution_and_aggregate_detection():
    _, q = _bind("select k % 2 as bucket, count(*) c from t group by bucket having c > 1 order by bucket")
    assert q.is_aggregate
    assert q.group_by[0].key() == q.select[0].key()              # alias replaced by its expression
    assert q.having.args[0].function_name == "count_star"
    _, q = _bind("select k from t")
    assert not q.is_aggregate
    _, q = _bind("select sum(v) from t")
    assert q.is_aggregate and q.group_by == ()


@pytest.mark.parametrize("sql, fragment", [
    ("select nope from t", "Column 'nope' is not found."),
    ("select k from t where nope > 1", "Column 'nope' is not found."),
    ("select k, v from t group by k", 'Column "v" is not part of the "GROUP BY" clause'),
    ("select k, v + 1 from t group by k", "is neither aggregate function nor part of the"),
    ("select k, 4 from t group by k", "is not allowed in the aggregate query"),
])
def test_binder_errors_use_the_reference_messages(sql, fragment):
    with pytest.raises(ParserError) as e:
        _bind(sql)
    assert fragment in str(e.value)


@pytest.mark.parametrize("sql, streamed", [
    ("select k, count(*), sum(v) from t where v > 0.5 group by k", True),
    ("select k, sum(v) s from t where 2 < v group by k having s > 1 order by s desc limit 3", True),
    ("select k, sum(v) from t group by k", True),
    ("select k, sum(v + 1) from t group by k", False),          # computed argument
    ("select k % 2, sum(v) from t group by k % 2", False),      # computed key
    ("select s, count(*) from t group by s", False),            # string key: dictionary path
    ("select n, count(*) from t group by n", False),            # NULLs in the key
    ("select k, sum(n) from t group by k", False),              # NULLs in the argument
    ("select k, sum(v) from t where v > 0.5 and k > 0 group by k", False),   # compound predicate -> mask
    ("select b, count(*) from t group by b", False),            # boolean key
    ("select distinct k from t", False),
    ("select sum(v) from t", False),                            # un-grouped
])
def test_streaming_aggregate_eligibility(sql, streamed):
    eng, q = _bind(sql)
    assert eng._streamable(q) is streamed


def test_string_key_dictionary_persists_across_batches():
    table, values = {}, []
    a = E._encode_with_dictionary(pa.array(["x", "y", None, "x"]), table, values)
    b = E._encode_with_dictionary(pa.chunked_array([pa.array(["z", "y"]), pa.array(["x", None])]), table, values)
    assert a.to_pylist() == [0, 1, None, 0] and b.to_pylist() == [2, 1, 0, None]
    assert values == ["x", "y", "z"] and a.type == pa.int32()
    c = E._encode_with_dictionary(pa.array([], type=pa.string()), table, values)
    assert len(c) == 0


def test_host_function_registry_matches_numpy_and_reference_casts():
    x = np.array([1.5, -2.25, 9.0])
    assert np.array_equal(F.call_host_function("abs", [x]), np.absolute(x))
    assert np.array_equal(F.call_host_function("np.floor", [x]), np.floor(x))
    assert F.call_host_function("to_int", [x]).tolist() == [1, -2, 9]              # np.array(..., dtype='int')
    assert F.call_host_function("to_int", ["1", "2", "3"]).tolist() == [1, 2, 3]   # several arguments -> one array
    assert F.call_host_function("to_str", [np.array([1, 2])]).tolist() == ["1", "2"]
    assert F.call_host_function("upper", [pa.array(["ab", None])]).to_pylist() == ["AB", None]
    assert F.call_host_function("concat", [pa.array(["a", "b"]), "-", np.array([1, 2])]).to_pylist() == ["a-1", "b-2"]
    assert F.call_host_function("pi", []) == np.pi
    assert F.call_host_function("datetime", ["2020-10"]).dtype == np.dtype("datetime64[D]")        # month -> day
    assert F.call_host_function("datetime", ["2020-10-07 19"]).dtype == np.dtype("datetime64[s]")  # hour -> second
    assert F.call_host_function("from_timestamp", [np.array([0.0, np.nan])]).astype("int64").tolist()[0] == 0
    with pytest.raises(F.FunctionError):
        F.call_host_function("date", ["2020-10-07", "s"])                          # DateFunction.UNITS = ['D']
    with pytest.raises(F.FunctionError):
        F.call_host_function("no_such_function", [x])
    F.register_numpy("twice", lambda a: a * 2)
    assert F.call_host_function("TWICE", [x]).tolist() == [3.0, -4.5, 18.0]
