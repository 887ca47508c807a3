"""Fused expression chains (-m gpu), SURVEY 8a row a3: every chain must be bit-identical to the
node-by-node NumPy evaluation the reference performs (VectorizedExpression.evaluate,
vinum/core/base.py:105-125, with the ufuncs of vinum/core/expressions.py:13-36), as a projection, as a
mask, and as a predicate fused into the filter and aggregate kernels."""
import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

NP = {"+": np.add, "-": np.subtract, "*": np.multiply, "/": np.divide, "%": np.mod, "&": np.bitwise_and,
      "|": np.bitwise_or, "#": np.bitwise_xor}
CMP = {"==": np.equal, "!=": np.not_equal, ">": np.greater, ">=": np.greater_equal, "<": np.less, "<=": np.less_equal}


@pytest.fixture(scope="module")
def vb(stream):
    import vinum_b200
    return vinum_b200


@pytest.fixture(scope="module")
def cols(vb, stream):
    rng = np.random.default_rng(12)
    n = 200_003
    host = {
        "a": rng.integers(-1000, 1000, n),
        "b": rng.integers(-2**62, 2**62, n),         # products wrap
        "z": rng.integers(-3, 4, n),                 # zeros: x % 0, x / 0
        "x": rng.normal(0, 100, n),
        "y": np.where(rng.random(n) < 0.05, 0.0, rng.normal(0, 1, n)),
    }
    host["y"][::97] = np.nan
    host["x"][::101] = np.inf
    dev = {k: vb.DeviceColumn.from_numpy(v, stream) for k, v in host.items()}
    return host, dev


def _numpy_chain(chain, host):
    with np.errstate(all="ignore"):
        acc = host[chain[0][1]] if isinstance(chain[0][1], str) else chain[0][1]
        for op, t in chain[1:]:
            acc = NP[op](acc, host[t] if isinstance(t, str) else t)
    return acc


def _dev_chain(chain, dev):
    return [(op, dev[t] if isinstance(t, str) else t) for op, t in chain]


CHAINS = [
    [(None, "a"), ("*", 10), ("+", "z")],
    [(None, "b"), ("*", "a"), ("-", 7), ("#", "z")],
    [(None, "a"), ("%", "z"), ("+", 1)],
    [(None, "a"), ("/", "z"), ("*", 2)],                 # int / int -> float64, then float
    [(None, "x"), ("*", 2.5), ("-", "y"), ("/", "y")],
    [(None, "x"), ("%", "y"), ("+", "a")],               # float mod (sign of the divisor, NaN for 0), int promoted
    [(None, "a"), ("+", 0.5), ("*", "b")],               # int column meets a float literal
    [(None, 3), ("-", "a"), ("&", 255)],
    [(None, "b"), ("+", "b"), ("+", "b"), ("+", "b")],   # wraps
]


@pytest.mark.parametrize("chain", CHAINS, ids=lambda c: " ".join(str(t) if o is None else f"{o} {t}" for o, t in c))
def test_chain_projection_is_bit_identical_to_numpy(vb, stream, cols, chain):
    from vinum_b200 import ops
    host, dev = cols
    want = _numpy_chain(chain, host)
    got = ops.eval_chain(_dev_chain(chain, dev), stream).to_numpy(stream)
    assert got.dtype == want.dtype
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("op", sorted(CMP))
def test_chain_compare_mask_and_fused_predicates(vb, stream, cols, op):
    from vinum_b200 import ops, _lib as L
    host, dev = cols
    cases = [([(None, "a"), ("*", 10)], [(None, "z"), ("*", 1000)]),
             ([(None, "x"), ("*", 2.0), ("+", "y")], [(None, "a")]),
             ([(None, "b")], [(None, "a"), ("*", "b")]),
             ([(None, "a"), ("/", "z")], [(None, 0.5)])]
    for lhs, rhs in cases:
        with np.errstate(all="ignore"):
            want = CMP[op](_numpy_chain(lhs, host), _numpy_chain(rhs, host))
        dl, dr = _dev_chain(lhs, dev), _dev_chain(rhs, dev)
        mask = ops.compare_chains(dl, op, dr, stream).to_numpy(stream).astype(bool)
        assert np.array_equal(mask, want), (lhs, op, rhs)
        # the same comparison evaluated INSIDE the compaction kernel ...
        batch = vb.DeviceBatch([dev["a"], dev["x"]], ["a", "x"])
        out = ops.filter_batch(batch, ops.Predicate.expr(dl, op, dr), stream)
        assert out.num_rows == int(want.sum())
        assert np.array_equal(out.column("a").to_numpy(stream), host["a"][want])
        assert np.array_equal(out.column("x").to_numpy(stream).view(np.uint64), host["x"][want].view(np.uint64))
        # ... and inside the fused filter -> hash aggregate kernel (two chunks: the chain's columns are sliced)
        agg = vb.Aggregator([pa.int64()], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.int64())])
        n = len(want)
        half = (n // 2) & ~1
        for lo, hi in ((0, half), (half, n)):
            sl = lambda ch: [(o, t.slice(lo, hi - lo) if isinstance(t, vb.DeviceColumn) else t) for o, t in ch]
            agg.update([dev["z"].slice(lo, hi - lo)], [None, dev["a"].slice(lo, hi - lo)],
                       ops.Predicate.expr(sl(dl), op, sl(dr)), stream)
        keys, aggs = agg.result_arrays(stream)
        order = np.argsort(keys[0].to_numpy())
        uk = np.unique(host["z"][want])
        assert np.array_equal(keys[0].to_numpy()[order], uk)
        assert np.array_equal(aggs[0].to_numpy()[order], np.array([(host["z"][want] == k).sum() for k in uk], dtype=np.uint64))
        assert np.array_equal(aggs[1].to_numpy()[order], np.array([host["a"][want][host["z"][want] == k].sum() for k in uk]))


def test_sql_expressions_fuse_and_match_the_unfused_engine(vb, monkeypatch):
    rng = np.random.default_rng(3)
    n = 150_000
    t = pa.table({"a": rng.integers(-1000, 1000, n), "b": rng.integers(-10**6, 10**6, n), "x": rng.normal(0, 10, n),
                  "k": rng.integers(0, 50, n), "s": pa.array(rng.integers(0, 9, n).astype(str))})
    queries = [
        "SELECT a * 10 + b AS v, x / 2 - a AS w FROM t WHERE a * 10 > b",
        "SELECT k, COUNT(*) AS c, SUM(x) AS sx FROM t WHERE a * 10 > b - 5 GROUP BY k ORDER BY k",
        "SELECT a, s FROM t WHERE x * 2 + a > b / 1000 ORDER BY a, s LIMIT 50",
        "SELECT k, SUM(a * 2 + b) AS sv FROM t WHERE b + a * 3 <= x GROUP BY k ORDER BY k",
    ]
    tbl = vb.Table.from_arrow(t)
    for q in queries:
        monkeypatch.setenv("VINUM_B200_FUSE_EXPR", "1")
        fused = tbl.sql(q).to_arrow()
        assert tbl.last_stats.get("fused_expr", 0) >= 1, q
        monkeypatch.setenv("VINUM_B200_FUSE_EXPR", "0")
        plain = tbl.sql(q).to_arrow()
        assert tbl.last_stats.get("fused_expr", 0) == 0
        assert fused.schema == plain.schema, q
        for name in fused.column_names:
            f, p = fused.column(name).to_numpy(zero_copy_only=False), plain.column(name).to_numpy(zero_copy_only=False)
            if f.dtype.kind == "f" and name in ("sx",):
                assert np.allclose(f, p, rtol=1e-9, atol=0), (q, name)
            else:
                assert np.array_equal(f, p), (q, name)
