"""GPU parity tests (-m gpu): every CUDA operator, called through the C ABI, against
  * the reference-generated golden fixtures (tests/golden/*.arrow),
  * the reference's own gtest vectors (tests/golden/gtest_vectors.py),
  * the CPU oracle (oracle/vinum_oracle.py) on seeded random inputs.
Bar: bit-exact for masks, filters, integer/temporal aggregates, keys, permutations;
1e-6 relative for floating SUM/AVG (BASELINE.json north_star).
"""
import os

import numpy as np
import pyarrow as pa
import pytest

import gtest_vectors as G
from golden_util import assert_tables_match, manifest, read_arrow
from oracle import vinum_oracle as O

pytestmark = pytest.mark.gpu

MAN = manifest()
FLOAT_RTOL = 1e-6  # north_star tolerance for floating-point sums / averages


@pytest.fixture(scope="module")
def vb(stream):
    import vinum_b200
    return vinum_b200


def _dev(vb, arr, stream):
    return vb.DeviceColumn.from_arrow(arr, stream)


# ------------------------------------------------------------------ datagen ----
def test_datagen_bit_identical_host_device(vb, stream):
    from vinum_b200 import datagen
    for row0, n in ((0, 100_003), (7_777_777_777, 65_537)):
        for name in datagen.KINDS:
            dev = datagen.device_column(name, row0, n, stream=stream).to_numpy(stream)
            host = datagen.host_column(name, row0, n)
            assert dev.dtype == host.dtype
            assert np.array_equal(dev.view(np.uint8), host.view(np.uint8)), name


# -------------------------------------------------------- expressions (a1, a5) ----
@pytest.mark.parametrize("tname", ["t_nulls", "t_dense"])
def test_expressions_match_reference_fixture(vb, stream, tname):
    from vinum_b200 import ops
    table = read_arrow(f"{tname}.in.arrow").combine_chunks()
    outs = read_arrow(f"{tname}.expr.out.arrow")
    cols = {}

    def col(name):
        if name not in cols:
            cols[name] = _dev(vb, table.column(name).chunk(0), stream)
        return cols[name]

    checked = 0
    for spec in MAN["expr"]:
        if spec["table"] != tname:
            continue
        want = np.asarray(outs.column(spec["key"]).to_numpy(zero_copy_only=False))
        kind = spec["kind"]
        if kind == "cmp":
            got = ops.compare(col(spec["col"]), spec["op"], spec["scalar"], stream)
        elif kind == "cmpcol":
            got = ops.compare(col(spec["a"]), spec["op"], col(spec["b"]), stream)
        elif kind == "between":
            got = ops.between(col(spec["col"]), spec["lo"], spec["hi"], spec["negate"], stream)
        elif kind == "isin":
            got = ops.isin(col(spec["col"]), spec["values"], spec["negate"], stream)
        else:
            a = col(spec["a"]) if spec["a_is_col"] else spec["a"]
            b = col(spec["b"]) if spec["b_is_col"] else spec["b"]
            got = ops.arith(spec["op"], a, b, stream)
        g = got.to_numpy(stream)
        assert str(g.dtype) == spec["dtype"], (spec["key"], g.dtype, spec["dtype"])
        w = want.astype(g.dtype)
        if g.dtype.kind == "f":
            # bit-exact (single IEEE ops, -fmad=false); NaN payloads may differ -> compare NaN-ness
            nan = np.isnan(w)
            assert np.array_equal(np.isnan(g), nan), spec["key"]
            assert np.array_equal(g[~nan].view(np.uint8), w[~nan].view(np.uint8)), spec["key"]
        else:
            assert np.array_equal(g, w), spec["key"]
        checked += 1
    assert checked > 100


def test_mask_algebra_and_null_tests(vb, stream):
    from vinum_b200 import ops
    table = read_arrow("t_nulls.in.arrow").combine_chunks()
    x = table.column("v_f64_clean").chunk(0)
    y = table.column("v_i32").chunk(0)
    dx, dy = _dev(vb, x, stream), _dev(vb, y, stream)
    a = ops.compare(dx, ">", 0.0, stream)
    b = ops.compare(dy, "<", 5, stream)
    ha, hb = O.compare(x, ">", 0.0), O.compare(y, "<", 5)
    assert np.array_equal(ops.mask_and(a, b, stream).to_numpy(stream), O.mask_and(ha, hb))
    assert np.array_equal(ops.mask_or(a, b, stream).to_numpy(stream), O.mask_or(ha, hb))
    assert np.array_equal(ops.mask_not(a, stream).to_numpy(stream), O.mask_not(ha))
    assert np.array_equal(ops.is_null(dx, stream).to_numpy(stream), O.is_null(x))
    assert np.array_equal(ops.is_valid(dy, stream).to_numpy(stream), O.is_valid(y))
    # sliced arrays (non-zero Arrow offset, validity bit offset not byte aligned)
    xs = x.slice(5, 100)
    assert np.array_equal(ops.is_null(_dev(vb, xs, stream), stream).to_numpy(stream), O.is_null(xs))
    assert np.array_equal(ops.compare(_dev(vb, xs, stream), "<=", 1.0, stream).to_numpy(stream), O.compare(xs, "<=", 1.0))


# ------------------------------------------------------------------ filter (a4) ----
def _filter_check(vb, stream, table, pred_col, op, scalar, use_mask):
    from vinum_b200 import ops
    batch = table.combine_chunks().to_batches()[0] if table.num_rows else pa.RecordBatch.from_arrays(
        [pa.array([], type=f.type) for f in table.schema], schema=table.schema)
    dev = vb.DeviceBatch.from_arrow(batch, stream)
    mask = O.compare(batch.column(batch.schema.get_field_index(pred_col)), op, scalar)
    want = O.filter_batch(batch, mask)
    if use_mask:
        dmask = ops.compare(dev.column(pred_col), op, scalar, stream)
        pred = ops.Predicate.from_mask(dmask)
    else:
        pred = ops.Predicate.compare(dev.column(pred_col), op, scalar)
    got = ops.filter_batch(dev, pred, stream).to_arrow(stream)
    assert_tables_match(got, want, rtol=0.0, float_exact_cols=want.schema.names)


@pytest.mark.parametrize("tname", ["t_nulls", "t_dense"])
@pytest.mark.parametrize("use_mask", [False, True])
def test_filter_every_dtype_with_nulls(vb, stream, tname, use_mask):
    table = read_arrow(f"{tname}.in.arrow")
    for pred_col, op, scalar in (("v_f64_clean", ">", 0.0), ("v_i64", "<=", 1000), ("v_f64", "!=", 3.0),
                                 ("k_i8", "==", 1), ("v_u8", ">=", 128), ("v_f32", "<", 0.25)):
        _filter_check(vb, stream, table, pred_col, op, scalar, use_mask)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 2047, 2048, 2049, 4096, 100_001])
def test_filter_ragged_sizes_preserve_order(vb, stream, n):
    from vinum_b200 import datagen
    table = datagen.host_table(["i1", "i2", "f0", "f1", "k32"], 0, n)
    _filter_check(vb, stream, table, "f0", ">", 0.5, False)
    _filter_check(vb, stream, table, "f0", ">", 0.5, True)
    _filter_check(vb, stream, table, "i1", "<", 0, False)
    if n:
        _filter_check(vb, stream, table, "f0", ">", 2.0, False)   # selects nothing
        _filter_check(vb, stream, table, "f0", ">=", 0.0, False)  # selects everything


def test_filter_many_columns(vb, stream):
    # more columns than one kernel pass carries (FT_MAX_COLS = 12)
    rng = np.random.default_rng(5)
    n = 10_000
    table = pa.table({f"c{i}": rng.integers(0, 100, n) for i in range(15)})
    _filter_check(vb, stream, table, "c3", ">", 50, False)


# ---------------------------------------------------------- aggregate (a8-a14) ----
def _lib_funcs(vl, funcs):
    return [vl.AggFuncDef(getattr(vl.AggFuncType, t), c, o) for t, c, o in funcs]


def _run_vinum_lib(vb, batches, gb, ac, funcs):
    vl = vb.vinum_lib
    fd = _lib_funcs(vl, funcs)
    if len(gb) == 0:
        agg = vl.OneGroupAggregate(fd)
    elif len(gb) == 1:
        agg = vl.SingleNumericalHashAggregate(gb, ac, fd)
    else:
        agg = vl.MultiNumericalHashAggregate(gb, ac, fd)
    for b in batches:
        agg.next(b)
    return agg.result()


@pytest.mark.parametrize("case", G.cases(), ids=lambda c: c[0])
def test_aggregate_matches_gtest_vectors(vb, case):
    """The reference's own expected batches, through the drop-in vinum_lib classes,
    table fed in two halves like the gtest (hash_agg_test.cpp:108-133)."""
    name, table, gb, ac, funcs, expected, sort_cols = case
    got = G.sort_result(_run_vinum_lib(vb, G.split_in_two(table), gb, ac, funcs), sort_cols)
    assert_tables_match(got, expected, rtol=FLOAT_RTOL)


@pytest.mark.parametrize("case", G.generic_cases(), ids=lambda c: c[0])
def test_generic_aggregate_matches_gtest_vectors(vb, case):
    """GenericHashAggregate (string / bool keys, string MIN/MAX): the reference's expected
    batches (hash_agg_test.cpp:286-338, :439-477, :601-650), table fed in two halves."""
    name, table, gb, ac, funcs, expected, sort_cols = case
    agg = vb.vinum_lib.GenericHashAggregate(gb, ac, _lib_funcs(vb.vinum_lib, funcs))
    for b in G.split_in_two(table):
        agg.next(b)
    got = G.sort_result(agg.result(), sort_cols)
    assert_tables_match(got, expected, rtol=FLOAT_RTOL)


def test_generic_aggregate_random_vs_compiled_reference(vb):
    """String + numeric keys with NULLs, many batches, against the reference's own
    GenericHashAggregate (oracle/_ref)."""
    from oracle import ref
    lib = ref.ref_lib()
    if lib is None or not hasattr(lib, "GenericHashAggregate"):
        pytest.skip("oracle/_ref with GenericHashAggregate is not built")
    rng = np.random.default_rng(7)
    n = 50_000
    words = np.array(["alpha", "beta", "gamma", "delta", "", "epsilon", "zeta", "a much longer key value"], dtype=object)
    table = pa.table({
        "s": pa.array(words[rng.integers(0, len(words), n)], type=pa.string(), mask=rng.random(n) < 0.05),
        "k": pa.array(rng.integers(0, 5, n).astype(np.int16), mask=rng.random(n) < 0.05),
        "b": pa.array(rng.random(n) < 0.5, mask=rng.random(n) < 0.1),
        "v": pa.array(rng.normal(0, 10, n), mask=rng.random(n) < 0.1),
        "w": pa.array(rng.integers(-1000, 1000, n).astype(np.int64)),
        "t": pa.array(words[rng.integers(0, len(words), n)], type=pa.string(), mask=rng.random(n) < 0.2),
    })
    funcs = [("COUNT_STAR", "", "c"), ("COUNT", "t", "ct"), ("SUM", "v", "sv"), ("AVG", "w", "aw"), ("MIN", "v", "mv"),
             ("MAX", "w", "xw"), ("MIN", "t", "mt"), ("MAX", "t", "xt")]
    for gb in (["s"], ["s", "k"], ["b", "s", "k"]):
        want_agg = lib.GenericHashAggregate(gb, gb, [lib.AggFuncDef(getattr(lib.AggFuncType, t), c, o) for t, c, o in funcs])
        got_agg = vb.vinum_lib.GenericHashAggregate(gb, gb, _lib_funcs(vb.vinum_lib, funcs))
        for b in table.to_batches(max_chunksize=7000):
            want_agg.next(b)
            got_agg.next(b)
        assert_tables_match(got_agg.result(), want_agg.result(), key_cols=gb, rtol=FLOAT_RTOL,
                            float_exact_cols=["mv"])


def test_string_min_max_with_numeric_keys_vs_compiled_reference(vb):
    """MIN / MAX / COUNT over a string column under every aggregate class (StringMinMaxFunc,
    agg_funcs.h:219-261): rank codes reduced on the device batch by batch, winners merged on the host.
    NULL strings, a group whose strings are all NULL, NaN keys, many batches."""
    from oracle import ref
    lib = ref.ref_lib()
    if lib is None:
        pytest.skip("oracle/_ref is not built")
    rng = np.random.default_rng(11)
    n = 60_000
    words = np.array(["pear", "apple", "fig", "", "Apple", "zucchini", "fig tree", "\u00e9clair", "apple pie"], dtype=object)
    k = rng.integers(0, 40, n).astype(np.int64)
    strings = words[rng.integers(0, len(words), n)]
    null_s = rng.random(n) < 0.15
    null_s[k == 7] = True                      # a group whose strings are all NULL
    f = rng.integers(0, 6, n).astype(np.float64)
    f[rng.random(n) < 0.1] = np.nan
    table = pa.table({
        "k": pa.array(k),          # (the compiled reference's numerical classes crash on NULL keys: none here)
        "j": pa.array(rng.integers(0, 3, n).astype(np.int32)),
        "f": pa.array(f),
        "t": pa.array(strings, type=pa.string(), mask=null_s),
        "v": pa.array(rng.normal(0, 10, n)),
    })
    funcs = [("COUNT_STAR", "", "c"), ("MIN", "t", "mt"), ("MAX", "t", "xt"), ("COUNT", "t", "ct"), ("SUM", "v", "sv")]
    for cls, gb in (("SingleNumericalHashAggregate", ["k"]), ("SingleNumericalHashAggregate", ["f"]),
                    ("MultiNumericalHashAggregate", ["k", "j"]), ("OneGroupAggregate", [])):
        mk = lambda m: (getattr(m, cls)([m.AggFuncDef(getattr(m.AggFuncType, a), c, o) for a, c, o in funcs]) if not gb else
                        getattr(m, cls)(gb, gb, [m.AggFuncDef(getattr(m.AggFuncType, a), c, o) for a, c, o in funcs]))
        got_agg = mk(vb.vinum_lib)
        want_agg = mk(lib) if gb else None     # the compiled reference's OneGroupAggregate crashes on a string MIN
        for b in table.to_batches(max_chunksize=9000):
            if want_agg is not None:
                want_agg.next(b)
            got_agg.next(b)
        got = got_agg.result()
        if gb:
            assert_tables_match(got, want_agg.result(), key_cols=gb, rtol=FLOAT_RTOL)
        else:
            import pyarrow.compute as pc
            assert got.column("mt").to_pylist() == [pc.min(table.column("t")).as_py()]
            assert got.column("xt").to_pylist() == [pc.max(table.column("t")).as_py()]
            assert got.column("ct").to_pylist() == [len(table) - table.column("t").null_count]


def test_sql_string_min_max(vb):
    """The same through Table.sql, with a WHERE (the fused predicate applies to the rank-code aggregate too)
    and a string GROUP BY key."""
    import pyarrow.compute as pc
    rng = np.random.default_rng(12)
    n = 200_003
    words = np.array([f"w{i:04d}" for i in range(3000)], dtype=object)
    table = pa.table({
        "k": rng.integers(0, 500, n).astype(np.int64),
        "g": pa.array(np.array(["x", "y", "z"], dtype=object)[rng.integers(0, 3, n)], type=pa.string()),
        "t": pa.array(words[rng.integers(0, len(words), n)], type=pa.string(), mask=rng.random(n) < 0.1),
        "p": rng.random(n),
    })
    tbl = vb.Table.from_arrow(table)
    kept = table.filter(pc.greater(table.column("p"), 0.3))
    got = tbl.sql("SELECT k, MIN(t) AS mn, MAX(t) AS mx, COUNT(t) AS c FROM t WHERE p > 0.3 GROUP BY k ORDER BY k").to_arrow()
    want = kept.group_by("k", use_threads=False).aggregate([("t", "min"), ("t", "max"), ("t", "count")]).sort_by("k")
    assert got.column("k").to_pylist() == want.column("k").to_pylist()
    assert got.column("mn").to_pylist() == want.column("t_min").to_pylist()
    assert got.column("mx").to_pylist() == want.column("t_max").to_pylist()
    assert got.column("c").to_pylist() == want.column("t_count").to_pylist()
    got = tbl.sql("SELECT g, MAX(t) AS mx FROM t GROUP BY g ORDER BY g").to_arrow()
    want = table.group_by("g", use_threads=False).aggregate([("t", "max")]).sort_by("g")
    assert got.column("mx").to_pylist() == want.column("t_max").to_pylist()
    got = tbl.sql("SELECT MIN(t) AS mn, MAX(t) AS mx FROM t WHERE p < 0.001").to_arrow()
    few = table.filter(pc.less(table.column("p"), 0.001))
    assert got.column("mn").to_pylist() == [pc.min(few.column("t")).as_py()]
    assert got.column("mx").to_pylist() == [pc.max(few.column("t")).as_py()]


@pytest.mark.parametrize("case", MAN["agg"], ids=lambda c: f"{c['table']}.{c['name']}")
def test_aggregate_matches_reference_fixture(vb, case):
    table = read_arrow(f"{case['table']}.in.arrow")
    want = read_arrow(f"{case['table']}.{case['name']}.out.arrow")
    funcs = [tuple(f) for f in case["funcs"]]
    got = _run_vinum_lib(vb, table.to_batches(max_chunksize=64), case["groupby"], case["agg_cols"], funcs)
    exact = [f[2] for f in funcs if f[0] in ("MIN", "MAX")]  # MIN/MAX are selections: bit-exact
    assert_tables_match(got, want, key_cols=case["agg_cols"], rtol=FLOAT_RTOL, float_exact_cols=exact)


# Every kernel path of the fused aggregate, forced through per-object options (vk_set_option is sampled by
# vk_agg_create): direct group ids vs the CTA hash table vs the global-table kernel, 8 vs 12 warps, L2
# prefetch off / on.  `path` is what vk_agg_last_path must report.
AGG_PATHS = {
    "direct_auto": (dict(), 1),
    "direct_w8_pf6": (dict(AGG_WARPS=8, AGG_PF=6), 1),
    "direct_w12_pf0": (dict(AGG_WARPS=12, AGG_PF=0), 1),
    "hash_w8": (dict(AGG_DIRECT=0, AGG_WARPS=8), 1),
    "hash_insert_w8": (dict(AGG_DIRECT=0, AGG_DICT=0, AGG_WARPS=8), 1),
    "hash_w12_small_table": (dict(AGG_DIRECT=0, AGG_WARPS=12, AGG_LOG2S=11), 1),
    "hash_w4": (dict(AGG_DIRECT=0, AGG_WARPS=4, AGG_PF=0), 1),
    "direct_split_entries": (dict(AGG_ENTRY=2), 1),
    "direct_tag_arbitration": (dict(AGG_ENTRY=0), 1),
    "hash_split_entries": (dict(AGG_DIRECT=0, AGG_ENTRY=2), 1),
    "general_kernel": (dict(AGG_NOFAST=1), 2),
    "general_kernel_one_row": (dict(AGG_NOFAST=1, AGG_WIDE=0), 2),
    "partitioned": (dict(AGG_NOFAST=1, AGG_PARTITION=2), 4),
}


@pytest.mark.parametrize("path", sorted(AGG_PATHS))
@pytest.mark.parametrize("keyname", ["i0", "k32"])
def test_northstar_filter_aggregate_vs_oracle(vb, stream, path, keyname):
    """SELECT k, COUNT(*), SUM(f1), AVG(f1) FROM t WHERE f0 > 0.5 GROUP BY k -- every kernel path of the
    fused aggregate against the oracle's restatement of the reference chain
    (single_numerical_hash_aggregate.cpp:15-46, agg_funcs.h:97-127,280-317,439-542)."""
    from vinum_b200 import datagen, ops, _lib as L
    opts, want_path = AGG_PATHS[path]
    n = 1_000_003
    table = datagen.host_table([keyname, "f0", "f1"], 0, n)
    funcs = [("COUNT_STAR", "", "count_star"), ("SUM", "f1", "sum_f1"), ("AVG", "f1", "avg_f1")]
    want = O.filter_hash_aggregate(table, "f0", ">", 0.5, [keyname], funcs)
    dev = datagen.device_table([keyname, "f0", "f1"], 0, n, stream=stream)
    with vb.options(AGG_LEARN_LOG2=16, **opts):   # a short learning launch: the main launch sees most rows
        agg = vb.Aggregator([table.schema.field(keyname).type], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64()),
                                                                  (L.AGG_AVG, pa.float64())])
    half = 500_736
    for lo, hi in ((0, half), (half, n)):
        part = dev.slice(lo, hi - lo)
        agg.update([part.column(keyname)], [None, part.column("f1"), part.column("f1")],
                   ops.Predicate.compare(part.column("f0"), ">", 0.5), stream)
    assert agg.last_path == want_path
    keys, aggs = agg.result_arrays(stream)
    got = pa.table([keys[0]] + aggs, names=[keyname, "count_star", "sum_f1", "avg_f1"])
    assert_tables_match(got, want, key_cols=[keyname], rtol=FLOAT_RTOL)


def test_aggregate_random_types_vs_oracle(vb):
    rng = np.random.default_rng(11)
    n = 200_000
    table = pa.table({
        "k": pa.array(rng.integers(-500, 500, n).astype(np.int32), mask=rng.random(n) < 0.01),
        "k2": pa.array(rng.integers(0, 4, n).astype(np.int16)),
        "f": pa.array(rng.normal(0, 1e3, n), mask=rng.random(n) < 0.1),
        "f32": pa.array(rng.normal(0, 10, n).astype(np.float32)),
        "i": pa.array(rng.integers(-2**62, 2**62, n), mask=rng.random(n) < 0.1),
        "u": pa.array(rng.integers(0, 2**64 - 1, n, dtype=np.uint64)),
        "s": pa.array(rng.integers(-100, 100, n).astype(np.int8), mask=rng.random(n) < 0.5),
        "ts": pa.array(rng.integers(0, 10**12, n), type=pa.timestamp("us")),
    })
    funcs = [("COUNT_STAR", "", "c"), ("COUNT", "f", "cf"), ("SUM", "f", "sf"), ("AVG", "f", "af"), ("MIN", "f", "mnf"),
             ("MAX", "f", "mxf"), ("SUM", "f32", "sf32"), ("MIN", "f32", "mnf32"), ("AVG", "f32", "af32"),
             ("MAX", "i", "mxi"), ("MIN", "i", "mni"), ("SUM", "i", "si"), ("AVG", "i", "ai"),
             ("SUM", "u", "su"), ("AVG", "u", "au"), ("MAX", "u", "mxu")]
    funcs2 = [("SUM", "s", "ss"), ("AVG", "s", "as"), ("MIN", "s", "mns"), ("MIN", "ts", "mnts"), ("MAX", "ts", "mxts"),
              ("COUNT", "s", "cs")]
    batches = table.to_batches(max_chunksize=33_333)
    for gb in (["k"], ["k", "k2"], ["k2"], []):
        for fs in (funcs, funcs2):
            want = O.hash_aggregate(batches, gb, gb, fs)
            got = _run_vinum_lib(vb, batches, gb, gb, fs)
            exact = [f[2] for f in fs if f[0] in ("MIN", "MAX")]
            assert_tables_match(got, want, key_cols=gb, rtol=FLOAT_RTOL, float_exact_cols=exact)


def test_aggregate_high_cardinality_growth_and_replay(vb, stream):
    """More groups than the initial table holds: rows deferred to the replay list, table
    rehashed, nothing lost."""
    from vinum_b200 import _lib as L
    rng = np.random.default_rng(3)
    n = 3_000_000
    keys = rng.integers(0, 2_500_000, n)
    vals = rng.normal(0, 1, n)
    agg = vb.Aggregator([pa.int64()], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())])
    dk, dv = vb.DeviceColumn.from_numpy(keys, stream), vb.DeviceColumn.from_numpy(vals, stream)
    agg.update([dk], [None, dv], None, stream)
    k, kv, cnt, lo, hi, valid = agg.result_raw(stream)
    uk, inv = np.unique(keys, return_inverse=True)
    order = np.argsort(k[0].view(np.int64))
    assert np.array_equal(k[0].view(np.int64)[order], uk)
    assert np.array_equal(cnt[order], np.bincount(inv).astype(np.uint64))
    assert np.allclose(lo[1].view(np.float64)[order], np.bincount(inv, weights=vals), rtol=FLOAT_RTOL, atol=1e-9)
    assert int(cnt.sum()) == n


def test_aggregate_key_semantics(vb):
    """A.3: -0.0 and +0.0 are distinct groups, equal NaN bit patterns group, NULL is its own group."""
    t = pa.table({"k": pa.array([0.0, -0.0, float("nan"), float("nan"), None, None, 1.0]),
                  "v": pa.array([1, 2, 3, 4, 5, 6, 7], type=pa.int64())})
    funcs = [("COUNT_STAR", "", "c"), ("SUM", "v", "s")]
    got = _run_vinum_lib(vb, t.to_batches(), ["k"], ["k"], funcs)
    want = O.hash_aggregate_rowwise(t.to_batches(), ["k"], ["k"], funcs)
    assert got.num_rows == 5
    assert_tables_match(got, want, key_cols=["k"], rtol=0.0)
    # NULL group is emitted last (single_numerical_hash_aggregate.cpp:54-60)
    assert got.column(0).to_pylist()[-1] is None


def test_aggregate_all_null_group_and_empty_inputs(vb):
    t = pa.table({"k": pa.array([1, 1, 2], type=pa.int32()), "v": pa.array([None, None, 2.5])})
    funcs = [("COUNT_STAR", "", "c"), ("COUNT", "v", "cv"), ("SUM", "v", "s"), ("MAX", "v", "m"), ("AVG", "v", "a")]
    got = _run_vinum_lib(vb, t.to_batches(), ["k"], ["k"], funcs)
    want = O.hash_aggregate_rowwise(t.to_batches(), ["k"], ["k"], funcs)
    assert_tables_match(got, want, key_cols=["k"], rtol=0.0)
    empty = t.slice(0, 0).to_batches() or [pa.RecordBatch.from_arrays(
        [pa.array([], type=f.type) for f in t.schema], schema=t.schema)]
    got = _run_vinum_lib(vb, empty, ["k"], ["k"], funcs)
    assert got.num_rows == 0 and got.schema.names == ["k", "c", "cv", "s", "m", "a"]
    assert got.schema.field("s").type == pa.float64() and got.schema.field("c").type == pa.uint64()


def test_aggregate_int64_min_sum_is_decimal(vb):
    """By the code (huge_int.cpp:341-361) a sum of exactly -2^63 fails the int64 cast."""
    t = pa.table({"k": pa.array([1, 1, 2], type=pa.int64()),
                  "v": pa.array([-(2**62), -(2**62), 5], type=pa.int64())})
    funcs = [("SUM", "v", "s")]
    got = _run_vinum_lib(vb, t.to_batches(), ["k"], ["k"], funcs)
    want = O.hash_aggregate_rowwise(t.to_batches(), ["k"], ["k"], funcs)
    assert want.schema.field("s").type == pa.decimal128(38, 0)
    assert_tables_match(got, want, key_cols=["k"], rtol=0.0)


def test_vinum_lib_error_behaviour(vb):
    vl = vb.vinum_lib
    t = pa.table({"k": pa.array([1, 2], type=pa.int32()), "v": pa.array([1.0, 2.0]), "b": pa.array([True, False]),
                  "d": pa.array([1, 2], type=pa.date32())})
    b = t.to_batches()[0]
    agg = vl.SingleNumericalHashAggregate(["nope"], ["nope"], [vl.AggFuncDef(vl.AggFuncType.COUNT_STAR, "", "c")])
    with pytest.raises(RuntimeError, match="Column not found: nope"):
        agg.next(b)
    agg = vl.SingleNumericalHashAggregate(["k"], ["k"], [vl.AggFuncDef(vl.AggFuncType.SUM, "b", "s")])
    with pytest.raises(RuntimeError, match=r"not supported by sum\(\)"):
        agg.next(b)
    agg = vl.SingleNumericalHashAggregate(["k"], ["k"], [vl.AggFuncDef(vl.AggFuncType.AVG, "d", "s")])
    with pytest.raises(RuntimeError, match=r"not supported by avg\(\)"):
        agg.next(b)
    with pytest.raises(TypeError):
        vl.SingleNumericalHashAggregate(["k"], ["k"], []).next("not a batch")
    assert vl.import_pyarrow() == 0
    assert repr(vl.AggFuncDef(vl.AggFuncType.SUM, "v", "o")) == "<AggFuncDef col_name: v, out_col_name: o>"
    rdr = vl.TableBatchReader(pa.table({"x": list(range(25))}))
    rdr.set_batch_size(10)
    sizes = []
    while True:
        rb = rdr.next()
        if rb is None:
            break
        sizes.append(rb.num_rows)
    assert sizes == [10, 10, 5]


# -------------------------------------------------------------------- sort (a16) ----
def _sorted_rows(vb, table, cols, orders):
    vl = vb.vinum_lib
    s = vl.Sort(cols, [getattr(vl.SortOrder, o) for o in orders])
    for b in table.to_batches(max_chunksize=64):
        s.next(b)
    return s.sorted()


@pytest.mark.parametrize("case", MAN["sort"], ids=lambda c: f"{c['table']}.{c['name']}")
def test_sort_matches_reference_fixture(vb, case):
    table = read_arrow(f"{case['table']}.in.arrow")
    want_rows = read_arrow(f"{case['table']}.{case['name']}.out.arrow").column("row").to_numpy()
    got = _sorted_rows(vb, table, case["cols"], case["orders"])
    assert np.array_equal(got.column(got.schema.get_field_index("row")).to_numpy(), want_rows)
    # Take of every column: the sorted batch equals the table gathered by the reference permutation
    want = table.combine_chunks().take(pa.array(want_rows))
    assert_tables_match(got, want, rtol=0.0, float_exact_cols=want.schema.names)


def test_sort_special_values_probe(vb):
    """SURVEY A.5 probe: stable, -0.0 == +0.0, NaN then NULL last in both directions."""
    t = pa.table({"x": pa.array([2, 1, 2, None, float("nan"), 1, -0.0, 0.0, float("inf"), float("-inf")]),
                  "row": pa.array(range(10))})
    asc = _sorted_rows(vb, t, ["x"], ["ASC"]).column(1).to_pylist()
    desc = _sorted_rows(vb, t, ["x"], ["DESC"]).column(1).to_pylist()
    assert asc == [9, 6, 7, 1, 5, 0, 2, 8, 4, 3]
    assert desc == [8, 0, 2, 1, 5, 6, 7, 9, 4, 3]


@pytest.mark.parametrize("n", [1, 2, 4095, 4096, 4097, 300_001])
def test_sort_random_vs_oracle(vb, stream, n):
    from vinum_b200 import ops, _lib as L
    rng = np.random.default_rng(n)
    table = pa.table({
        "f": pa.array(np.where(rng.random(n) < 0.05, np.nan, rng.normal(0, 1, n).round(2)), mask=rng.random(n) < 0.05),
        "i": pa.array(rng.integers(-2**63, 2**63 - 1, n), mask=rng.random(n) < 0.05),
        "s": pa.array(rng.integers(-3, 3, n).astype(np.int8)),
        "u": pa.array(rng.integers(0, 2**64 - 1, n, dtype=np.uint64)),
    })
    dev = vb.DeviceBatch.from_arrow(table, stream)
    for cols, orders in ((["f"], ["ASC"]), (["f"], ["DESC"]), (["i"], ["DESC"]), (["s", "f"], ["DESC", "ASC"]),
                         (["u"], ["ASC"]), (["s", "i", "f"], ["ASC", "ASC", "DESC"])):
        want = O.sort_indices(table, cols, orders)
        got = ops.sort_indices([dev.column(c) for c in cols], [L.DESC if o == "DESC" else L.ASC for o in orders],
                               stream).to_numpy(stream)
        assert np.array_equal(got, want), (cols, orders)


def test_sort_rejects_boolean_key(vb):
    t = pa.table({"b": pa.array([True, False]), "x": pa.array([1, 2])})
    with pytest.raises(RuntimeError):
        _sorted_rows(vb, t, ["b"], ["ASC"])
