"""GPU parity on hostile inputs and on every selectable kernel (-m gpu).

The headline benchmark feeds the fused aggregate the friendliest input there is (dense keys 0..999,
uniform).  These cases are the ones SURVEY 7.3 calls hard -- one hot key, Zipf, keys that leave the
range the learning launch saw, the table's own sentinel value as a key, more groups than a CTA holds
showing up late, float keys with NaN payloads and both zeros -- each checked against the REFERENCE's
own C++ operators (oracle/_ref: single_numerical_hash_aggregate.cpp:15-46 unmodified) on the same
Arrow input, through every kernel path.  The checker falls back to the oracle's restatement when
oracle/_ref is not built.  Also: every run-time option of include/vinum_b200.h's vk_set_option is
exercised here, so no kernel ships untested.
"""
import zlib

import numpy as np
import pyarrow as pa
import pytest

from golden_util import assert_tables_match
from oracle import ref as R
from oracle import vinum_oracle as O

pytestmark = pytest.mark.gpu
FLOAT_RTOL = 1e-6
FUNCS = [("COUNT_STAR", "", "c"), ("SUM", "v", "s")]


@pytest.fixture(scope="module")
def vb(stream):
    import vinum_b200
    return vinum_b200


def _reference_groupby(table, where=None):
    """Reference chain over the host table (the compiled reference when present)."""
    if where is not None:
        if R.ref_lib() is not None:
            return pa.Table.from_batches([R.ref_filter_hash_aggregate(table, where[0], where[1], where[2], ["k"], FUNCS,
                                                                      batch_size=100_000)])
        return pa.Table.from_batches([O.filter_hash_aggregate(table, where[0], where[1], where[2], ["k"], FUNCS)])
    batches = table.to_batches(max_chunksize=100_000)
    if R.ref_lib() is not None:
        return pa.Table.from_batches([R.ref_aggregate(batches, ["k"], ["k"], FUNCS)])
    return pa.Table.from_batches([O.hash_aggregate(batches, ["k"], ["k"], FUNCS)])


def _device_groupby(vb, stream, table, where=None, opts=None, chunks=1):
    from vinum_b200 import _lib as L, ops
    with vb.options(**(opts or {})):
        agg = vb.Aggregator([table.schema.field("k").type], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, table.schema.field("v").type)])
    n = table.num_rows
    step = -(-n // chunks)
    step += step % 2
    paths = []
    for lo in range(0, n, step):
        part = table.slice(lo, min(step, n - lo)).combine_chunks()
        k = vb.DeviceColumn.from_arrow(part.column("k").chunk(0), stream)
        v = vb.DeviceColumn.from_arrow(part.column("v").chunk(0), stream)
        pred = None
        if where is not None:
            pred = ops.Predicate.compare(vb.DeviceColumn.from_arrow(part.column(where[0]).chunk(0), stream), where[1], where[2])
        agg.update([k], [None, v], pred, stream)
        paths.append(agg.last_path)
    keys, aggs = agg.result_arrays(stream)
    agg.close()
    return pa.table([keys[0]] + aggs, names=["k", "c", "s"]), paths


PATH_OPTS = {
    "auto": {},
    "hash": {"AGG_DIRECT": 0},                                     # read-only cuckoo dictionary after the learning launch
    "hash_insert": {"AGG_DIRECT": 0, "AGG_DICT": 0},               # insert-as-you-go CTA table
    "hash_small": {"AGG_DIRECT": 0, "AGG_LOG2S": 10, "AGG_WARPS": 12},
    "tag_arbitration": {"AGG_ENTRY": 0},
    "no_hot_group_step": {"AGG_HOT": 0},
    "hot_group_step_always": {"AGG_HOT": 2},
    "split_hot_group_step_always": {"AGG_ENTRY": 2, "AGG_HOT": 2},
    "split_entries": {"AGG_ENTRY": 2},                             # SUM array + {tag | COUNT} array
    "general": {"AGG_NOFAST": 1},                                  # global table, four rows per thread (agg_wide_kernel)
    "general_one_row": {"AGG_NOFAST": 1, "AGG_WIDE": 0},           # agg_general_kernel
    "partitioned": {"AGG_NOFAST": 1, "AGG_PARTITION": 2},          # scatter into table slices + slice-by-slice update
    "general_hot": {"AGG_NOFAST": 1, "AGG_HOT": 2},                # global-table kernel that merges a warp's updates of one slot
    "partitioned_hot": {"AGG_NOFAST": 1, "AGG_PARTITION": 2, "AGG_HOT": 2},
    "no_partition": {"AGG_PARTITION": 0},
}


def _keys(kind: str, n: int, rng) -> np.ndarray:
    if kind == "one_hot_100":
        return np.full(n, 7, dtype=np.int64)
    if kind == "one_hot_90":
        k = rng.integers(0, 1000, n)
        k[rng.random(n) < 0.9] = 123
        return k.astype(np.int64)
    if kind == "zipf":
        return (rng.zipf(1.2, n) % 5000).astype(np.int64)
    if kind == "leaves_window":
        # the learning launch (2^16 rows here) sees 0..99; later rows jump far outside any window around it
        k = rng.integers(0, 100, n).astype(np.int64)
        k[1 << 17:] += rng.choice(np.array([0, 10**6, -10**9, 2**40], dtype=np.int64), n - (1 << 17))
        return k
    if kind == "sentinel":
        # 0xFFFF...F is the empty-slot tag of both the CTA table and the global table
        k = rng.integers(-3, 3, n).astype(np.int64)
        k[::5] = -1
        return k
    if kind == "late_groups":
        # 50 groups for almost all of the input, 40 000 new ones in the last rows (more than a CTA holds)
        k = rng.integers(0, 50, n).astype(np.int64)
        k[-60_000:] = rng.integers(10_000, 50_000, 60_000)
        return k
    if kind == "wide_sparse":
        return (rng.integers(0, 900, n).astype(np.int64) * 2654435761) - 2**40
    raise ValueError(kind)


@pytest.mark.parametrize("path", sorted(PATH_OPTS))
@pytest.mark.parametrize("kind", ["one_hot_100", "one_hot_90", "zipf", "leaves_window", "sentinel", "late_groups",
                                  "wide_sparse"])
def test_group_by_hostile_key_distributions(vb, stream, kind, path):
    rng = np.random.default_rng(zlib.crc32(kind.encode()))
    n = 700_000
    table = pa.table({"k": _keys(kind, n, rng), "v": rng.normal(0, 100, n), "p": rng.random(n)})
    want = _reference_groupby(table, ("p", ">", 0.25))
    got, paths = _device_groupby(vb, stream, table, ("p", ">", 0.25), dict(AGG_LEARN_LOG2=16, **PATH_OPTS[path]))
    assert_tables_match(got, want, key_cols=["k"], rtol=FLOAT_RTOL)
    if path == "general":
        assert paths == [2]


def _many_keys(kind: str, n: int, rng) -> np.ndarray:
    if kind == "uniform_2e6":        # the estimate g = G (1 - exp(-n / G)) is exact in expectation
        return rng.integers(0, 2_000_000, n).astype(np.int64)
    if kind == "all_distinct":       # g == n: no estimate, every row left is taken for a new group
        return rng.permutation(n).astype(np.int64) * 7919
    if kind == "sorted_runs":        # new groups arrive at a constant rate: the even-draw estimate is far too low
        return (np.arange(n, dtype=np.int64) // 3) - 12345
    if kind == "skewed_long_tail":   # half the rows on 100 keys, the rest spread thin
        k = rng.integers(0, 3_000_000, n).astype(np.int64)
        hot = rng.random(n) < 0.5
        k[hot] = rng.integers(0, 100, int(hot.sum()))
        return k
    raise ValueError(kind)


@pytest.mark.parametrize("path", ["auto", "general", "general_one_row", "partitioned", "partitioned_hot", "no_partition"])
@pytest.mark.parametrize("chunks", [1, 3])
@pytest.mark.parametrize("kind", ["uniform_2e6", "all_distinct", "sorted_runs", "skewed_long_tail"])
def test_group_by_more_groups_than_the_first_table_holds(vb, stream, kind, chunks, path):
    """Millions of groups: the global table (2 Mi slots at first) is sized from the estimate the host makes at
    every chunk boundary, grown when the estimate was low, and rows that found it full are replayed; nothing
    is lost or counted twice whichever of these happens (unordered_map semantics of
    single_numerical_hash_aggregate.cpp:15-46)."""
    rng = np.random.default_rng(zlib.crc32(kind.encode()) + chunks)
    n = 4_300_007
    table = pa.table({"k": _many_keys(kind, n, rng), "v": rng.normal(0, 100, n), "p": rng.random(n)})
    want = _reference_groupby(table, ("p", ">", 0.25))
    got, paths = _device_groupby(vb, stream, table, ("p", ">", 0.25), dict(AGG_LEARN_LOG2=16, **PATH_OPTS[path]), chunks=chunks)
    assert_tables_match(got, want, key_cols=["k"], rtol=FLOAT_RTOL)
    assert paths[-1] in (2, 4)
    if path.startswith("partitioned"):
        assert paths == [4] * chunks


@pytest.mark.parametrize("path", ["partitioned", "partitioned_hot", "general", "general_hot", "general_one_row"])
@pytest.mark.parametrize("pred_kind", ["none", "f64_cmp", "i64_cmp", "i32_cmp", "mask", "expr"])
def test_global_table_plans_every_predicate_kind_and_cell(vb, stream, pred_kind, path):
    """The plans behind the shared-memory kernel (partitioned scatter + reduce, four rows per thread, one row per
    thread) under every predicate specialisation (PK_NONE / F64_VEC / I64_VEC / GENERIC / MASK / fused expression),
    an int32 key, and MIN(int64) / SUM(int32) / AVG(float32) accumulators (CELL_MAXORD, CELL_ADD_I64, CELL_ADD_F64
    with a widened float32), in two ragged updates."""
    from vinum_b200 import ops, _lib as L
    import pyarrow.compute as pc
    rng = np.random.default_rng(zlib.crc32(pred_kind.encode()))
    n = 600_011
    host = {"k": rng.integers(-20_000, 20_000, n).astype(np.int32), "a": rng.integers(-1000, 1000, n).astype(np.int32),
            "x": rng.normal(0, 50, n), "y": rng.random(n).astype(np.float32), "w": rng.integers(-10**12, 10**12, n),
            "q": rng.integers(-5, 5, n).astype(np.int32)}
    sel = {"none": np.ones(n, bool), "f64_cmp": host["x"] > 3.0, "i64_cmp": host["w"] <= 0, "i32_cmp": host["q"] != 2,
           "mask": (host["x"] > 0) & (host["q"] < 3), "expr": host["x"] * 2.0 > host["w"] / 1e10}[pred_kind]
    with vb.options(AGG_LEARN_LOG2=14, **PATH_OPTS[path]):
        # three value columns and three accumulator cells: the most the fused / partitioned plans take
        agg = vb.Aggregator([pa.int32()], [(L.AGG_COUNT_STAR, None), (L.AGG_MIN, pa.int64()), (L.AGG_SUM, pa.int32()),
                                           (L.AGG_AVG, pa.float32())])
    cut = 250_002
    for lo, hi in ((0, cut), (cut, n)):
        d = {c: vb.DeviceColumn.from_numpy(np.ascontiguousarray(v[lo:hi]), stream) for c, v in host.items()}
        pred = {"none": lambda: None,
                "f64_cmp": lambda: ops.Predicate.compare(d["x"], ">", 3.0),
                "i64_cmp": lambda: ops.Predicate.compare(d["w"], "<=", 0),
                "i32_cmp": lambda: ops.Predicate.compare(d["q"], "!=", 2),
                "mask": lambda: ops.Predicate.from_mask(ops.mask_and(ops.compare(d["x"], ">", 0.0, stream), ops.compare(d["q"], "<", 3, stream), stream)),
                "expr": lambda: ops.Predicate.expr([(None, d["x"]), ("*", 2.0)], ">", [(None, d["w"]), ("/", 1e10)])}[pred_kind]()
        agg.update([d["k"]], [None, d["w"], d["a"], d["y"]], pred, stream)
        assert agg.last_path == (4 if path.startswith("partitioned") else 2)
    keys, aggs = agg.result_arrays(stream)
    agg.close()
    got = pa.table([keys[0]] + aggs, names=["k", "c", "mn", "sa", "ay"]).sort_by("k")
    t = pa.table({c: v[sel] for c, v in host.items()})
    want = t.group_by("k", use_threads=False).aggregate([("k", "count"), ("w", "min"), ("a", "sum")]).sort_by("k")
    assert got.column("k").to_pylist() == want.column("k").to_pylist()
    assert got.column("c").to_pylist() == want.column("k_count").to_pylist()
    assert got.column("mn").to_pylist() == want.column("w_min").to_pylist()
    assert got.column("sa").to_pylist() == want.column("a_sum").to_pylist()
    inv = np.unique(host["k"][sel], return_inverse=True)[1]
    want_avg = np.bincount(inv, weights=host["y"][sel].astype(np.float64)) / np.bincount(inv)
    np.testing.assert_allclose(got.column("ay").to_numpy().astype(np.float64), want_avg, rtol=1e-5)


@pytest.mark.parametrize("path", ["auto", "hash", "hash_insert", "tag_arbitration", "split_entries", "general", "partitioned"])
def test_group_by_without_predicate_and_streamed_chunks(vb, stream, path):
    """C3's shape (no WHERE), fed in five ragged chunks: state carries across vk_agg_update calls
    (BaseAggregate::Next, base_aggregate.cpp:23-45)."""
    rng = np.random.default_rng(5)
    n = 900_001
    table = pa.table({"k": rng.integers(0, 1000, n).astype(np.int32), "v": rng.normal(0, 1, n)})
    want = _reference_groupby(table)
    got, _ = _device_groupby(vb, stream, table, None, dict(AGG_LEARN_LOG2=14, **PATH_OPTS[path]), chunks=5)
    assert_tables_match(got, want, key_cols=["k"], rtol=FLOAT_RTOL)


@pytest.mark.parametrize("path", ["auto", "hash", "hash_insert", "general", "partitioned"])
def test_float_keys_nan_payloads_and_signed_zero(vb, stream, path):
    """Float keys group by BIT PATTERN (FloatArrayIter::floatToInt, array_iterators.h:239-248):
    -0.0 and +0.0 are two groups, NaNs with different payloads are different groups."""
    rng = np.random.default_rng(9)
    n = 400_000
    base = np.array([0.0, -0.0, 1.5, -1.5, np.inf, -np.inf], dtype=np.float64)
    nan_a = np.array([0x7FF8000000000000], dtype=np.uint64).view(np.float64)
    nan_b = np.array([0x7FF8000000000001], dtype=np.uint64).view(np.float64)
    nan_c = np.array([0xFFF8000000000000], dtype=np.uint64).view(np.float64)
    pool = np.concatenate([base, nan_a, nan_b, nan_c])
    k = pool[rng.integers(0, len(pool), n)]
    table = pa.table({"k": pa.array(k), "v": rng.normal(0, 1, n), "p": rng.random(n)})
    got, _ = _device_groupby(vb, stream, table, ("p", "<=", 0.8), dict(AGG_LEARN_LOG2=14, **PATH_OPTS[path]))
    want = _reference_groupby(table, ("p", "<=", 0.8))
    assert got.num_rows == want.num_rows == len(pool)
    # compare by bit pattern: NaN != NaN under ordinary equality
    gk = got.column("k").to_numpy().view(np.uint64)
    wk = want.column("k").to_numpy().view(np.uint64)
    go, wo = np.argsort(gk), np.argsort(wk)
    assert np.array_equal(gk[go], wk[wo])
    assert np.array_equal(got.column("c").to_numpy()[go], want.column("c").to_numpy()[wo])
    assert np.allclose(got.column("s").to_numpy()[go], want.column("s").to_numpy()[wo], rtol=FLOAT_RTOL, atol=0)


# ---------------------------------------------------------------- options ----
def _c2_table(n, seed=0):
    rng = np.random.default_rng(seed)
    return pa.table({"i1": rng.integers(-2**40, 2**40, n), "i2": np.arange(n, dtype=np.int64), "f0": rng.random(n),
                     "f1": rng.normal(0, 1000, n), "s": rng.integers(-100, 100, n).astype(np.int16)})


@pytest.mark.parametrize("opts", [dict(), dict(FILTER_STAGE=0), dict(FILTER_PF=0), dict(FILTER_PF=1), dict(FILTER_STAGE=0, FILTER_PF=0)], ids=str)
@pytest.mark.parametrize("n", [1, 2047, 2048, 2049, 300_001])
def test_filter_every_option_vs_numpy_indexing(vb, stream, opts, n):
    """RecordBatch.filter (record_batch.py:85-90): every column compacted, input order kept -- the
    staged-predicate kernel, the plain one, both tile sizes, with and without the L2 prefetch."""
    from vinum_b200 import ops
    table = _c2_table(n)
    dev = vb.DeviceBatch.from_arrow(table, stream)
    for op, c in ((">", 0.5), ("<=", 0.01), (">=", 0.0), ("<", 0.0)):
        with vb.options(**opts):
            out = ops.filter_batch(dev, ops.Predicate.compare(dev.column("f0"), op, c), stream)
        m = {">": np.greater, "<=": np.less_equal, ">=": np.greater_equal, "<": np.less}[op](table.column("f0").to_numpy(), c)
        assert out.num_rows == int(m.sum())
        for name in table.column_names:
            assert np.array_equal(out.column(name).to_numpy(stream), table.column(name).to_numpy()[m]), (name, op, c)
    # int64 predicate column that is also an output (the staged path's other instantiation)
    with vb.options(**opts):
        out = ops.filter_batch(dev, ops.Predicate.compare(dev.column("i1"), ">", 0), stream)
    m = table.column("i1").to_numpy() > 0
    for name in table.column_names:
        assert np.array_equal(out.column(name).to_numpy(stream), table.column(name).to_numpy()[m]), name


@pytest.mark.parametrize("opts", [dict(), dict(SORT_FUSE_LAST=0), dict(SORT_PREP=0), dict(SORT_PREP=2)], ids=str)
def test_sort_every_option_vs_reference(vb, stream, opts):
    """Sort::Sorted (sort.cpp:15-63): identical permutation, ties / NaN / NULL / both zeros included,
    multi-key with integer NULLs, for every selectable prepare / last-pass kernel."""
    from vinum_b200 import ops, _lib as L
    rng = np.random.default_rng(21)
    n = 200_003
    f = rng.integers(-50, 50, n).astype(np.float64) / 4
    f[rng.random(n) < 0.02] = np.nan
    f[rng.random(n) < 0.02] = -0.0
    table = pa.table({
        "f": pa.array(f, mask=rng.random(n) < 0.03),
        "i": pa.array(rng.integers(-5, 5, n), mask=rng.random(n) < 0.05),
        "u": pa.array(rng.integers(0, 2**63, n, dtype=np.uint64) * 2),
        "d": rng.normal(0, 1, n),
    })
    dev = vb.DeviceBatch.from_arrow(table, stream)
    for cols, orders in ((["f"], ["DESC"]), (["i", "f"], ["ASC", "DESC"]), (["u"], ["DESC"]), (["d"], ["ASC"]),
                         (["f", "i", "d"], ["ASC", "DESC", "DESC"])):
        with vb.options(**opts):
            idx = ops.sort_indices([dev.column(c) for c in cols], [L.DESC if o == "DESC" else L.ASC for o in orders], stream)
        want = O.sort_indices(table, cols, orders)
        assert np.array_equal(idx.to_numpy(stream), want), (cols, orders)


@pytest.mark.parametrize("dtype", ["f64", "i64", "u64"])
@pytest.mark.parametrize("order", ["ASC", "DESC"])
def test_sorted_key_column_from_last_pass_is_bit_identical_to_take(vb, stream, dtype, order):
    """vk_sort_indices_keys: the first key column comes out of the last radix pass; it must equal
    Take(key, SortIndices) bit for bit (sort.cpp:40-48) -- NaN payloads and the sign of zero included."""
    from vinum_b200 import ops, _lib as L
    rng = np.random.default_rng(33)
    n = 150_001
    if dtype == "f64":
        k = rng.integers(-20, 20, n).astype(np.float64) / 2
        bits = k.view(np.uint64).copy()
        sel = rng.random(n)
        bits[sel < 0.05] = 0x8000000000000000           # -0.0
        bits[(sel >= 0.05) & (sel < 0.08)] = 0x7FF8000000000123   # NaN with a payload
        bits[(sel >= 0.08) & (sel < 0.10)] = 0xFFF8000000000000   # negative NaN
        k = bits.view(np.float64)
    elif dtype == "i64":
        k = rng.integers(-2**62, 2**62, n)
        k[:10] = [-2**63, 2**63 - 1, 0, -1, 1, 0, 0, -2**63, 2**63 - 1, 5]
    else:
        k = rng.integers(0, 2**64 - 1, n, dtype=np.uint64)
        k[:4] = [0, 2**64 - 1, 2**63, 2**63 - 1]
    second = rng.integers(0, 3, n).astype(np.int64)
    kd, sd = vb.DeviceColumn.from_numpy(k, stream), vb.DeviceColumn.from_numpy(second, stream)
    o = L.DESC if order == "DESC" else L.ASC
    for keys, orders in (([kd], [o]), ([kd, sd], [o, L.ASC])):
        idx, sorted0 = ops.sort_indices_keys(keys, orders, stream)
        assert sorted0 is not None
        plain = ops.sort_indices(keys, orders, stream)
        assert np.array_equal(idx.to_numpy(stream), plain.to_numpy(stream))
        want = k[plain.to_numpy(stream)]
        assert np.array_equal(sorted0.to_numpy(stream).view(np.uint64), want.view(np.uint64))
    # a constant first key: no radix pass runs for it, the library gathers instead
    const = vb.DeviceColumn.from_numpy(np.full(1000, k[0]), stream)
    tail = vb.DeviceColumn.from_numpy(rng.integers(0, 50, 1000), stream)
    idx, sorted0 = ops.sort_indices_keys([const, tail], [o, L.ASC], stream)
    assert np.array_equal(sorted0.to_numpy(stream).view(np.uint64), np.full(1000, k[0]).view(np.uint64))


@pytest.mark.parametrize("opts", [dict(CMP_FAST=0), dict(CMP_FAST=2), dict(ARITH_FAST=0), dict(ARITH_FAST=2),
                                  dict(ARITH_FAST=4)], ids=str)
def test_elementwise_options_vs_numpy(vb, stream, opts):
    from vinum_b200 import ops
    rng = np.random.default_rng(2)
    n = 100_003
    a, b = rng.normal(0, 1, n), rng.normal(0, 1, n)
    i = rng.integers(-2**40, 2**40, n)
    da, db, di = (vb.DeviceColumn.from_numpy(x, stream) for x in (a, b, i))
    with vb.options(**opts):
        assert np.array_equal(ops.compare(da, ">", 0.25, stream).to_numpy(stream).astype(bool), a > 0.25)
        assert np.array_equal(ops.compare(da, "<=", db, stream).to_numpy(stream).astype(bool), a <= b)
        assert np.array_equal(ops.compare(di, "!=", 7, stream).to_numpy(stream).astype(bool), i != 7)
        assert np.array_equal(ops.arith("+", da, db, stream).to_numpy(stream).view(np.uint64), (a + b).view(np.uint64))
        assert np.array_equal(ops.arith("*", da, 2.5, stream).to_numpy(stream).view(np.uint64), (a * 2.5).view(np.uint64))
        assert np.array_equal(ops.arith("*", di, 3, stream).to_numpy(stream), i * 3)


@pytest.mark.parametrize("opts", [dict(ONEGROUP_FAST=0), dict(ONEGROUP_FAST=2), dict(ONEGROUP_FAST=4)], ids=str)
def test_one_group_options_vs_reference(vb, stream, opts):
    """OneGroupAggregate (one_group_aggregate.cpp:9-26) through each un-grouped reduction kernel."""
    from vinum_b200 import _lib as L, ops
    rng = np.random.default_rng(4)
    n = 500_001
    v, p, i = rng.normal(0, 10, n), rng.random(n), rng.integers(-2**50, 2**50, n)
    dv, dp, di = (vb.DeviceColumn.from_numpy(x, stream) for x in (v, p, i))
    with vb.options(**opts):
        agg = vb.Aggregator([], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64()), (L.AGG_MAX, pa.float64()),
                                 (L.AGG_SUM, pa.int64()), (L.AGG_MIN, pa.int64())])
        agg.update([], [None, dv, dv, di, di], ops.Predicate.compare(dp, ">", 0.5), stream)
        _, aggs = agg.result_arrays(stream)
    m = p > 0.5
    assert aggs[0][0].as_py() == int(m.sum())
    assert np.isclose(aggs[1][0].as_py(), v[m].sum(), rtol=FLOAT_RTOL, atol=0)
    assert aggs[2][0].as_py() == v[m].max()
    assert aggs[3][0].as_py() == int(i[m].sum())
    assert aggs[4][0].as_py() == int(i[m].min())


def test_unknown_option_is_an_error_and_reset_restores_defaults(vb):
    with pytest.raises(vb.VinumB200Error):
        vb.set_option("NO_SUCH_KNOB", 1)
    before = vb.get_option("SORT_FUSE_LAST")
    vb.set_option("SORT_FUSE_LAST", 1 - before)
    assert vb.get_option("SORT_FUSE_LAST") == 1 - before
    vb.lib.vk_reset_options()
    assert vb.get_option("SORT_FUSE_LAST") == before


def test_pageable_and_pinned_ingest_agree(vb, stream):
    """vk_memcpy_h2d_auto: a pageable source goes through the pinned bounce pool (ragged tail piece
    included), a pinned one is DMA'd directly; both must land the same bytes."""
    rng = np.random.default_rng(8)
    n = (9 << 20) // 8 + 12345          # > 2 pieces of 4 MB, ragged
    a = rng.integers(-2**62, 2**62, n)
    pinned = vb.pinned_array(a)
    d1 = vb.DeviceColumn.from_arrow(pa.array(a), stream)
    d2 = vb.DeviceColumn.from_arrow(pa.array(pinned), stream)
    with vb.options(INGEST_STAGED=0):
        d3 = vb.DeviceColumn.from_arrow(pa.array(a), stream)
    for d in (d1, d2, d3):
        assert np.array_equal(d.to_numpy(stream), a)
    assert vb.lib.raw.vk_ingest_threads() >= 1


def test_packed_result_grows_past_its_first_block(vb, stream):
    """vk_agg_result_packed: more groups than the block holds -> the count comes back, the caller retries."""
    from vinum_b200 import _lib as L
    rng = np.random.default_rng(6)
    n = 300_000
    k = rng.integers(0, 50_000, n)
    v = rng.normal(0, 1, n)
    agg = vb.Aggregator([pa.int64()], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())])
    agg.update([vb.DeviceColumn.from_numpy(k, stream)], [None, vb.DeviceColumn.from_numpy(v, stream)], None, stream)
    keys, kv, cnt, lo, hi, valid = agg.result_raw(stream)
    uk, uc = np.unique(k, return_counts=True)
    order = np.argsort(keys[0].view(np.int64))
    assert np.array_equal(keys[0].view(np.int64)[order], uk)
    assert np.array_equal(cnt[order], uc.astype(np.uint64))
    assert np.allclose(lo[1].view(np.float64)[order], np.bincount(np.searchsorted(uk, k), weights=v), rtol=1e-9, atol=1e-9)
