"""Headline-size parity (-m gpu): BASELINE.json's configs at their FULL sizes, checked against an
independent CPU computation over the same rows -- not against another of this repo's kernels.

The synthetic columns are a pure function of the row number (vinum_b200/datagen.py), so any row
range is regenerated on the host with NumPy and reduced there in chunks by a fork pool:
  * north-star (1e9 rows): all 1000 (key, COUNT(*)) pairs bit-exact, SUM(f1) within 1e-6 relative
    (BASELINE.json's tolerance), against np.bincount over the selected rows -- the quantity
    SingleNumericalHashAggregate computes (single_numerical_hash_aggregate.cpp:15-46,
    agg_funcs.h:97-127,280-317);
  * C3 (1e9 rows, int32 key, no predicate): the same;
  * C2 (1e8 rows, 4 columns): every output column equals NumPy boolean indexing, bit for bit
    (RecordBatch.filter keeps input order, record_batch.py:85-90);
  * C4 (1e8 rows): the permutation equals np.argsort(kind="stable") on a LOW-cardinality key, where
    almost every comparison is a tie and any instability shows, and on the float64 key DESC through
    the order-preserving code (ties keep input order in both directions, sort.cpp:22-48).
Sizes shrink with VK_TEST_FULL_ROWS / VK_TEST_SCALE_ROWS for quick local runs.
"""
import multiprocessing as mp
import os

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu
FLOAT_RTOL = 1e-6
CHUNK = 1 << 24


@pytest.fixture(scope="module")
def vb(stream):
    import vinum_b200
    return vinum_b200


def _groupby_chunk(args):
    key, val, pred, row0, rows = args
    from vinum_b200.datagen import host_column
    k = host_column(key, row0, rows).astype(np.int64)
    v = host_column(val, row0, rows)
    if pred is not None:
        m = host_column(pred, row0, rows) > 0.5
        k, v = k[m], v[m]
    return np.bincount(k, minlength=1000), np.bincount(k, weights=v, minlength=1000)


def _host_groupby(key, val, pred, n):
    jobs = [(key, val, pred, r0, min(CHUNK, n - r0)) for r0 in range(0, n, CHUNK)]
    cnt = np.zeros(1000, dtype=np.int64)
    sm = np.zeros(1000, dtype=np.float64)
    with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 16)) as pool:
        for c, s in pool.imap_unordered(_groupby_chunk, jobs):
            cnt += c
            sm += s
    return cnt, sm


@pytest.mark.parametrize("config", ["northstar", "c3"])
def test_group_by_at_1e9_rows_vs_numpy(vb, stream, config):
    from vinum_b200 import datagen, ops, _lib as L
    n = int(os.environ.get("VK_TEST_FULL_ROWS", 1_000_000_000))
    key, pred_col = ("i0", "f0") if config == "northstar" else ("k32", None)
    names = [key, "f1"] + ([pred_col] if pred_col else [])
    dev = datagen.device_table(names, 0, n, stream=stream)
    pred = ops.Predicate.compare(dev.column(pred_col), ">", 0.5) if pred_col else None
    agg = vb.Aggregator([pa.int64() if key == "i0" else pa.int32()], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())])
    agg.profile(True)
    agg.update([dev.column(key)], [None, dev.column("f1")], pred, stream)
    keys, kv, cnt, lo, hi, valid = agg.result_raw(stream)
    ms, launches, rows = agg.profile_read(1)
    assert agg.last_path == 1 and launches == 2 and rows >= n - 4096   # learning launch + ONE main launch
    agg.close()
    del dev
    want_cnt, want_sum = _host_groupby(key, "f1", pred_col, n)
    order = np.argsort(keys[0].view(np.int64))
    assert np.array_equal(keys[0].view(np.int64)[order], np.arange(1000)) and kv.all()
    assert np.array_equal(cnt[order].astype(np.int64), want_cnt)            # bit-exact, all 1000 groups
    assert np.array_equal(lo[0][order].astype(np.int64), want_cnt)          # COUNT(*) as a function result
    got_sum = lo[1].view(np.float64)[order]
    assert np.allclose(got_sum, want_sum, rtol=FLOAT_RTOL, atol=0), float(np.max(np.abs(got_sum - want_sum) / np.abs(want_sum)))
    assert valid.all()


def test_filter_at_1e8_rows_vs_numpy_indexing(vb, stream):
    """C2: SELECT * FROM t WHERE f0 > 0.5 over {i1, i2, f0, f1}."""
    from vinum_b200 import datagen, ops
    n = int(os.environ.get("VK_TEST_SCALE_ROWS", 100_000_000))
    names = ["i1", "i2", "f0", "f1"]
    dev = datagen.device_table(names, 0, n, stream=stream)
    out = ops.filter_batch(dev, ops.Predicate.compare(dev.column("f0"), ">", 0.5), stream)
    m = datagen.host_column("f0", 0, n) > 0.5
    assert out.num_rows == int(m.sum())
    for name in names:
        got = out.column(name).to_numpy(stream)
        want = datagen.host_column(name, 0, n)[m]
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), name
        del got, want


@pytest.mark.parametrize("key,order", [("i0", "ASC"), ("i0", "DESC"), ("f3", "DESC")])
def test_sort_at_1e8_rows_vs_numpy_stable_argsort(vb, stream, key, order):
    """C4.  i0 has 1000 distinct values in 1e8 rows: the order inside every run of equal keys is
    input order, under DESC too (sort.cpp:22-29 + Arrow's stable SortIndices)."""
    from vinum_b200 import datagen, ops, _lib as L
    n = int(os.environ.get("VK_TEST_SCALE_ROWS", 100_000_000))
    col = datagen.device_column(key, 0, n, stream=stream)
    idx, sorted0 = ops.sort_indices_keys([col], [L.DESC if order == "DESC" else L.ASC], stream)
    got = idx.to_numpy(stream)
    host = datagen.host_column(key, 0, n)
    if host.dtype.kind == "f":
        code = host.view(np.int64).copy()      # f3 >= 0: the bit pattern orders like the value
        assert (host >= 0).all() and not np.isnan(host).any()
    else:
        code = host
    want = np.argsort(-code if order == "DESC" else code, kind="stable")
    assert np.array_equal(got, want)
    assert np.array_equal(sorted0.to_numpy(stream).view(np.uint64), host[want].view(np.uint64))
