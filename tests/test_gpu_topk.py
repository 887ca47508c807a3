"""GPU parity of the device top-k (`ORDER BY ... LIMIT k`): vk_topk_candidates +
ops.sort_top against the prefix of the full stable sort, and through Table.sql().

Written without a GPU at the end of round 1 and green on first contact (gpurun_out/final_experimental.log,
42 passed); the engine uses the select by default for tables of >= 2^20 rows, VINUM_B200_TOPK=0 turns it off.
"""
import numpy as np
import pyarrow as pa
import pytest

from oracle import vinum_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vb(stream):
    import vinum_b200
    return vinum_b200


def _table(n, seed):
    rng = np.random.default_rng(seed)
    f = rng.random(n) * 1e6 - 2e5
    f[rng.integers(0, n, 40)] = np.nan
    f[rng.integers(0, n, 40)] = 0.0
    f[rng.integers(0, n, 40)] = -0.0
    f[rng.integers(0, n, 40)] = np.inf
    f[rng.integers(0, n, 40)] = -np.inf
    return pa.table({
        "f": pa.array(f),
        "fn": pa.array(f, mask=rng.random(n) < 0.01),          # float key with NULLs (in-band code)
        "i": pa.array(rng.integers(-2**62, 2**62, n)),
        "dup": pa.array(rng.integers(0, 50, n).astype(np.int64)),   # heavy ties: stability matters
        "in": pa.array(rng.integers(0, 1000, n).astype(np.int32), mask=rng.random(n) < 0.01),  # int NULLs: falls back
    })


@pytest.mark.parametrize("keys,orders", [
    (["f"], ["ASC"]), (["f"], ["DESC"]), (["fn"], ["ASC"]), (["fn"], ["DESC"]), (["i"], ["ASC"]), (["i"], ["DESC"]),
    (["dup"], ["ASC"]), (["dup", "f"], ["DESC", "ASC"]), (["in"], ["ASC"]), (["f", "dup"], ["ASC", "DESC"]),
])
@pytest.mark.parametrize("k", [1, 7, 1000, 40_000])
def test_sort_top_equals_prefix_of_full_sort(vb, stream, keys, orders, k):
    from vinum_b200 import _lib as L, ops
    n = 1_200_003
    table = _table(n, 11)
    dev = [vb.DeviceColumn.from_arrow(table.column(c).combine_chunks(), stream) for c in keys]
    codes = [L.DESC if o == "DESC" else L.ASC for o in orders]
    got = ops.sort_top(dev, codes, k, stream).to_numpy(stream)
    want = O.sort_indices(table, keys, orders)[:k]
    assert np.array_equal(got, want)


def test_topk_candidates_contract(vb, stream):
    """Candidates are a superset of the first k rows, few, and the select declines when it cannot pay."""
    from vinum_b200 import _lib as L, ops
    n = 2_000_000
    table = _table(n, 5)
    f = vb.DeviceColumn.from_arrow(table.column("f").combine_chunks(), stream)
    cand = ops.topk_candidates(f, L.DESC, 100, stream)
    assert cand is not None and cand.length <= max(1 << 16, n // 4)
    ids = set(cand.to_numpy(stream).tolist())
    assert len(ids) == cand.length
    assert set(O.sort_indices(table, ["f"], ["DESC"])[:100].tolist()) <= ids
    # k too large, integer key with NULLs, every key equal: the caller is told to sort everything
    assert ops.topk_candidates(f, L.ASC, n // 2, stream) is None
    nul = vb.DeviceColumn.from_arrow(table.column("in").combine_chunks(), stream)
    assert ops.topk_candidates(nul, L.ASC, 10, stream) is None
    same = vb.DeviceColumn.from_arrow(pa.array(np.full(n, 42, dtype=np.int64)), stream)
    assert ops.topk_candidates(same, L.ASC, 10, stream) is None


def test_sql_order_by_limit_with_topk(vb, stream, monkeypatch):
    from vinum_b200 import datagen
    n = 3_000_000
    host = datagen.host_table(["i0", "f0", "f1", "i2"], 0, n)
    tbl = vb.Table.from_arrow(host)
    queries = ["SELECT i2, f1 FROM t ORDER BY f1 DESC LIMIT 25",
               "SELECT i2, i0, f1 FROM t WHERE f0 > 0.25 ORDER BY i0, f1 DESC LIMIT 40 OFFSET 3",
               "SELECT i2 FROM t ORDER BY i0 LIMIT 10"]
    for q in queries:
        monkeypatch.setenv("VINUM_B200_TOPK", "0")       # full sort, then slice
        want = tbl.sql(q).to_arrow()
        assert not tbl.last_stats.get("sort_topk")
        monkeypatch.delenv("VINUM_B200_TOPK")             # default: radix select + sort of the candidates
        got = tbl.sql(q).to_arrow()
        assert tbl.last_stats.get("sort_topk")
        assert got.equals(want), q
