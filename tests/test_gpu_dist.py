"""Multi-rank group-by exchange on ONE GPU (-m gpu): G logical ranks in one process, each with its
own aggregate, stream and exchange window (SURVEY 8e "fake cluster": same kernels as the torchrun
path, the windows are attached directly instead of through cudaIpc).  Checked against the
reference's operators over the concatenation of the shards.
"""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

from golden_util import assert_tables_match
from oracle import ref as R
from oracle import vinum_oracle as O

pytestmark = pytest.mark.gpu
FUNCS = [("COUNT_STAR", "", "c"), ("SUM", "v", "s"), ("MIN", "v", "mn"), ("AVG", "v", "av"), ("SUM", "i", "si")]


@pytest.fixture(scope="module")
def vb(stream):
    import vinum_b200
    return vinum_b200


def _reference(table):
    batches = table.to_batches(max_chunksize=50_000)
    if R.ref_lib() is not None:
        return pa.Table.from_batches([R.ref_aggregate(batches, ["k"], ["k"], FUNCS)])
    return pa.Table.from_batches([O.hash_aggregate(batches, ["k"], ["k"], FUNCS)])


class _Rank:
    def __init__(self, vb, rank, world, cap, words):
        self.stream = vb.Stream()
        h = C.c_void_p()
        vb.lib.vk_peer_create(C.byref(h), rank, world, cap, words)
        self.peer = h


def _run(vb, shards, cap, epochs=1):
    """Every shard aggregated by its own rank; ranks 1.. send to rank 0, rank 0 merges.  Returns
    (decision, merged table or None) of the LAST epoch."""
    from vinum_b200 import _lib as L
    world = len(shards)
    spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64()), (L.AGG_MIN, pa.float64()), (L.AGG_AVG, pa.float64()),
            (L.AGG_SUM, pa.int64())]
    words = 1 + 2 + 3 * len(spec)
    ranks = [_Rank(vb, r, world, cap, words) for r in range(world)]
    for a in ranks:
        for r, b in enumerate(ranks):
            if a is not b:
                vb.lib.vk_peer_attach_local(a.peer, r, b.peer)
    pinned = vb.PinnedBuffer(64)
    word = pinned.as_numpy(np.uint64, 8)
    out = None
    try:
        for epoch in range(1, epochs + 1):
            aggs = []
            for rk, shard in zip(ranks, shards):
                agg = vb.Aggregator([pa.int64()], spec)
                if shard.num_rows:
                    cols = {n: vb.DeviceColumn.from_arrow(shard.column(n).chunk(0), rk.stream) for n in ("k", "v", "i")}
                    agg.update([cols["k"]], [None, cols["v"], cols["v"], cols["v"], cols["i"]], None, rk.stream)
                aggs.append(agg)
            # senders first (their kernels wait for the owner's acknowledgement on the device)
            for r in range(1, world):
                vb.lib.vk_agg_peer_send(aggs[r]._h, ranks[r].peer, 0, epoch, ranks[r].stream.ptr)
            vb.lib.vk_agg_peer_merge(aggs[0]._h, ranks[0].peer, epoch, ranks[0].stream.ptr)
            dptr = vb.lib.raw.vk_peer_decision_ptr(ranks[0].peer, epoch)
            raw = aggs[0].result_raw(ranks[0].stream, extra_d2h=(pinned.ptr, dptr, 8))
            decision = int(word[0])
            for r in range(1, world):
                ranks[r].stream.sync()
                # the owner's decision reached every window
                vb.lib.vk_memcpy_d2h(C.c_void_p(pinned.ptr + 8), C.c_void_p(vb.lib.raw.vk_peer_decision_ptr(ranks[r].peer, epoch)),
                                     8, ranks[r].stream.ptr)
                ranks[r].stream.sync()
                assert int(word[1]) == decision
            if decision == 1:
                keys, aggs_out = aggs[0].result_arrays(ranks[0].stream, raw=raw)
                out = pa.table([keys[0]] + aggs_out, names=["k"] + [f[2] for f in FUNCS])
            else:
                out = None
            for a in aggs:
                a.close()
    finally:
        for rk in ranks:
            vb.lib.vk_peer_destroy(rk.peer)
    return decision, out


def _table(rng, n, nkeys):
    return pa.table({"k": rng.integers(-nkeys // 2, nkeys // 2, n), "v": rng.normal(0, 50, n),
                     "i": rng.integers(-2**62, 2**62, n)})


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_exchange_matches_reference(vb, world):
    rng = np.random.default_rng(world)
    n = 240_000
    table = _table(rng, n, 1500)
    bounds = [n * r // world for r in range(world + 1)]
    shards = [table.slice(bounds[r], bounds[r + 1] - bounds[r]).combine_chunks() for r in range(world)]
    decision, got = _run(vb, shards, cap=2048, epochs=3)   # three queries reuse both slot parities
    assert decision == 1
    assert_tables_match(got, _reference(table), key_cols=["k"], rtol=1e-6,
                        float_exact_cols=["mn"])


def test_peer_exchange_with_empty_and_disjoint_shards(vb):
    rng = np.random.default_rng(77)
    a = pa.table({"k": rng.integers(0, 100, 50_000), "v": rng.normal(0, 1, 50_000), "i": rng.integers(-9, 9, 50_000)})
    b = pa.table({"k": rng.integers(1000, 1100, 30_000), "v": rng.normal(0, 1, 30_000), "i": rng.integers(-9, 9, 30_000)})
    empty = a.slice(0, 0)
    for shards in ([a, empty, b], [empty, a, b]):
        decision, got = _run(vb, shards, cap=512)
        assert decision == 1
        assert_tables_match(got, _reference(pa.concat_tables([s for s in shards if s.num_rows])), key_cols=["k"], rtol=1e-6,
                            float_exact_cols=["mn"])


def test_peer_exchange_reports_overflow_unanimously(vb):
    """A rank with more partial groups than a slot holds: nothing is merged, every rank reads decision 2
    (and falls back to the all-to-all repartition in vinum_b200.dist)."""
    rng = np.random.default_rng(5)
    shards = [_table(rng, 20_000, 64), _table(rng, 20_000, 6000)]
    decision, got = _run(vb, shards, cap=256)
    assert decision == 2 and got is None
