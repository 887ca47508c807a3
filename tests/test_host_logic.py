"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header
declares, the host-side logic (type tables, result assembly, scalar conversion,
exchange planning) is right, and the multi-rank repartition plan works over gloo."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pyarrow as pa
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_functions():
    text = (ROOT / "include" / "vinum_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from vinum_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 50
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/vinum_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in vinum_b200/_lib.py"
    assert set(_lib.SIGNATURES) <= set(names), set(_lib.SIGNATURES) - set(names)
    assert raw.vk_abi_version() == 1


def test_library_is_sm100a_only():
    from vinum_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_no_gpu_call_fails_loudly_not_silently():
    """Without a device every compute entry point raises; nothing falls back to the CPU."""
    import vinum_b200 as vb
    try:
        n = vb.device_count()
    except vb.VinumB200Error:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(vb.VinumB200Error):
        vb.DeviceColumn.from_numpy(np.arange(10))


def test_product_never_imports_oracle():
    for p in (ROOT / "vinum_b200").rglob("*.py"):
        src = p.read_text()
        assert "oracle" not in re.sub(r'"""(.|\n)*?"""', "", src), f"{p} references the oracle"
    for p in (ROOT / "vinum_b200" / "csrc").glob("*"):
        assert "oracle" not in p.read_text()


def test_make_scalar():
    from vinum_b200 import _lib as L
    s = L.make_scalar(5)
    assert s.dtype == L.I64 and s.v.i == 5
    s = L.make_scalar(-(2**63))
    assert s.dtype == L.I64 and s.v.i == -(2**63)
    s = L.make_scalar(2**64 - 1)
    assert s.dtype == L.U64 and s.v.u == 2**64 - 1
    s = L.make_scalar(0.5)
    assert s.dtype == L.F64 and s.v.f == 0.5
    s = L.make_scalar(np.float32(0.25))
    assert s.dtype == L.F64 and s.v.f == 0.25
    with pytest.raises(OverflowError):
        L.make_scalar(2**64)
    with pytest.raises(TypeError):
        L.make_scalar("x")


def test_output_type_table_matches_reference_factory():
    """SURVEY A.2 / agg_func_factory.cpp: the type of every aggregate output."""
    from vinum_b200 import _lib as L
    from vinum_b200.aggregate import check_supported, output_type
    ints = [pa.int8(), pa.int16(), pa.int32(), pa.int64()]
    uints = [pa.uint8(), pa.uint16(), pa.uint32(), pa.uint64()]
    for t in ints + uints + [pa.float32(), pa.float64(), pa.timestamp("ms")]:
        assert output_type(L.AGG_COUNT, t) == pa.uint64()
        assert output_type(L.AGG_MIN, t) == t and output_type(L.AGG_MAX, t) == t
    assert output_type(L.AGG_COUNT_STAR, None) == pa.uint64()
    for t in ints:
        assert output_type(L.AGG_SUM, t) == pa.int64()
    for t in uints:
        assert output_type(L.AGG_SUM, t) == pa.uint64()
    assert output_type(L.AGG_SUM, pa.float32()) == pa.float64()
    assert output_type(L.AGG_SUM, pa.time32("ms")) == pa.time32("ms")
    for t in (pa.int8(), pa.int16(), pa.uint8(), pa.uint16()):
        assert output_type(L.AGG_AVG, t) == pa.float32()
    for t in (pa.int32(), pa.int64(), pa.uint32(), pa.uint64(), pa.float32(), pa.float64(), pa.time64("us")):
        assert output_type(L.AGG_AVG, t) == pa.float64()
    for t in (pa.bool_(), pa.date32(), pa.date64(), pa.timestamp("s")):
        with pytest.raises(RuntimeError, match=r"not supported by sum\(\)"):
            check_supported(L.AGG_SUM, t)
        with pytest.raises(RuntimeError, match=r"not supported by avg\(\)"):
            check_supported(L.AGG_AVG, t)


def test_sum_int64_result_assembly_decimal_switch():
    """128-bit (lo, hi) lanes -> int64 unless any group overflows -> whole column decimal128(38,0)."""
    import decimal
    from vinum_b200 import _lib as L
    from vinum_b200.aggregate import _agg_array

    def lanes(vals):
        lo = np.array([v & (2**64 - 1) for v in vals], dtype=np.uint64)
        hi = np.array([(v >> 64) & (2**64 - 1) for v in vals], dtype=np.uint64)
        return lo, hi

    valid = np.array([True, True, False])
    lo, hi = lanes([5, -7, 0])
    arr = _agg_array(L.AGG_SUM, pa.int64(), L.I64, lo, hi, valid)
    assert arr.type == pa.int64() and arr.to_pylist() == [5, -7, None]
    lo, hi = lanes([2**63, -7, 0])
    arr = _agg_array(L.AGG_SUM, pa.int64(), L.I64, lo, hi, valid)
    assert arr.type == pa.decimal128(38, 0)
    assert arr.to_pylist() == [decimal.Decimal(2**63), decimal.Decimal(-7), None]
    # -2^63 exactly fails the reference's int64 cast (huge_int.cpp:341-361)
    lo, hi = lanes([-(2**63), 1, 0])
    assert _agg_array(L.AGG_SUM, pa.int64(), L.I64, lo, hi, valid).type == pa.decimal128(38, 0)
    lo, hi = lanes([2**64 - 1, 3, 0])
    arr = _agg_array(L.AGG_SUM, pa.uint64(), L.U64, lo, hi, valid)
    assert arr.type == pa.uint64() and arr.to_pylist() == [2**64 - 1, 3, None]
    lo, hi = lanes([2**64, 3, 0])
    assert _agg_array(L.AGG_SUM, pa.uint64(), L.U64, lo, hi, valid).type == pa.decimal128(38, 0)
    # an overflowing lane in an INVALID group does not switch the column
    lo, hi = lanes([1, 2, 2**70])
    assert _agg_array(L.AGG_SUM, pa.int64(), L.I64, lo, hi, valid).type == pa.int64()


def test_narrowing_and_temporal_result_arrays():
    from vinum_b200 import _lib as L
    from vinum_b200.aggregate import _agg_array, _key_array
    valid = np.array([True, False])
    lo = np.array([np.int64(-5).astype(np.uint64), 0], dtype=np.uint64)
    arr = _agg_array(L.AGG_MIN, pa.int8(), L.I8, lo, np.zeros(2, np.uint64), valid)
    assert arr.type == pa.int8() and arr.to_pylist() == [-5, None]
    arr = _agg_array(L.AGG_MAX, pa.timestamp("ms"), L.I64, np.array([1611664420588, 0], np.uint64),
                     np.zeros(2, np.uint64), valid)
    assert arr.type == pa.timestamp("ms") and arr.cast(pa.int64()).to_pylist() == [1611664420588, None]
    f = np.array([np.float64(1.5)]).view(np.uint64)
    arr = _agg_array(L.AGG_AVG, pa.int8(), L.I8, f, np.zeros(1, np.uint64), np.array([True]))
    assert arr.type == pa.float32() and arr.to_pylist() == [1.5]
    k = _key_array(np.array([np.float32(0.5).view(np.uint32)], dtype=np.uint64), np.array([True]), pa.float32(), L.F32)
    assert k.type == pa.float32() and k.to_pylist() == [0.5]


def test_arith_result_dtype_follows_numpy():
    from vinum_b200 import _lib as L
    from vinum_b200.ops import result_dtype
    from vinum_b200.device import DeviceColumn

    def col(dt, nulls=False):
        c = DeviceColumn(None, object() if nulls else None, 0, 4, dt, None, 1 if nulls else 0, data_ptr=0)
        return c

    assert result_dtype(L.ADD, col(L.I64), 3) == np.int64
    assert result_dtype(L.ADD, col(L.I32), col(L.I8)) == np.int32
    assert result_dtype(L.DIV, col(L.I64), col(L.I64)) == np.float64
    assert result_dtype(L.MUL, col(L.F32), 0.1) == np.float32
    assert result_dtype(L.MUL, col(L.I16), 2.5) == np.float64
    assert result_dtype(L.ADD, col(L.I64, nulls=True), 1) == np.float64  # NULL -> NaN view
    assert result_dtype(L.NEG, col(L.I8), None) == np.int8
    assert result_dtype(L.ADD, col(L.U8), col(L.U32)) == np.uint32


def test_shard_ranges_cover_rows_exactly():
    from vinum_b200.dist import shard_range
    for n in (0, 1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]


def test_exchange_plan():
    from vinum_b200.dist import exchange_plan
    counts = np.array([[1, 2, 3], [4, 0, 6], [7, 8, 9]])
    so, ss, ro, rs = exchange_plan(counts)
    assert so.tolist() == [[0, 1, 3], [0, 4, 4], [0, 7, 15]]
    assert ss.tolist() == counts.tolist()
    assert rs.tolist() == counts.T.tolist()
    assert ro.tolist() == [[0, 1, 5], [0, 2, 2], [0, 3, 9]]


def test_merge_sorted_runs_is_a_global_stable_sort():
    """dist.merge_sorted_runs (SURVEY 8e host merge): merging per-shard stable sorts in shard order
    equals one stable sort of the whole column, ties and NaN placement included."""
    from vinum_b200.dist import merge_sorted_runs
    rng = np.random.default_rng(3)
    for dtype, desc in [(np.float64, False), (np.float64, True), (np.int64, False), (np.int64, True), (np.uint64, True)]:
        n = 5000
        if dtype == np.float64:
            col = np.round(rng.normal(0, 3, n), 1)
            col[rng.random(n) < 0.02] = np.nan
        else:
            col = rng.integers(0, 40, n).astype(dtype)
        bounds = [0, 1200, 1200, 3100, n]   # one empty shard
        runs_k, runs_i = [], []
        for lo, hi in zip(bounds, bounds[1:]):
            part = col[lo:hi]
            if dtype == np.float64:
                key = np.where(np.isnan(part), np.inf, -part if desc else part)
                order = np.lexsort((key, np.isnan(part)))
            else:
                order = np.argsort(part if not desc else (part.max() - part if len(part) else part), kind="stable")
            runs_k.append(part[order])
            runs_i.append(order.astype(np.int64) + lo)
        k, ids = merge_sorted_runs(runs_k, runs_i, desc)
        if dtype == np.float64:
            key = np.where(np.isnan(col), np.inf, -col if desc else col)
            want = np.lexsort((key, np.isnan(col)))
        else:
            want = np.argsort(col if not desc else col.max() - col, kind="stable")
        assert np.array_equal(ids, want), (dtype, desc)
        assert np.array_equal(k, col[want], equal_nan=True)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from vinum_b200.dist import exchange_plan, all_to_all_records, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
words = 3
# rank r sends (r+1)*(d+1) records to rank d; record = [src, dst, seq]
counts_local = torch.tensor([(rank + 1) * (d + 1) for d in range(world)], dtype=torch.int64)
allc = [torch.empty_like(counts_local) for _ in range(world)]
dist.all_gather(allc, counts_local)
counts = torch.stack(allc).numpy()
so, ss, ro, rs = exchange_plan(counts)
recs = []
for d in range(world):
    for s in range(int(ss[rank][d])):
        recs += [rank, d, s]
send = torch.tensor(recs, dtype=torch.int64)
recv = all_to_all_records(send, ss[rank], rs[rank], words).view(-1, words).numpy()
assert recv.shape[0] == int(rs[rank].sum())
assert (recv[:, 1] == rank).all()
for src in range(world):
    part = recv[int(ro[rank][src]): int(ro[rank][src]) + int(rs[rank][src])]
    assert (part[:, 0] == src).all() and (part[:, 2] == np.arange(len(part))).all()
lo, hi = shard_range(1001, rank, world)
tot = torch.tensor([hi - lo]); dist.all_reduce(tot); assert int(tot) == 1001
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


_GLOO_ONESHOT_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from vinum_b200.dist import partial_block_words, collect_partial_blocks
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
w, cap = 5, 8
for groups in ([3, 6], [0, 8], [9, 2], [0, 0]):          # third case: rank 0 overflows its block
    g = groups[rank]
    blk = torch.zeros(partial_block_words(cap, w), dtype=torch.int64)
    blk[0] = g
    n = min(g, cap)
    # record j of rank r = [r, j, r*100+j, 7, 7]
    for j in range(n):
        blk[1 + j * w: 1 + (j + 1) * w] = torch.tensor([rank, j, rank * 100 + j, 7, 7])
    allb = torch.empty(world * blk.numel(), dtype=torch.int64)
    dist.all_gather_into_tensor(allb, blk)
    heads = [int(x) for x in allb.view(world, -1)[:, 0].tolist()]
    assert heads == groups                                   # every rank sees every header
    res = collect_partial_blocks(allb, heads, cap, w)
    if max(groups) > cap:
        assert res is None                                   # unanimous fallback to the repartition path
        continue
    total, recs = res
    assert total == sum(groups) and recs.numel() == total * w
    recs = recs.view(-1, w)
    want = [[r, j, r * 100 + j, 7, 7] for r in range(world) for j in range(groups[r])]
    assert recs.tolist() == want
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_one_shot_partial_block_exchange_over_gloo_world2(tmp_path):
    """The low-cardinality exchange of DistributedAggregator.finish() (one all-gather of fixed-size
    `[n_groups | records]` blocks, unanimous overflow fallback) with 2 CPU ranks."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker_oneshot.py"
    script.write_text(_GLOO_ONESHOT_WORKER.format(root=str(ROOT), port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_partial_group_exchange_over_gloo_world2(tmp_path):
    """The all-to-all-v of partial-aggregate records with 2 CPU ranks."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=str(ROOT), port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


_GLOO_SAMPLE_SORT_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from vinum_b200.dist import order_codes, choose_splitters, split_counts, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
rng = np.random.default_rng(5)
n = 40_001
for kind, desc in (("f", False), ("f", True), ("i", True), ("ties", False)):
    if kind == "f":
        col = rng.normal(0, 1, n); col[::97] = np.nan; col[::31] = -0.0; col[::29] = 0.0
    elif kind == "i":
        col = rng.integers(-2**62, 2**62, n)
    else:
        col = rng.integers(0, 4, n).astype(np.int64)      # four values: every splitter sits inside a run of ties
    lo, hi = shard_range(n, rank, world)
    mine = torch.from_numpy(col[lo:hi].copy())
    codes = order_codes(mine, desc)
    perm = torch.argsort(codes, stable=True)              # stands in for the device radix sort of the shard
    codes, keys, ids = codes[perm], mine[perm], perm + lo
    S = 64
    take = torch.linspace(0, codes.numel() - 1, S).long()
    allc = [torch.empty(S, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, codes[take].contiguous())
    splitters = choose_splitters(torch.stack(allc), world)
    sc = split_counts(codes, splitters)
    cm = [torch.empty(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cm, sc.contiguous())
    counts = torch.stack(cm)
    in_split, out_split = [int(x) for x in counts[rank]], [int(x) for x in counts[:, rank]]
    rc = torch.empty(sum(out_split), dtype=torch.int64); ri = torch.empty(sum(out_split), dtype=torch.int64)
    dist.all_to_all_single(rc, codes.contiguous(), out_split, in_split)
    dist.all_to_all_single(ri, ids.contiguous(), out_split, in_split)
    p2 = torch.argsort(rc, stable=True)                   # the owner's stable sort over runs in source-rank order
    out_ids = ri[p2]
    parts = [torch.empty(int(counts[:, r].sum()), dtype=torch.int64) for r in range(world)] if rank == 0 else None
    sizes = [int(counts[:, r].sum()) for r in range(world)]
    if rank == 0:
        parts[0] = out_ids
        for r in range(1, world):
            dist.recv(parts[r], src=r)
        got = torch.cat(parts).numpy()
        want = torch.argsort(order_codes(torch.from_numpy(col.copy()), desc), stable=True).numpy()
        assert np.array_equal(got, want), (kind, desc)     # == ONE stable sort of the whole column
    else:
        dist.send(out_ids, dst=0)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_sample_sort_plan_over_gloo_world2(tmp_path):
    """The splitter / all-to-all plan of dist.sample_sort_sharded with 2 CPU ranks (torch's stable
    argsort stands in for the device radix sort): the concatenation of the ranks' ranges must be
    exactly one stable sort of the whole column -- NaN last in both directions, both zeros equal,
    ties in global row order even when a splitter falls inside a run of equal keys."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker_ssort.py"
    script.write_text(_GLOO_SAMPLE_SORT_WORKER.format(root=str(ROOT), port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_every_option_has_a_default_and_a_name():
    import vinum_b200 as vb
    names = ("FILTER_STAGE FILTER_PF CMP_FAST ARITH_FAST ONEGROUP_FAST SORT_FUSE_LAST SORT_PREP "
             "AGG_LOG2S AGG_PF AGG_WARPS AGG_DIRECT AGG_DICT AGG_ENTRY AGG_HOT AGG_NOFAST AGG_WIDE AGG_PARTITION AGG_LEARN_LOG2 LIST_LOG2 DEBUG INGEST_STAGED "
             "INGEST_THREADS INGEST_PIECE_KB").split()
    for n in names:
        assert isinstance(vb.get_option(n), int)
    assert vb.get_option("VINUM_B200_SORT_PREP") == vb.get_option("SORT_PREP")   # the environment spelling resolves too
    with vb.options(AGG_DICT=0, SORT_PREP=0):
        assert vb.get_option("AGG_DICT") == 0 and vb.get_option("SORT_PREP") == 0
    assert vb.get_option("AGG_DICT") == 1 and vb.get_option("SORT_PREP") == 4


def test_string_rank_codes_are_order_preserving_and_keep_nulls():
    """vinum_b200.strings: the int32 code of a row is the rank of its string among the batch's distinct values
    (binary order, like std::string::operator< in StringMinMaxFunc, agg_funcs.h:219-261); NULL stays NULL."""
    import pyarrow as pa
    from vinum_b200.strings import key_matrix, rank_codes
    values = ["pear", None, "apple", "pear", "zz", "", "Apple", "éclair", "apple pie", None]
    codes, ordered = rank_codes(pa.array(values, type=pa.string()))
    assert ordered.to_pylist() == sorted({v for v in values if v is not None}, key=lambda s: s.encode())
    got = codes.to_pylist()
    for v, c in zip(values, got):
        assert (c is None) == (v is None)
        if v is not None:
            assert ordered[c].as_py() == v
    a, b = np.array(values, dtype=object), np.array(got, dtype=object)
    for i in range(len(values)):
        for j in range(len(values)):
            if values[i] is not None and values[j] is not None:
                assert (a[i].encode() < a[j].encode()) == (b[i] < b[j])
    codes, ordered = rank_codes(pa.array([None, None], type=pa.string()))
    assert codes.to_pylist() == [None, None] and len(ordered) == 0
    # group identity for the host merge: NULL == NULL, NaN payloads by their bits, -0.0 != +0.0 (as the device table)
    m = key_matrix([pa.array([1, None, 1, None]), pa.array([float("nan"), 0.0, float("nan"), -0.0])], 4)
    assert m.shape == (4, 4)
    assert (m[0] == m[2]).all() and not (m[1] == m[3]).all()
    assert key_matrix([], 3).shape == (3, 0)


def test_pinned_result_pool_keeps_the_large_blocks():
    """aggregate._pinned_give_back: over its budget the free list drops the SMALLEST block (pinning and unpinning a
    400 MB block costs ~0.1 s each, and a loop of large queries asks for the same size again)."""
    import vinum_b200.aggregate as A

    class Block:
        def __init__(self, mb):
            self.nbytes = mb << 20

    saved = list(A._PINNED_POOL)
    try:
        A._PINNED_POOL.clear()
        for mb in (50, 50, 50, 400, 400, 400, 400, 400, 400):
            A._pinned_give_back(Block(mb))
        sizes = sorted(b.nbytes >> 20 for b in A._PINNED_POOL)
        assert sum(sizes) << 20 <= A._PINNED_POOL_BYTES and len(sizes) <= A._PINNED_POOL_MAX
        assert sizes.count(400) >= 4 and sizes.count(50) <= 1
        got = A._pinned_take(300 << 20)
        assert got.nbytes == 400 << 20 and got not in A._PINNED_POOL
    finally:
        A._PINNED_POOL[:] = saved


def test_cardinality_estimate_behind_the_table_sizing():
    """vk_agg_estimate_groups: from (rows that passed the predicate, groups so far) to the final number of groups under
    even draws; the host grows the global table to 1.02 x this BEFORE the rows arrive."""
    import vinum_b200 as vb
    est = vb.lib.vk_agg_estimate_groups
    rng = np.random.default_rng(3)
    for true_groups in (1_000, 100_000, 1_000_000):
        for share in (0.1, 0.45, 1.0, 3.0):
            n = int(true_groups * share)
            seen = len(np.unique(rng.integers(0, true_groups, n)))
            got = est(float(n), float(seen))
            if share >= 0.45 or true_groups >= 100_000:     # enough collisions for the estimate to be tight
                assert abs(got - true_groups) / true_groups < 0.05, (true_groups, share, seen, got)
            else:
                assert got > seen
    assert est(1000.0, 1000.0) == 0.0        # every row its own group: no upper bound
    assert est(0.0, 0.0) == 0.0
    assert est(1e9, 1000.0) == pytest.approx(1000.0, rel=1e-6)   # saturated
