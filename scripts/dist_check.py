"""Multi-GPU parity check (run under torchrun, one rank per GPU): every rank aggregates its
own row range with the fused kernel, partial groups are repartitioned with one NCCL
all-to-all, rank 0 gathers and compares with the oracle over the whole row range.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py [ROWS_PER_RANK]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa, torch, torch.distributed as dist
import vinum_b200 as vb
from vinum_b200 import _lib as L, datagen, ops
from vinum_b200.aggregate import Aggregator
from vinum_b200.dist import DistributedAggregator
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_001
vb.lib.vk_set_device(lr)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
st = vb.default_stream()
ok = True
for keyname, keyt, card, mode in (("i0", pa.int64(), 1000, "finish"), ("i0", pa.int64(), 1000, "repartition"),
                                  ("i3", pa.int64(), 1_000_000, "finish")):
    t = datagen.device_table([keyname, "f0", "f1", "i1"], rank * n, n, stream=st)
    spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64()), (L.AGG_MIN, pa.int64()), (L.AGG_AVG, pa.int64())]
    d = DistributedAggregator(Aggregator([keyt], spec), st)
    d.update([t.column(keyname)], [None, t.column("f1"), t.column("i1"), t.column("i1")], ops.Predicate.compare(t.column("f0"), ">", 0.5))
    if mode == "finish":
        raw = d.finish()          # low cardinality: one all-gather; high: all-to-all repartition
    else:
        d.repartition()
        raw = d.gather_raw()
    if rank == 0:
        keys, kv, cnt, lo, hi, valid = raw
        host = {c: datagen.host_column(c, 0, n * world) for c in (keyname, "f0", "f1", "i1")}
        m = host["f0"] > 0.5
        k = host[keyname][m]
        order = np.argsort(keys[0].view(np.int64), kind="stable")
        uk, inv = np.unique(k, return_inverse=True)
        want_cnt = np.bincount(inv, minlength=len(uk))
        want_sum = np.bincount(inv, weights=host["f1"][m], minlength=len(uk))
        want_min = np.full(len(uk), np.iinfo(np.int64).max); np.minimum.at(want_min, inv, host["i1"][m])
        got_k = keys[0].view(np.int64)[order]
        good = (np.array_equal(got_k, uk) and np.array_equal(cnt[order].astype(np.int64), want_cnt)
                and np.allclose(lo[1].view(np.float64)[order], want_sum, rtol=1e-6, atol=1e-6))
        # MIN comes back as an order-preserving code only after finalize: check through result arrays on world==1 only
        print(f"dist_check key={keyname} mode={mode} world={world} rows={n * world} groups={len(uk)} exchange_bytes={d.exchange_bytes} ok={good}", flush=True)
        ok = ok and good
    d.agg.close()
# ---- C5 tail: ORDER BY over the gathered groups on rank 0 (device radix sort) ----
t = datagen.device_table(["i0", "f0", "f1"], rank * n, n, stream=st)
d = DistributedAggregator(Aggregator([pa.int64()], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())]), st)
d.update([t.column("i0")], [None, t.column("f1")], ops.Predicate.compare(t.column("f0"), ">", 0.5))
raw = d.finish()
if rank == 0:
    keys = vb.DeviceColumn.from_numpy(np.ascontiguousarray(raw[0][0].view(np.int64)), st)
    order = ops.sort_indices([keys], [L.ASC], st).to_numpy(st)
    good = np.array_equal(raw[0][0].view(np.int64)[order], np.arange(1000))
    print(f"dist_check C5 pipeline (filter->group-by->agg->order-by) world={world} ok={good}", flush=True)
    ok = ok and good
d.agg.close()

# ---- sharded ORDER BY: per-GPU radix sort + host k-way merge ----
from vinum_b200.dist import sort_sharded
for colname, desc in (("f3", True), ("i3", False)):
    col = datagen.device_column(colname, rank * n, n, stream=st)
    res = sort_sharded(col, rank * n, desc, st)
    if rank == 0:
        k, ids = res
        host = datagen.host_column(colname, 0, n * world)
        want = np.argsort(-host if desc else host, kind="stable")
        good = np.array_equal(ids, want) and np.array_equal(k, host[want])
        print(f"dist_check sharded sort {colname} desc={desc} world={world} rows={n * world} ok={good}", flush=True)
        ok = ok and good

flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
