"""One logical query over a table that is row-range sharded across the GPUs of one box.

`north_star`: "batches shard by row-range across the 8 GPUs of one box with an NCCL all-to-all over
NVLink only for the hash-aggregate key repartition and a final host merge for sort".  The
reference has one entry point, `Table.sql` (vinum/api/table.py:266-274), and runs it on one thread
(vinum/executor/executor.py:24-31); here every rank of a `torchrun` job holds the rows
[g*N/G, (g+1)*N/G) of the logical table as its own host `pyarrow.Table`, calls the same
`ShardedTable.sql(query)` and rank 0 gets the answer:

* aggregate queries: each rank runs the fused filter -> hash aggregate over its shard (the same
  streaming path as `Table.sql`), the partial groups meet on rank 0 through
  `DistributedAggregator.finish()` (peer-memory exchange at low cardinality, hash(key) mod world
  all-to-all otherwise), and HAVING / ORDER BY / LIMIT / the final projection run on rank 0 over the
  merged groups;
* any other query: each rank runs the WHERE (and, when there is an ORDER BY ... LIMIT k, keeps only
  its own first offset + k rows), the surviving rows are gathered on rank 0 in rank order -- which
  is row order, so a stable sort of the concatenation is the stable sort of the whole table -- and
  ORDER BY / LIMIT / projection finish there.

The other ranks return an empty table with the result's schema.
"""
from __future__ import annotations

from typing import Optional

import pyarrow as pa

from .sql.ast import Column, Query, contains_aggregate, walk
from .sql.engine import Engine, _used_columns, execute_sql
from .sql.parser import parse_sql


def init(backend: str = "nccl") -> int:
    """Join the job `torchrun` started: bind this process to GPU `LOCAL_RANK`, create the process
    group.  Returns the rank."""
    import os
    import torch
    import torch.distributed as dist
    from ._lib import lib
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        lib.vk_set_device(local)
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return dist.get_rank()


def bind_to_gpu_numa(device_index: Optional[int] = None) -> Optional[list]:
    """Pin this process (and the ingest workers it starts later) to the CPU cores NVML reports as
    local to its GPU, so that pinned buffers and bounce copies stay on the GPU's socket.  Returns the
    core list, or None when NVML has no answer (single-socket boxes report every core for every GPU)."""
    import os
    try:
        import pynvml as nv
        nv.nvmlInit()
        idx = int(os.environ.get("LOCAL_RANK", "0")) if device_index is None else device_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis and all(x.strip().isdigit() for x in vis.split(",")):
            idx = int(vis.split(",")[idx])
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cores = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def _exchange(group):
    def run(agg, stream):
        from .dist import DistributedAggregator
        d = DistributedAggregator(agg, stream, group)
        raw = d.finish()
        run.mode = getattr(d, "exchange_mode_used", None)
        if raw is None:
            raw = d.agg.empty_raw()
        return d.agg, raw
    run.mode = None
    return run


class ShardedTable:
    """This rank's row range of a logical table; `sql()` is collective (every rank calls it)."""

    def __init__(self, local_table: pa.Table, group=None):
        self._table = local_table
        self._group = group
        self.last_stats: dict = {}

    @property
    def schema(self) -> pa.Schema:
        return self._table.schema

    def sql(self, query: str):
        from .table import Table
        import torch.distributed as dist
        stats: dict = {}
        q = parse_sql(query, self._table.schema.names)
        is_agg = q.distinct or q.has_group_clause or any(contains_aggregate(e) for e in q.select)
        if is_agg:
            ex = _exchange(self._group)
            out = execute_sql(query, self._table, stats=stats, exchange=ex)
            stats["exchange"] = ex.mode
        else:
            out = self._rows_query(q, query, stats)
        stats["world"] = dist.get_world_size(self._group)
        self.last_stats = stats
        return Table(out)

    def _rows_query(self, q: Query, text: str, stats: dict) -> pa.Table:
        """WHERE (+ a local ORDER BY ... LIMIT) on every rank, the rest on rank 0."""
        import torch.distributed as dist
        rank, world = dist.get_rank(self._group), dist.get_world_size(self._group)
        eng = Engine(self._table)
        used = _used_columns(eng._bind(q))
        need = None if q.limit is None else q.offset + q.limit
        local = Query(select=tuple(Column(n) for n in used) or q.select, where=q.where,
                      order_by=q.order_by if need is not None else (), sort_order=q.sort_order if need is not None else (),
                      limit=need, offset=0)
        part = Engine(self._table).execute(local) if used else self._table.slice(0, 0)
        parts = [None] * world if rank == 0 else None
        dist.gather_object(part, parts, dst=0, group=self._group)
        rest = Query(select=q.select, distinct=q.distinct, order_by=q.order_by, sort_order=q.sort_order,
                     limit=q.limit, offset=q.offset)
        if rank == 0:
            merged = pa.concat_tables(parts).combine_chunks() if used else self._table
            stats["gathered_rows"] = merged.num_rows
        else:
            merged = part.slice(0, 0)
        e2 = Engine(merged)
        out = e2.execute(rest)
        stats.update({k: v for k, v in e2.stats.items() if k != "kernels_before"})
        return out
