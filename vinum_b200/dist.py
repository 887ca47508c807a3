"""Row-range data parallelism across the GPUs of one box.

The reference has no parallel execution at all (vinum/executor/executor.py:24-31); the
only step of the hot path that needs communication is the group-by: every rank
aggregates its own row range locally (fused filter -> hash aggregate), then the PARTIAL
groups are repartitioned by hash(key) mod world_size with ONE all-to-all over
NVLink/NVSwitch (NCCL through torch.distributed; gloo in the CPU tests), merged into the
owner's table and gathered on rank 0.  Filter, projection and local sort need no
collective; sort results merge on the host.

One process per GPU (`torchrun`); torch is used for the process group and the exchange
buffers only.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import _lib as L
from ._lib import lib


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) owned by `rank` (shard g of G owns
    [g*N/G, (g+1)*N/G))."""
    lo = (n_rows * rank) // world
    hi = (n_rows * (rank + 1)) // world
    return lo, hi


def exchange_plan(counts: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Given the world x world matrix counts[src][dst] of records, the per-rank send
    offsets/sizes and receive offsets/sizes (in records) of the all-to-all-v."""
    counts = np.asarray(counts, dtype=np.int64)
    send_sizes = counts
    send_offsets = np.concatenate([np.zeros((counts.shape[0], 1), dtype=np.int64),
                                   np.cumsum(counts, axis=1)[:, :-1]], axis=1)
    recv_sizes = counts.T.copy()
    recv_offsets = np.concatenate([np.zeros((counts.shape[0], 1), dtype=np.int64),
                                   np.cumsum(recv_sizes, axis=1)[:, :-1]], axis=1)
    return send_offsets, send_sizes, recv_offsets, recv_sizes


def partial_block_words(cap_groups: int, record_words: int) -> int:
    """Words of one rank's block in the one-shot exchange: [n_groups | cap_groups records]."""
    return 1 + cap_groups * record_words


def collect_partial_blocks(all_blocks: "torch.Tensor", heads: Sequence[int], cap_groups: int, record_words: int):
    """After the all-gather of every rank's `[n_groups | records...]` block: None when any rank
    overflowed its block (every rank sees every header, so the fallback to the all-to-all
    repartition is unanimous), else (total records, the ranks' valid records as one tensor)."""
    import torch
    if max(heads) > cap_groups:
        return None
    world = len(heads)
    blocks = all_blocks.view(world, partial_block_words(cap_groups, record_words))
    parts = [blocks[r, 1:1 + int(heads[r]) * record_words] for r in range(world) if heads[r]]
    total = int(sum(int(h) for h in heads))
    return total, (torch.cat(parts) if parts else blocks.new_empty(0))


def all_to_all_records(send: "torch.Tensor", send_sizes: Sequence[int], recv_sizes: Sequence[int], words: int,
                       group=None) -> "torch.Tensor":
    """all-to-all-v of fixed-width u64 records (int64 tensor view): `send` holds the
    records for rank 0, then rank 1, ... contiguously."""
    import torch
    import torch.distributed as dist
    total = int(sum(recv_sizes))
    recv = torch.empty(max(total, 1) * words, dtype=torch.int64, device=send.device)
    in_split = [int(s) * words for s in send_sizes]
    out_split = [int(s) * words for s in recv_sizes]
    dist.all_to_all_single(recv[: total * words], send[: sum(in_split)], out_split, in_split, group=group)
    return recv[: total * words]


PEER_OK, PEER_OVERFLOW, PEER_NEED_GROW, PEER_TIMEOUT = 1, 2, 3, 4
_MAX_RECORD_WORDS = 8 + 2 + 3 * 16   # VK_AGG_MAX_KEYS + nullmask + count + 3 words per function (16 functions)


class PeerWindow:
    """This rank's exchange window (include/vinum_b200.h "peer exchange") with every peer's window
    mapped through cudaIpc.  One per process group, created on first use (the handle exchange is the
    only collective it ever runs); `next_epoch()` numbers the queries, identically on every rank."""

    def __init__(self, group=None, cap_groups: int = 4096):
        import torch
        import torch.distributed as dist
        from .device import PinnedBuffer
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.cap = cap_groups
        h = C.c_void_p()
        lib.vk_peer_create(C.byref(h), self.rank, self.world, cap_groups, _MAX_RECORD_WORDS)
        self._h = h
        mine = C.create_string_buffer(64)
        lib.vk_peer_handle(h, mine)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine.raw), group=group)
        for r, raw in enumerate(handles):
            if r != self.rank:
                lib.vk_peer_open(h, r, C.create_string_buffer(raw, 64))
        self.epoch = 0
        self._pinned = PinnedBuffer(64)
        self.host_word = self._pinned.as_numpy(np.uint64, 8)
        torch.cuda.synchronize()
        dist.barrier(group=group)   # every window is mapped before anybody writes into one

    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    def decision_ptr(self, epoch: int) -> int:
        return int(L._lib.vk_peer_decision_ptr(self._h, epoch))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            L._lib.vk_peer_destroy(self._h)
            self._h = C.c_void_p()


_WINDOWS = {}


def peer_window(group=None) -> "PeerWindow":
    key = id(group) if group is not None else 0
    w = _WINDOWS.get(key)
    if w is None:
        w = _WINDOWS[key] = PeerWindow(group)
    return w


def close_peer_windows() -> None:
    """Unmap and free the exchange windows (before `destroy_process_group`)."""
    for w in _WINDOWS.values():
        w.close()
    _WINDOWS.clear()


def exchange_mode() -> str:
    """`peer` (default on CUDA + NCCL: no collective on the query path), `allgather` (round 1: one
    all-gather of fixed-size blocks + merge on rank 0) or `repartition` (always the all-to-all)."""
    import os
    import torch.distributed as dist
    mode = os.environ.get("VINUM_B200_DIST_MODE", "peer")
    if mode == "peer" and dist.get_backend() != "nccl":
        mode = "allgather"
    return mode


class DistributedAggregator:
    """Local Aggregator + partial-group repartition.  `gather_raw()` returns, on rank 0, the
    raw finalised groups of the whole job (same tuple as Aggregator.result_raw); None on
    the other ranks.

    Everything between the local aggregate and the final read-back is ordered on ONE CUDA
    stream (torch is pointed at the aggregate's stream through `torch.cuda.ExternalStream`),
    so the only host synchronisations are the two places where the host needs a number:
    the count matrix (all-to-all split sizes) and the number of groups per rank."""

    def __init__(self, aggregator, stream, group=None):
        import torch
        import torch.distributed as dist
        self.agg = aggregator
        self.stream = stream
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.exchange_bytes = 0
        self._recv_max = 0
        self._tstream = torch.cuda.ExternalStream(int(stream.handle)) if stream.handle else torch.cuda.default_stream()

    def update(self, keys, values, pred=None) -> None:
        self.agg.update(keys, values, pred, self.stream)

    def _mark(self, label: str) -> None:
        """Phase trace (VINUM_B200_DIST_TRACE=1): host wall clock between marks, stream drained."""
        import os
        import time
        if not os.environ.get("VINUM_B200_DIST_TRACE"):
            return
        self.stream.sync()
        now = time.perf_counter()
        if getattr(self, "_t_last", None) is not None and self.rank == 0:
            print(f"[dist] {label}: {(now - self._t_last) * 1e3:.3f} ms", flush=True)
        self._t_last = now

    def repartition(self) -> None:
        """hash(key) mod world all-to-all of the partial groups, merged into the owner."""
        import torch
        import torch.distributed as dist
        from .aggregate import Aggregator
        if self.world == 1:
            return
        self._mark("local aggregate")
        dev = torch.device("cuda", torch.cuda.current_device())
        words = C.c_int()
        lib.vk_agg_record_words(self.agg._h, C.byref(words))
        w = words.value
        st = self.stream
        with torch.cuda.stream(self._tstream):
            counts_dev = torch.empty(self.world, dtype=torch.int64, device=dev)
            lib.vk_agg_partition_counts(self.agg._h, self.world, C.c_void_p(counts_dev.data_ptr()), st.ptr)
            all_counts = torch.empty(self.world * self.world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(all_counts, counts_dev, group=self.group)
            counts = all_counts.view(self.world, self.world).cpu().numpy()   # host sync 1: split sizes
            self._mark("partition counts + all_gather + D2H")
            send_off, send_sz, _recv_off, recv_sz = exchange_plan(counts)
            self._recv_max = int(recv_sz.sum(axis=1).max())
            my_total = int(send_sz[self.rank].sum())
            send = torch.empty(max(my_total, 1) * w, dtype=torch.int64, device=dev)
            offs = torch.from_numpy(np.ascontiguousarray(send_off[self.rank])).to(dev, non_blocking=True)
            lib.vk_agg_export_partials(self.agg._h, self.world, C.c_void_p(offs.data_ptr()),
                                       C.c_void_p(send.data_ptr()), st.ptr)
            self._mark("export partials")
            recv = all_to_all_records(send, send_sz[self.rank], recv_sz[self.rank], w, self.group)
            self._mark("all_to_all")
            self.exchange_bytes = int(send_sz[self.rank].sum() - send_sz[self.rank][self.rank]) * w * 8
            # the owner's table = merge of everything it received (its own share included)
            n_recv = int(recv_sz[self.rank].sum())
            fresh = Aggregator(self.agg.key_types, self.agg.funcs, expected_groups=max(n_recv, 1))
            lib.vk_agg_merge_partials(fresh._h, C.c_void_p(recv.data_ptr()), n_recv, st.ptr)
            # `send` / `recv` / `offs` are torch allocations used by kernels on this same stream
            recv.record_stream(self._tstream)
        self._mark("merge partials")
        self.agg.close()   # drains the stream: the exchange buffers may be released after this
        self.agg = fresh
        self._mark("close old aggregate")

    # Partial groups per rank up to which the exchange is ONE all-gather: at low cardinality the
    # all-to-all's messages are a few KB and the cost is pure collective latency (measured:
    # ~0.1 ms per collective + host round trip, 4 of them per query), so every rank publishes
    # its whole partial table in one fixed-size block and rank 0 merges.
    SMALL_GROUPS = 4096

    def finish(self):
        """Whole-job result on rank 0 (None elsewhere).  Low cardinality: one all-gather of
        fixed-size blocks `[n_groups | records...]`, merged on rank 0.  Otherwise (any rank
        holds more than SMALL_GROUPS partial groups -- every rank sees every header, so the
        decision is unanimous): hash(key) mod world all-to-all repartition + gather."""
        import torch
        import torch.distributed as dist
        from .aggregate import Aggregator
        import os
        if self.world == 1:
            return self.agg.result_raw(self.stream)
        mode = exchange_mode()
        if mode == "repartition":   # A/B switch for tuning runs
            self.repartition()
            return self.gather_raw()
        if mode == "peer" and len(self.agg.key_vk) > 0:
            return self._finish_peer()
        self._mark("local aggregate")
        st = self.stream
        dev = torch.device("cuda", torch.cuda.current_device())
        words = C.c_int()
        lib.vk_agg_record_words(self.agg._h, C.byref(words))
        w = words.value
        cap = self.SMALL_GROUPS
        g = self.agg.num_groups(st)                                   # host sync 1
        blk_words = partial_block_words(cap, w)
        with torch.cuda.stream(self._tstream):
            blk = torch.empty(blk_words, dtype=torch.int64, device=dev)
            blk[0] = g
            if 0 < g <= cap:
                if getattr(self, "_zero_off", None) is None:
                    self._zero_off = torch.zeros(1, dtype=torch.int64, device=dev)
                lib.vk_agg_export_partials(self.agg._h, 1, C.c_void_p(self._zero_off.data_ptr()),
                                           C.c_void_p(blk.data_ptr() + 8), st.ptr)
            allb = torch.empty(self.world * blk_words, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allb, blk, group=self.group)
            heads = [int(x) for x in allb.view(self.world, blk_words)[:, 0].cpu().tolist()]      # host sync 2
            self._mark("export + all_gather of partial blocks")
            collected = collect_partial_blocks(allb, heads, cap, w)
            small = collected is not None
            if small and self.rank == 0:
                total, recs = collected
                fresh = Aggregator(self.agg.key_types, self.agg.funcs, expected_groups=max(total, 1))
                if total:
                    lib.vk_agg_merge_partials(fresh._h, C.c_void_p(recs.data_ptr()), total, st.ptr)
        if not small:
            self.repartition()
            return self.gather_raw()
        if self.rank != 0:
            return None
        raw = fresh.result_raw(st)
        self._mark("merge + finalize on rank 0")
        self.agg.close()
        self.agg = fresh
        return raw

    def _finish_peer(self):
        """The low-cardinality exchange without a collective: every rank's kernel stores its partial
        groups into rank 0's window over NVLink and releases a flag; rank 0's kernels acquire the
        flags, merge into rank 0's own table and acknowledge; ONE host synchronisation per rank (the
        final read-back on rank 0, the decision word elsewhere)."""
        self._mark("local aggregate")
        st = self.stream
        win = peer_window(self.group)
        epoch = win.next_epoch()
        dptr = win.decision_ptr(epoch)
        haddr = win._pinned.ptr
        raw = None
        if self.rank == 0:
            lib.vk_agg_peer_merge(self.agg._h, win._h, epoch, st.ptr)
            raw = self.agg.result_raw(st, extra_d2h=(haddr, dptr, 8))
        else:
            lib.vk_agg_peer_send(self.agg._h, win._h, 0, epoch, st.ptr)
            lib.vk_memcpy_d2h(C.c_void_p(haddr), C.c_void_p(dptr), 8, st.ptr)
            st.sync()
        decision = int(win.host_word[0])
        self._mark("peer exchange + merge + read-back")
        self.exchange_mode_used = "peer"
        if decision == PEER_OK:
            return raw
        if decision in (PEER_OVERFLOW, PEER_NEED_GROW):
            # unanimous (rank 0 wrote the same word into every window): nothing was merged
            self.exchange_mode_used = "peer->repartition"
            self.repartition()
            return self.gather_raw()
        raise RuntimeError(f"vinum_b200.dist: peer exchange failed on rank {self.rank} (decision {decision}: "
                           "a peer did not answer within 4 s)")

    def gather_raw(self):
        """Concatenate every rank's finalised groups on rank 0 (after repartition each
        group lives on exactly one rank): finalise into a padded device block, one
        `gather` of u64 words (the block's last column carries the group count)."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return self.agg.result_raw(self.stream)
        st = self.stream
        dev = torch.device("cuda", torch.cuda.current_device())
        nk, nf = len(self.agg.key_vk), len(self.agg.funcs)
        g = self.agg.num_groups(st)                               # host sync 2
        self._mark("num_groups")
        if self._recv_max == 0:
            # no repartition happened: agree on the padded width the slow way
            with torch.cuda.stream(self._tstream):
                m = torch.tensor([g], dtype=torch.int64, device=dev)
                dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
                self._recv_max = int(m.item())
        gmax = max(self._recv_max, 1)   # identical on every rank; g <= records received <= _recv_max
        n64 = nk + 1 + 2 * nf          # keys | count | lo | hi
        n8 = nk + nf                   # key_valid | valid (one byte per group, widened on the wire)
        with torch.cuda.stream(self._tstream):
            blk = torch.zeros((n64 + n8, gmax + 1), dtype=torch.int64, device=dev)
            b8 = torch.zeros((max(n8, 1), gmax), dtype=torch.uint8, device=dev)
            base, row = blk.data_ptr(), (gmax + 1) * 8
            p64 = lambda i: base + i * row
            p8 = lambda i: b8.data_ptr() + i * gmax
            okeys = (C.c_void_p * max(nk, 1))(*[p64(i) for i in range(nk)])
            okv = (C.c_void_p * max(nk, 1))(*[p8(i) for i in range(nk)])
            olo = (C.c_void_p * max(nf, 1))(*[p64(nk + 1 + f) for f in range(nf)])
            ohi = (C.c_void_p * max(nf, 1))(*[p64(nk + 1 + nf + f) for f in range(nf)])
            oval = (C.c_void_p * max(nf, 1))(*[p8(nk + f) for f in range(nf)])
            lib.vk_agg_result(self.agg._h, g, okeys, okv, C.c_void_p(p64(nk)), olo, ohi, oval, st.ptr)
            if n8:
                blk[n64:, :gmax] = b8[:n8]
            blk[0, gmax] = g
            self._mark("finalize into padded block")
            gathered = [torch.empty_like(blk) for _ in range(self.world)] if self.rank == 0 else None
            dist.gather(blk, gathered, dst=0, group=self.group)
            self._mark("gather")
            if self.rank != 0:
                return None
            allb = torch.stack(gathered).cpu().numpy().view(np.uint64)   # host sync 3 (rank 0)
        parts = []
        for r in range(self.world):
            gr = int(allb[r, 0, gmax])
            parts.append(allb[r, :, :gr])
        allp = np.concatenate(parts, axis=1)
        return (allp[:nk], allp[n64:n64 + nk].astype(bool), allp[nk], allp[nk + 1:nk + 1 + nf],
                allp[nk + 1 + nf:nk + 1 + 2 * nf], allp[n64 + nk:].astype(bool))


# ------------------------------------------------------------------------- sort ----
def merge_sorted_runs(keys: Sequence[np.ndarray], row_ids: Sequence[np.ndarray], descending: bool = False):
    """Host k-way merge of per-shard sorted runs ("per-GPU radix sort -> D2H of sorted
    (key, global row id) runs -> host k-way merge").  Run g holds shard g's rows in sorted order;
    shards are contiguous row ranges in rank order, so breaking ties by (run, position) keeps
    the GLOBAL sort stable, exactly like a single stable sort of the whole column.  NaN sorts
    after every number in both directions (Arrow SortIndices semantics, sort.cpp:33).

    NumPy's stable sort is a run-detecting merge sort: on a concatenation of k sorted runs it
    does the k-way merge in O(n log k)."""
    if not keys:
        return np.empty(0), np.empty(0, dtype=np.int64)
    k = np.concatenate(keys)
    ids = np.concatenate(row_ids)
    if k.dtype.kind == "f":
        nan = np.isnan(k)
        rank_key = np.where(nan, np.inf, -k if descending else k)
        # NaN after +inf: order by (is_nan, key); lexsort is stable
        order = np.lexsort((rank_key, nan))
    else:
        if descending:
            rank_key = ~k if k.dtype.kind == "u" else -1 - k   # order-reversing, overflow-free
        else:
            rank_key = k
        order = np.argsort(rank_key, kind="stable")
    return k[order], ids[order]


def _key_bits64(keys: np.ndarray) -> np.ndarray:
    """Sort keys as int64 words for the wire, losslessly: 8-byte types are reinterpreted, narrower
    floats widen to float64 first (exact, NaN / inf included), narrower integers sign / zero extend."""
    if keys.dtype.itemsize == 8:
        return keys.view(np.int64)
    if keys.dtype.kind == "f":
        return keys.astype(np.float64).view(np.int64)
    return keys.astype(np.int64)


def _key_from_bits64(words: np.ndarray, dtype) -> np.ndarray:
    dtype = np.dtype(dtype)
    if dtype.itemsize == 8:
        return words.view(dtype)
    if dtype.kind == "f":
        return words.view(np.float64).astype(dtype)
    return words.astype(dtype)


def sort_sharded(column, row0: int, descending: bool, stream, group=None):
    """ORDER BY one numeric column over row-range shards: every rank sorts its shard on the
    device (stable LSD radix sort), rank 0 gathers the sorted (key, global row id) runs and
    merges them on the host.  Returns (sorted keys, global row ids) on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    from . import ops
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    idx = ops.sort_indices([column], [L.DESC if descending else L.ASC], stream)
    keys_sorted = ops.take(column, idx, stream).to_numpy(stream)
    ids = idx.to_numpy(stream) + np.int64(row0)
    if world == 1:
        return keys_sorted, ids
    dev = torch.device("cuda", torch.cuda.current_device())
    n = torch.tensor([len(ids)], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    nmax = max(max(sizes), 1)
    pack = np.zeros((2, nmax), dtype=np.int64)
    pack[0, :len(ids)] = _key_bits64(keys_sorted)
    pack[1, :len(ids)] = ids
    t = torch.from_numpy(pack).to(dev)
    gathered = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0, group=group)
    if rank != 0:
        return None
    runs_k, runs_i = [], []
    for g, s in zip(gathered, sizes):
        a = g.cpu().numpy()
        runs_k.append(_key_from_bits64(a[0, :s], keys_sorted.dtype))
        runs_i.append(a[1, :s])
    return merge_sorted_runs(runs_k, runs_i, descending)


# ------------------------------------------------------------------ sample sort ----
def order_codes(values: "torch.Tensor", descending: bool) -> "torch.Tensor":
    """Signed int64 codes that order like Arrow's SortIndices orders the values (sort.cpp:22-48):
    ascending or descending numbers, -0.0 == +0.0, NaN after every number in BOTH directions.
    `values`: int64 or float64 tensor."""
    import torch
    if values.dtype == torch.float64:
        bits = values.view(torch.int64)
        nan = torch.isnan(values)
        bits = torch.where(values == 0, torch.zeros_like(bits), bits)                  # -0.0 -> +0.0
        code = bits ^ ((bits >> 63) & 0x7FFFFFFFFFFFFFFF)                               # monotonic in the value
        if descending:
            code = ~code
        return torch.where(nan, torch.full_like(code, 0x7FFFFFFFFFFFFFFF), code)
    if values.dtype != torch.int64:
        raise TypeError("order_codes: int64 or float64")
    return ~values if descending else values


def choose_splitters(all_samples: "torch.Tensor", world: int) -> "torch.Tensor":
    """world - 1 splitters at the equal quantiles of the gathered (sorted-per-rank) samples."""
    import torch
    s, _ = torch.sort(all_samples.flatten())
    n = s.numel()
    pos = [(n * (r + 1)) // world for r in range(world - 1)]
    return s[torch.tensor([min(max(p, 1), n) - 1 for p in pos], dtype=torch.long, device=s.device)] if n else s[:0]


def split_counts(sorted_codes: "torch.Tensor", splitters: "torch.Tensor") -> "torch.Tensor":
    """Rows of an ascending code run that go to each of len(splitters) + 1 destinations: destination r
    gets the codes in (splitter[r-1], splitter[r]] -- equal codes always travel together, whatever rank
    they come from, so the destination's stable sort keeps ties in global row order."""
    import torch
    cuts = torch.searchsorted(sorted_codes, splitters, right=True)
    edges = torch.cat([cuts.new_zeros(1), cuts, cuts.new_full((1,), sorted_codes.numel())])
    return edges[1:] - edges[:-1]


def sample_sort_sharded(column, row0: int, descending: bool, stream, group=None, samples_per_rank: int = 256):
    """ORDER BY one null-free int64 / float64 column over row-range shards, entirely on the GPUs
    (SURVEY 8f row 4): every rank radix-sorts its shard, the ranks agree on world - 1 splitters from a
    sample, ONE all-to-all (NCCL over NVLink) sends every row to the rank that owns its key range, and
    that rank's stable radix sort over the received runs -- concatenated in source-rank order, which is
    row order -- finishes its range.  Rank r returns (keys, global row ids) of the r-th range of the
    global order as NumPy arrays; the concatenation over ranks is what one stable sort of the whole
    column returns (Sort::Sorted, sort.cpp:15-63).  Replaces the rank-0 host merge of `sort_sharded`."""
    import torch
    import torch.distributed as dist
    from . import ops
    from .device import DeviceColumn
    world = dist.get_world_size(group)
    st = stream
    ts = torch.cuda.ExternalStream(int(st.handle)) if st.handle else torch.cuda.default_stream()
    order = L.DESC if descending else L.ASC
    if column.has_nulls or column.dtype not in (L.I64, L.F64):
        raise TypeError("sample_sort_sharded: null-free int64 / float64 keys (use sort_sharded otherwise)")
    n = column.length
    idx, keys_sorted = ops.sort_indices_keys([column], [order], st)
    if world == 1:
        return keys_sorted.to_numpy(st), idx.to_numpy(st) + np.int64(row0)
    dev = torch.device("cuda", torch.cuda.current_device())
    tdt = torch.float64 if column.dtype == L.F64 else torch.int64
    with torch.cuda.stream(ts):
        keys_t = _as_tensor(keys_sorted, n, tdt, dev)
        ids_t = _as_tensor(idx, n, torch.int64, dev) + row0
        codes = order_codes(keys_t, descending)
        # ---- splitters from an evenly spaced sample of every rank's sorted run ----
        take = torch.linspace(0, max(n - 1, 0), samples_per_rank, device=dev).long() if n else torch.zeros(0, dtype=torch.long, device=dev)
        mine = codes[take] if n else codes.new_full((samples_per_rank,), 0x7FFFFFFFFFFFFFFF)
        if mine.numel() < samples_per_rank:
            mine = torch.cat([mine, mine.new_full((samples_per_rank - mine.numel(),), 0x7FFFFFFFFFFFFFFF)])
        gathered = torch.empty(world * samples_per_rank, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, mine.contiguous(), group=group)
        splitters = choose_splitters(gathered, world)
        send_counts = split_counts(codes, splitters)
        counts = torch.empty(world * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, send_counts.contiguous(), group=group)
        counts = counts.view(world, world).cpu()                                   # host sync: split sizes
        rank = dist.get_rank(group)
        in_split = [int(x) for x in counts[rank]]
        out_split = [int(x) for x in counts[:, rank]]
        total = sum(out_split)
        recv_k = torch.empty(max(total, 1), dtype=tdt, device=dev)[:total]
        recv_i = torch.empty(max(total, 1), dtype=torch.int64, device=dev)[:total]
        dist.all_to_all_single(recv_k, keys_t.contiguous(), out_split, in_split, group=group)
        dist.all_to_all_single(recv_i, ids_t.contiguous(), out_split, in_split, group=group)
        ts.synchronize()
    if total == 0:
        return np.empty(0, dtype=np.float64 if tdt == torch.float64 else np.int64), np.empty(0, dtype=np.int64)
    # ---- the received runs, in source-rank order, through one more stable device sort ----
    rk = DeviceColumn.from_device_ptr(recv_k.data_ptr(), total, column.dtype, owner=recv_k)
    ri = DeviceColumn.from_device_ptr(recv_i.data_ptr(), total, L.I64, owner=recv_i)
    perm, out_keys = ops.sort_indices_keys([rk], [order], st)
    out_ids = ops.take(ri, perm, st)
    return out_keys.to_numpy(st), out_ids.to_numpy(st)


def _as_tensor(col, n: int, dtype, dev):
    """A torch view of a device column's values (no copy; the column must outlive the tensor's use)."""
    import torch
    if n == 0:
        return torch.empty(0, dtype=dtype, device=dev)
    iface = {"shape": (n,), "typestr": "<f8" if dtype == torch.float64 else "<i8",
             "data": (col.data_ptr + col.offset * 8, False), "version": 3, "strides": None}

    class _Holder:
        __cuda_array_interface__ = iface
        _keep = col
    return torch.as_tensor(_Holder(), device=dev)
