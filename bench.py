#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows R] [--impl ours|reference]

Workload (config.workload = "northstar_filter_hashagg_T8"): the table T8 of SURVEY 8d
-- 1e9 rows x 8 int64/float64 columns per GPU, generated on the device -- and the query
    SELECT i0, COUNT(*), SUM(f1) FROM t WHERE f0 > 0.5 GROUP BY i0
One STEP = one full execution of that query over the resident table (create the
aggregate state, fused filter->hash-aggregate over every row, finalise, read the 1000
groups back).  `value` = rows/s with the table already in HBM; the table (64 GB, 24 GB
touched per step) is far larger than L2, so no flush is needed between steps.

`e2e` runs the same query through the public API the reference's users call --
`Table.from_arrow(t).sql(QUERY)` -- on a host Arrow table in ordinary PAGEABLE memory (what
`pa.table(numpy arrays)`, `from_pandas` and `read_csv` produce): every step wraps the table, parses
and plans the SQL, moves the three referenced columns host->device through the library's pinned
bounce-buffer pool (double-buffered chunks) and reads the result back.  `e2e_pinned` is the same with
the caller's buffers already page-locked (plain DMA).

N > 1 (torchrun): each rank owns its own 1e9-row shard (weak scaling), aggregates it locally; the
partial groups meet on rank 0 through the peer-memory exchange (kernels store into rank 0's window
over NVLink; no collective on the step), rank 0 merges, finalises and orders by key -- BASELINE
config 5.  Its `e2e` is ONE logical query over the row-range-sharded host table:
`Table.from_arrow(shard).shard().sql(QUERY ORDER BY i0)` on every rank, the answer on rank 0.

Also in the line: `configs` -- one sub-record per BASELINE.json config and per kernel path that the
headline does not exercise (C2 filter, C3 group-by, C4 sort, the north-star query on non-dense keys =
CTA hash table instead of direct group ids, on 1e6 groups = global table, on 8.4e6 groups = partitioned plan), each with ms, rows/s
and its own roofline fraction; `verified` -- the timed result checked, outside the timed region,
against NumPy over the regenerated rows (all 1000 counts exact, sums to 1e-6).

--impl reference: the reference's STOCK code path -- `vinum.Table.from_arrow(t).sql(QUERY)`: its own
binder, QueryPlanner, RecursiveExecutor and compiled C++ operators (oracle/_ref; only the SQL parser is
this repo's stand-in, pglast cannot be installed) -- on ONE core, which is how the reference runs
(vinum/executor/executor.py:24-31).  `all_cores` in the same line is the labelled extra: one such process
per host core over disjoint row ranges.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QUERY = "SELECT i0, COUNT(*), SUM(f1) FROM t WHERE f0 > 0.5 GROUP BY i0"
ALG_BYTES_PER_ROW = 24  # f0 + i0 + f1, 8 B each (SURVEY 8d north-star row)
FUNCS = [("COUNT_STAR", "", "count_star"), ("SUM", "f1", "sum_f1")]


def _traffic_per_row():
    """DRAM bytes per row of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(p) as f:
            return float(json.load(f)["dram_bytes_per_row"])
    except Exception:
        return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons during the timed region.  NVML is read in-process
    (nvidia_ml_py) every 25 ms -- polling the nvidia-smi binary instead takes the driver lock for
    milliseconds at a time and shows up as step-time jitter; nvidia-smi is the fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device_index: int):
        self.idx = device_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.source = None

    def _nvml_loop(self, nv, h):
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.025)

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                self.sm.append(float(parts[0]))
                self.mx.append(float(parts[1]))
                for nm, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(0.25)

    def prepare(self):
        """NVML initialisation costs milliseconds: do it BEFORE the barrier that opens the timed
        region (rank 0 alone samples; a late start of rank 0 would be charged to every other rank)."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES-relative index -> NVML handle through the UUID-free common case
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.idx
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
        except Exception:
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)

    def start(self):
        if self.thread is None:
            self.prepare()
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler not started"]}
        self.stop_flag.set()
        self.thread.join(timeout=6)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "samples": len(self.sm), "source": self.source, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------- host checks ----
def _expect_chunk(args):
    row0, rows = args
    from vinum_b200.datagen import host_column
    k = host_column("i0", row0, rows)
    m = host_column("f0", row0, rows) > 0.5
    v = host_column("f1", row0, rows)
    k = k[m]
    return np.bincount(k, minlength=1000), np.bincount(k, weights=v[m], minlength=1000)


def host_expected(row0: int, rows: int, procs: int):
    """COUNT(*) and SUM(f1) per i0 over rows [row0, row0 + rows) with f0 > 0.5, by NumPy over the
    regenerated columns (the generator is a pure function of the row number).  Runs in a fork pool
    BEFORE this process touches CUDA."""
    import multiprocessing as mp
    chunk = 1 << 24
    jobs = [(r, min(chunk, row0 + rows - r)) for r in range(row0, row0 + rows, chunk)]
    cnt = np.zeros(1000, dtype=np.int64)
    sm = np.zeros(1000, dtype=np.float64)
    with mp.get_context("fork").Pool(max(1, procs)) as pool:
        for c, s in pool.imap_unordered(_expect_chunk, jobs):
            cnt += c
            sm += s
    return cnt, sm


def check_groups(keys, counts, sums, want_cnt, want_sum, what: str):
    keys = np.asarray(keys).astype(np.int64)
    order = np.argsort(keys, kind="stable")
    if not np.array_equal(keys[order], np.arange(1000)):
        raise AssertionError(f"{what}: group keys are not exactly 0..999")
    if not np.array_equal(np.asarray(counts).astype(np.int64)[order], want_cnt):
        raise AssertionError(f"{what}: COUNT(*) differs from the NumPy check")
    got = np.asarray(sums, dtype=np.float64)[order]
    if not np.allclose(got, want_sum, rtol=1e-6, atol=0):
        raise AssertionError(f"{what}: SUM(f1) differs from the NumPy check by more than 1e-6 relative")


# ---------------------------------------------------------------- reference arm ----
def _ref_stock_worker(args):
    """One process of the reference arm: regenerate a row range, run the reference's stock path."""
    row0, rows, repeat = args
    from oracle import ref_stack
    from vinum_b200.datagen import host_table  # pure-NumPy generator (no CUDA call)
    vn = ref_stack.reference_vinum()
    table = host_table(["i0", "f0", "f1"], row0, rows)
    tbl = vn.Table.from_arrow(table)
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        out = tbl.sql(QUERY)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out.to_arrow().num_rows if hasattr(out, "to_arrow") else len(out.to_pandas())


def run_reference(args) -> dict:
    from oracle import ref_stack
    if not ref_stack.available():
        return {"impl": "reference", "unavailable": "oracle/_ref (compiled reference operators + vinum_pyref.zip) is not built"}
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    rows = args.ref_rows
    # ---- the reference as it runs: one process, one thread ----
    times = []
    with ctx.Pool(1) as pool:
        for step in range(args.warmup + args.steps):
            dt, groups = pool.apply(_ref_stock_worker, ((0, rows, 1),))
            assert groups == 1000
            if step >= args.warmup:
                times.append(dt)
    t = statistics.median(times)
    value = rows / t
    # ---- labelled extra: one reference process per host core over disjoint row ranges ----
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, args.ref_procs if args.ref_procs > 0 else cores))
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_stock_worker, [(p * rows, rows, 2) for p in range(procs)])
    all_cores = {"value": rows * procs / max(r[0] for r in res), "unit": "rows/s", "processes": procs,
                 "note": "NOT the reference: one stock reference process per host core over disjoint row ranges, "
                         "slowest process, partial results not merged"}
    sample = (f"vinum.Table.from_arrow(first {rows} rows of the same generator).sql(QUERY): the reference's binder, "
              f"QueryPlanner, RecursiveExecutor (batch 10000) and compiled C++ operators (oracle/_ref), SQL text parsed by "
              f"this repo's stand-in parser; 1 process, 1 thread")
    return {
        "impl": "reference", "metric": "rows/sec filter->hash-agg", "value": value, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "northstar_filter_hashagg_T8", "query": QUERY, "rows_per_step": rows},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": sample},
        "all_cores": all_cores,
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def cpu_baseline_single(rows: int) -> dict:
    """The reference's stock path on ONE core (the reference is single-threaded by design,
    vinum/executor/executor.py:24-31) over a bounded sample of the same workload, timed in a child
    process so that this process's CUDA context and threads do not disturb it."""
    from oracle import ref_stack
    if not ref_stack.available():
        return {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
    code = ("import sys, json; sys.path.insert(0, %r); import bench; "
            "print(json.dumps(bench._ref_stock_worker((0, %d, 2))))" % (ROOT, rows))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        return {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": "failed: " + r.stderr[-300:]}
    best, groups = json.loads(r.stdout.strip().splitlines()[-1])
    assert groups == 1000
    return {"value": rows / best, "unit": "rows/s", "cores": 1, "kind": "reference",
            "sample": f"first {rows} rows of the same generator through vinum.Table.sql (reference planner + executor + "
                      f"compiled C++ operators, batch 10000), best of 2, one process / one thread"}


def _dropin_worker(rows: int):
    """The reference's own `vinum.Table.sql` (its binder, planner, executor, Python operators) on top of
    vinum_b200: `vinum_lib` = vinum_b200.vinum_lib and AggregateOperator dispatching scan -> filter ->
    aggregate plans to the fused device path (vinum_b200.compat.install)."""
    from oracle import ref_stack
    import pyarrow as pa
    from vinum_b200 import datagen
    vn = ref_stack.reference_vinum(gpu_operators=True)
    table = pa.table({n: pa.array(datagen.host_column(n, 0, rows)) for n in ("i0", "f0", "f1")})
    tbl = vn.Table.from_arrow(table)
    best, groups = None, 0
    for _ in range(4):
        t0 = time.perf_counter()
        out = tbl.sql(QUERY).to_arrow()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        groups = out.num_rows
    return best, groups


def dropin_reference_api(rows: int) -> dict:
    """Rows/s of the drop-in under the REFERENCE's Python layers (separate process: the reference binds
    `vinum_lib` at import time).  Reported beside cpu_baseline, which is the same call on the reference's
    own operators."""
    from oracle import ref_stack
    if not ref_stack.available():
        return {"value": None, "note": "oracle/_ref (vinum_pyref.zip) not built"}
    code = ("import sys, json; sys.path.insert(0, %r); import bench; "
            "print(json.dumps(bench._dropin_worker(%d)))" % (ROOT, rows))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        return {"value": None, "note": "failed: " + r.stderr[-400:]}
    best, groups = json.loads(r.stdout.strip().splitlines()[-1])
    assert groups == 1000
    return {"value": rows / best, "unit": "rows/s", "rows": rows, "ms": best * 1e3,
            "api": "reference's vinum.Table.from_arrow(pageable table).sql(QUERY) after vinum_b200.compat.install()",
            "note": "best of 4; the reference's parser stand-in, binder, planner and executor run unchanged; "
                    "AggregateOperator.next dispatches the scan -> filter -> aggregate plan to the fused device path"}


# --------------------------------------------------------------------- our arm ----
class _Timer:
    """CUDA-event timing on the package's stream (torch.cuda.Event only sees torch's stream)."""

    def __init__(self, lib, st):
        self.lib, self.st = lib, st
        self.e0, self.e1 = C.c_void_p(), C.c_void_p()
        lib.vk_event_create(C.byref(self.e0))
        lib.vk_event_create(C.byref(self.e1))

    def run(self, fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        self.st.sync()
        out = []
        for _ in range(reps):
            self.lib.vk_event_record(self.e0, self.st.ptr)
            fn()
            self.lib.vk_event_record(self.e1, self.st.ptr)
            self.lib.vk_event_sync(self.e1)
            ms = C.c_float()
            self.lib.vk_event_elapsed_ms(self.e0, self.e1, C.byref(ms))
            out.append(float(ms.value))
        return statistics.median(out), min(out)


def run_configs(vb, t8, rows, st, peak) -> dict:
    """One sub-record per BASELINE.json config / kernel path the headline does not exercise.  Median of 5
    after 2 warm-ups, CUDA events, inputs resident and larger than L2.  `frac` = algorithmic bytes
     / time / measured HBM peak."""
    from vinum_b200 import _lib as L, datagen, ops
    from vinum_b200.aggregate import Aggregator
    import pyarrow as pa
    lib = vb.lib
    tm = _Timer(lib, st)
    out = {}

    def rec(name, ms, best, n, bytes_per_row, launches, **extra):
        gbs = bytes_per_row * n / (ms / 1e3) / 1e9
        out[name] = dict(ms=round(ms, 4), ms_best=round(best, 4), rows=n, rows_per_s=n / (ms / 1e3),
                         algorithmic_bytes_per_row=bytes_per_row, achieved_gbs=round(gbs, 1), frac=round(gbs / peak, 4),
                         launches=launches, **extra)

    def launches_of(fn):
        a = lib.vk_launch_count()
        fn()
        st.sync()
        return int(lib.vk_launch_count() - a)

    spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())]

    def agg_query(key, key_t, pred_col, check=None):
        state = {}

        def fn():
            agg = Aggregator([key_t], spec)
            pred = ops.Predicate.compare(pred_col, ">", 0.5) if pred_col is not None else None
            agg.update([key], [None, t8.column("f1")], pred, st)
            state["raw"] = agg.result_raw(st)
            state["path"] = agg.last_path
            agg.close()
        return fn, state

    # ---- C3: SELECT k32, COUNT(*), SUM(f1) FROM t GROUP BY k32 (1e9 rows, int32 key, no predicate) ----
    k32 = datagen.device_column("k32", t8.row0 if hasattr(t8, "row0") else 0, rows, stream=st)
    fn, state = agg_query(k32, pa.int32(), None)
    ms, best = tm.run(fn)
    rec("C3_groupby_int32", ms, best, rows, 12, launches_of(fn), groups=int(len(state["raw"][2])), agg_path=state["path"],
        query="SELECT k32, COUNT(*), SUM(f1) FROM t GROUP BY k32")
    assert len(state["raw"][2]) == 1000 and int(state["raw"][2].sum()) == rows
    del k32
    # ---- north-star on non-dense keys: CTA hash table instead of direct group ids ----
    hk = ops.arith("*", t8.column("i0"), 2654435761, st)
    fn, state = agg_query(hk, pa.int64(), t8.column("f0"))
    ms, best = tm.run(fn)
    rec("northstar_hash_keys", ms, best, rows, 24, launches_of(fn), groups=int(len(state["raw"][2])), agg_path=state["path"],
        query="same query, key = i0 * 2654435761 (no dense range: shared-memory hash table)")
    assert len(state["raw"][2]) == 1000
    del hk
    # ---- north-star with 1e6 groups: more than the shared-memory tables hold ----
    fn, state = agg_query(t8.column("i3"), pa.int64(), t8.column("f0"))
    ms, best = tm.run(fn, reps=3, warm=1)
    rec("northstar_1e6_groups", ms, best, rows, 24, launches_of(fn), groups=int(len(state["raw"][2])), agg_path=state["path"],
        query="same query, GROUP BY i3 (1e6 groups)")
    assert len(state["raw"][2]) == 1_000_000
    state.clear()
    # ---- north-star with 8.4e6 groups: the table (805 MB) is beyond the L2 -> partitioned plan ----
    k23 = ops.arith("&", t8.column("i1"), (1 << 23) - 1, st)
    fn, state = agg_query(k23, pa.int64(), t8.column("f0"))
    ms, best = tm.run(fn, reps=3, warm=1)
    rec("northstar_8e6_groups", ms, best, rows, 24, launches_of(fn), groups=int(len(state["raw"][2])), agg_path=state["path"],
        query="same query, GROUP BY (i1 & 8388607) (8.4e6 groups; read-back of 200 MB of groups included)")
    if rows >= 200_000_000:
        assert len(state["raw"][2]) == 1 << 23
    state.clear()
    del k23
    # ---- C2: SELECT * FROM t WHERE f0 > 0.5 over {i1, i2, f0, f1}, 1e8 rows ----
    n2 = min(rows, 100_000_000)
    names = ["i1", "i2", "f0", "f1"]
    cols = [t8.column(c).slice(0, n2) for c in names]
    vp = ops.Predicate.compare(cols[2], ">", 0.5).vk()
    vcols = (L.VkColumn * 4)(*[c.vk() for c in cols])
    outs = [vb.DeviceColumn.empty(n2, c.dtype, stream=st) for c in cols]
    od = (C.c_void_p * 4)(*[o.data_ptr for o in outs])
    ov = (C.c_void_p * 4)()
    scratch = vb.DeviceBuffer(lib.vk_filter_scratch_bytes(n2), st)
    nout = vb.DeviceBuffer(8, st)
    fn = lambda: lib.vk_filter(C.byref(vp), n2, vcols, 4, od, ov, C.c_void_p(nout.ptr), C.c_void_p(scratch.ptr), st.ptr)
    ms, best = tm.run(fn)
    sel = int(nout.to_numpy(np.int64, 1, st)[0])
    rec("C2_filter_project", ms, best, n2, 32 + 32 * sel / n2, launches_of(fn), rows_out=sel,
        query="SELECT * FROM t WHERE f0 > 0.5 (i1, i2, f0, f1)")
    del outs, scratch
    # ---- C4: SELECT f3 FROM t ORDER BY f3 DESC, 1e8 rows (permutation + the sorted column) ----
    f3 = t8.column("f3").slice(0, n2)
    state = {}

    def sort_fn():
        state["idx"], state["sorted"] = ops.sort_indices_keys([f3], [L.DESC], st)
    ms, best = tm.run(sort_fn, reps=3, warm=1)
    # 8 B read + 12 B written by prepare, 24 B per pass x 8 passes (the last writes 8 + 8 instead of 12),
    # SURVEY 8d's accounting with P = 8 digits of 8 bits
    rec("C4_sort_f64_desc", ms, best, n2, 8 + 12 + 24 * 7 + 12 + 16, launches_of(sort_fn), passes=8,
        query="SELECT f3 FROM t ORDER BY f3 DESC (int64 permutation + sorted f3)")
    srt = state["sorted"]
    ok = ops.compare(srt.slice(0, n2 - 1), ">=", srt.slice(1, n2 - 1), st)
    chk = Aggregator([], [(L.AGG_COUNT_STAR, None)])
    chk.update_count_rows(n2 - 1, ops.Predicate.from_mask(ok), st)
    assert int(chk.result_raw(st)[2][0]) == n2 - 1, "C4 output is not sorted"
    return out


def run_ours(args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    distributed = world > 1
    rows = args.rows
    row0 = rank * rows
    e2e_rows = args.e2e_rows
    cores = os.cpu_count() or 1
    my_cores = max(1, cores // max(local_world, 1))
    # bounce-copy workers of the pageable ingest: 4 - 8 saturate one PCIe link (profiles/r02_tuning.md); more
    # only contend for host memory bandwidth, and N ranks share the host's cores
    os.environ.setdefault("VINUM_B200_INGEST_THREADS", str(max(2, min(8, my_cores))))

    # ---- what the answers must be (NumPy over the regenerated rows; before CUDA is touched) ----
    want = want_e2e = None
    if not args.no_verify:
        want = host_expected(row0, rows, my_cores)
        want_e2e = host_expected(row0, e2e_rows, my_cores) if e2e_rows != rows else want

    import vinum_b200 as vb
    from vinum_b200 import _lib as L, datagen, ops, sharded
    from vinum_b200.aggregate import Aggregator
    import pyarrow as pa
    lib = vb.lib
    numa = sharded.bind_to_gpu_numa(local_rank) if distributed else None
    lib.vk_set_device(local_rank)
    if distributed:
        import torch
        import torch.distributed as dist
        sharded.init("nccl")
        from vinum_b200.dist import DistributedAggregator, close_peer_windows
        _w = torch.zeros(world * 1024, dtype=torch.int64, device="cuda")
        for _ in range(3):   # NCCL sets its channels up lazily: part of start-up, not of a step
            dist.all_gather_into_tensor(_w, _w[:1024].clone())
            dist.all_reduce(_w[:8])
        torch.cuda.synchronize()
    st = vb.default_stream()

    # ---- resident table T8 (all 8 columns generated in HBM; the query touches 3) ----
    t8 = datagen.device_table(datagen.T8_COLUMNS, row0, rows, stream=st)
    st.sync()
    key, predc, val = t8.column("i0"), t8.column("f0"), t8.column("f1")
    key_t = pa.int64()
    spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())]

    def barrier():
        st.sync()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    kernel_ms, kernel_launches, kernel_rows = 0.0, 0, 0
    last = {}

    def step(profile: bool):
        nonlocal kernel_ms, kernel_launches, kernel_rows
        agg = Aggregator([key_t], spec)
        if profile:
            agg.profile(True)
        pred = ops.Predicate.compare(predc, ">", 0.5)
        if distributed:
            d = DistributedAggregator(agg, st)
            d.update([key], [None, val], pred)
            last["path"] = agg.last_path          # the LOCAL aggregate's kernel path
            raw = d.finish()
            last["exchange"] = getattr(d, "exchange_mode_used", None)
            inner = d.agg
            if raw is not None:                    # rank 0: ORDER BY i0 over the 1000 merged groups (config 5)
                order = np.argsort(raw[0][0].view(np.int64), kind="stable")
                last["keys"], last["count"], last["sum"] = raw[0][0].view(np.int64)[order], raw[2][order], raw[3][1].view(np.float64)[order]
        else:
            agg.update([key], [None, val], pred, st)
            raw = agg.result_raw(st)
            inner = agg
            last["path"] = agg.last_path
            last["keys"], last["count"], last["sum"] = raw[0][0].view(np.int64), raw[2], raw[3][1].view(np.float64)
        if profile:
            ms, ln, rw = agg.profile_read(1)
            kernel_ms += ms
            kernel_launches += ln
            kernel_rows += rw
        inner.close()
        if inner is not agg:
            agg.close()

    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.vk_event_create(C.byref(e0))
    lib.vk_event_create(C.byref(e1))
    # Python's cyclic collector walks every tracked object of the process (hundreds of thousands once
    # torch is imported: a 50-100 ms pause that lands in a random step).  Nothing in a step creates
    # cycles; collect now and keep the collector out of the timed regions.  This and the NVML start-up
    # come BEFORE the warm-up steps: an idle gap of tens of milliseconds right before the timed region
    # lets the SM clock drop, and the first timed step then ran 0.2-0.3 ms slow.
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.prepare()
    for _ in range(args.warmup):
        step(False)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = lib.vk_launch_count()
    lib.vk_event_record(e0, st.ptr)
    step_wall_ms = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        step(True)
        step_wall_ms.append(round((time.perf_counter() - t_s) * 1e3, 3))
    lib.vk_event_record(e1, st.ptr)
    lib.vk_event_sync(e1)
    barrier()
    launches1 = lib.vk_launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms = C.c_float()
    lib.vk_event_elapsed_ms(e0, e1, C.byref(ms))
    total_ms = float(ms.value)
    if distributed:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = rows * world / (ms_per_step / 1e3)

    # ---- the timed result against the NumPy check (outside the timed region, every N) ----
    verified = {"checked": False}
    if want is not None:
        wc, ws = want
        if distributed:
            tc = torch.from_numpy(wc).cuda()
            tsm = torch.from_numpy(ws).cuda()
            dist.all_reduce(tc)
            dist.all_reduce(tsm)
            wc, ws = tc.cpu().numpy(), tsm.cpu().numpy()
        if rank == 0:
            check_groups(last["keys"], last["count"], last["sum"], wc, ws, f"resident step at N={world}")
            if distributed:
                assert np.array_equal(last["keys"], np.arange(1000)), "rank 0's result is not ordered by i0"
            verified = {"checked": True, "rows": rows * world, "groups": 1000, "selected_rows": int(wc.sum()),
                        "how": "all 1000 COUNT(*) exact and SUM(f1) within 1e-6 of np.bincount over the regenerated rows "
                               "of every shard (sum over ranks)"}

    # ---- end to end through the public host API ----
    e2e_query = QUERY + (" ORDER BY i0" if distributed else "")
    np_cols = {n: datagen.host_column(n, row0, e2e_rows) for n in ("i0", "f0", "f1")}

    def e2e_run(table_of, steps):
        """table_of() -> host pyarrow.Table; returns (seconds per step, last result, stats)."""
        def once():
            tbl = vb.Table.from_arrow(table_of())
            if distributed:
                tbl = tbl.shard()
            res = tbl.sql(e2e_query)
            return res.to_arrow(), tbl.last_stats
        once()
        once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out, stats = once()
        st.sync()
        if distributed:
            dist.barrier()
        dt = (time.perf_counter() - t0) / steps
        if distributed:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, out, stats

    e2e_steps = max(1, min(args.steps, 5))
    pageable = pa.table({n: pa.array(a) for n, a in np_cols.items()})          # zero-copy views of NumPy memory
    e2e_t, out, stats = e2e_run(lambda: pageable, e2e_steps)
    assert stats.get("streamed") and stats.get("agg_path") == 1, stats
    if rank == 0:
        assert out.num_rows == 1000
        if want_e2e is not None:
            wc, ws = want_e2e
            if distributed:
                pass   # summed below
    if want_e2e is not None:
        wc, ws = want_e2e
        if distributed:
            tc, tsm = torch.from_numpy(wc).cuda(), torch.from_numpy(ws).cuda()
            dist.all_reduce(tc)
            dist.all_reduce(tsm)
            wc, ws = tc.cpu().numpy(), tsm.cpu().numpy()
        if rank == 0:
            check_groups(out.column(0).to_numpy(), out.column(1).to_numpy(), out.column(2).to_numpy(), wc, ws,
                         f"e2e query at N={world}")
            if distributed:
                assert np.array_equal(out.column(0).to_numpy(), np.arange(1000)), "ORDER BY i0 violated"
            verified["e2e_checked"] = True
    e2e_stats = dict(stats)
    del pageable
    pinned_cols = {n: vb.pinned_array(a) for n, a in np_cols.items()}
    del np_cols
    pinned = pa.table({n: pa.array(a) for n, a in pinned_cols.items()})
    pin_t, out2, _ = e2e_run(lambda: pinned, e2e_steps)
    del pinned, pinned_cols

    gc.enable()
    if rank != 0:
        if distributed:
            close_peer_windows()
            dist.destroy_process_group()
        return {}

    peak, peak_src = _peaks()
    h2d = int(e2e_stats.get("h2d_bytes", 0)) * world
    result = {
        "metric": "rows/sec filter->hash-agg", "value": value, "unit": "rows/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "northstar_filter_hashagg_T8", "query": QUERY + (" ORDER BY i0" if distributed else ""),
                   "rows_per_gpu": rows, "table": "T8: 8 columns int64/float64 resident in HBM, 3 touched",
                   "l2_policy": "inputs (24 GB touched per step) larger than L2; no flush",
                   "groups": int(len(last.get("keys", []))), "selected_rows": int(np.sum(last.get("count", [0]))),
                   "agg_path": last.get("path"), "exchange": last.get("exchange"), "parallelism": f"row-range x{world}",
                   "numa_cores": (f"{numa[0]}-{numa[-1]}" if numa else None)},
        "gpu_launches": int(launches1 - launches0),
        "step_wall_ms": step_wall_ms,  # host wall clock of each timed step (every step ends with a D2H read)
        "clocks": clocks,
        "verified": verified,
        "e2e": {"value": e2e_rows * world / e2e_t, "unit": "rows/s", "rows_per_step": e2e_rows * world,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(e2e_stats.get("d2h_bytes", 0)),
                "ms_per_step": e2e_t * 1e3, "host_memory": "pageable", "ingest_threads_per_gpu": int(lib.raw.vk_ingest_threads()),
                "exchange": e2e_stats.get("exchange"),
                "api": ("vinum_b200.Table.from_arrow(pa.table(numpy arrays)).sql(QUERY)" if not distributed else
                        "one logical query: Table.from_arrow(shard).shard().sql(QUERY ORDER BY i0) on every rank, answer on rank 0")},
        "e2e_pinned": {"value": e2e_rows * world / pin_t, "unit": "rows/s", "ms_per_step": pin_t * 1e3,
                       "host_memory": "page-locked by the caller (vb.pinned_array): plain DMA, no bounce copy"},
    }
    if kernel_launches:
        per_launch_ms = kernel_ms / kernel_launches
        per_launch_bytes = ALG_BYTES_PER_ROW * (kernel_rows / kernel_launches)
        achieved = per_launch_bytes / (per_launch_ms / 1e3) / 1e9
        tpr = _traffic_per_row()
        result["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                              "frac": achieved / peak,
                              "traffic": None if tpr is None else tpr * (kernel_rows / kernel_launches),
                              "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per row of one "
                                                "captured launch (profiles/r02_traffic.json) x rows per launch",
                              "algorithmic_bytes_per_launch": per_launch_bytes, "peak_source": peak_src,
                              "kernel": "agg_fast_kernel (rank 0's launches)" if distributed else "agg_fast_kernel",
                              "launches": kernel_launches, "avg_launch_ms": per_launch_ms,
                              "algorithmic_bytes_per_row": ALG_BYTES_PER_ROW,
                              "step_frac": ALG_BYTES_PER_ROW * rows / (ms_per_step / 1e3) / 1e9 / peak}
    if world == 1:
        if not args.no_configs:
            result["configs"] = run_configs(vb, t8, rows, st, peak)
        result["cpu_baseline"] = cpu_baseline_single(args.cpu_rows)
        result["dropin_reference_api"] = dropin_reference_api(e2e_rows)
    if distributed:
        close_peer_windows()
        dist.destroy_process_group()
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=int(os.environ.get("VK_BENCH_ROWS", 1_000_000_000)),
                    help="rows per GPU of the resident table")
    ap.add_argument("--e2e-rows", type=int, default=int(os.environ.get("VK_BENCH_E2E_ROWS", 100_000_000)))
    ap.add_argument("--cpu-rows", type=int, default=int(os.environ.get("VK_BENCH_CPU_ROWS", 20_000_000)))
    ap.add_argument("--ref-rows", type=int, default=int(os.environ.get("VK_BENCH_REF_ROWS", 20_000_000)),
                    help="reference arm: rows per step (and per process of the all-cores extra)")
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm, all-cores extra: processes (0 = all host cores)")
    ap.add_argument("--no-verify", action="store_true", help="skip the NumPy check of the timed results")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config sub-records (N = 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        print(json.dumps(run_reference(args)), flush=True)
        return
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    res = run_ours(args)
    if res:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
