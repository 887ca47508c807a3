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
`Table.sql(QUERY)` -- on a host Arrow table in pinned memory: every step parses and plans
the SQL, copies the three referenced columns host->device (double-buffered chunks) and
reads the result back.

N > 1 (torchrun): each rank owns its own 1e9-row shard (weak scaling), aggregates it
locally, repartitions the partial groups with one NCCL all-to-all and gathers on rank 0.

--impl reference: the reference's own C++/NumPy operator chain (oracle/_ref, the
unmodified reference sources compiled by oracle/build_ref.sh) on the host cores.
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
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
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


# ---------------------------------------------------------------- reference arm ----
def _ref_worker(args):
    """One process of the reference arm: regenerate a row range, run the reference chain."""
    row0, rows, batch = args
    from oracle import ref
    from vinum_b200.datagen import host_table  # pure-NumPy generator (no CUDA call)
    table = host_table(["i0", "f0", "f1"], row0, rows)
    t0 = time.perf_counter()
    out = ref.ref_filter_hash_aggregate(table, "f0", ">", 0.5, ["i0"], FUNCS, batch_size=batch)
    return time.perf_counter() - t0, out.num_rows


def run_reference(args) -> dict:
    from oracle import ref
    if ref.ref_lib() is None:
        return {"impl": "reference", "unavailable": "oracle/_ref (compiled reference operators) is not built"}
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, args.ref_procs if args.ref_procs > 0 else cores))
    rows_per_proc = args.ref_rows
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(procs) as pool:
        for step in range(args.warmup + args.steps):
            jobs = [(p * rows_per_proc, rows_per_proc, 10000) for p in range(procs)]
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, jobs)
            wall = time.perf_counter() - t0
            # data generation happens inside the workers but outside their timed region:
            # the step time is the slowest worker's reference run
            step_t = max(r[0] for r in res)
            assert all(r[1] == 1000 for r in res)
            if step >= args.warmup:
                times.append(step_t)
            del wall
    total_rows = rows_per_proc * procs
    t = statistics.median(times)
    value = total_rows / t
    sample = (f"{procs} processes x {rows_per_proc} rows each (disjoint row ranges of the same generator), "
              f"reference chain TableBatchReader(10000) -> numpy compare -> RecordBatch.filter -> "
              f"SingleNumericalHashAggregate; max over processes; partial-result merge not timed")
    return {
        "impl": "reference", "metric": "rows/sec filter->hash-agg", "value": value, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "northstar_filter_hashagg_T8", "query": QUERY, "rows_per_step": total_rows},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": procs, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def cpu_baseline_single(rows: int) -> dict:
    """Reference chain on ONE core (the reference is single-threaded by design,
    vinum/executor/executor.py:24-31) over a bounded sample of the same workload."""
    from oracle import ref
    if ref.ref_lib() is None:
        return {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
    from vinum_b200 import datagen
    table = datagen.host_table(["i0", "f0", "f1"], 0, rows)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        out = ref.ref_filter_hash_aggregate(table, "f0", ">", 0.5, ["i0"], FUNCS, batch_size=10000)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert out.num_rows == 1000
    # SURVEY 8d also asks for the reference at batch 1e6 (its per-batch Python overhead amortised)
    t0 = time.perf_counter()
    out = ref.ref_filter_hash_aggregate(table, "f0", ">", 0.5, ["i0"], FUNCS, batch_size=1_000_000)
    big = time.perf_counter() - t0
    assert out.num_rows == 1000
    return {"value": rows / best, "unit": "rows/s", "cores": 1, "kind": "reference",
            "sample": f"first {rows} rows of the same generator, batch 10000, best of 2 "
                      f"(oracle/_ref = unmodified reference C++ operators + the reference's NumPy/Arrow calls)",
            "value_batch_1e6": rows / big}


# --------------------------------------------------------------------- our arm ----
def run_ours(args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    import vinum_b200 as vb
    from vinum_b200 import _lib as L, datagen, ops
    from vinum_b200.aggregate import Aggregator
    import pyarrow as pa
    lib = vb.lib
    lib.vk_set_device(local_rank)
    if distributed:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from vinum_b200.dist import DistributedAggregator
        # NCCL sets its channels up lazily on the first collectives of each kind: part of start-up
        _w = torch.zeros(world * 1024, dtype=torch.int64, device="cuda")
        for _ in range(3):
            dist.all_gather_into_tensor(_w, _w[:1024].clone())
            dist.all_reduce(_w[:8])
        torch.cuda.synchronize()
    st = vb.default_stream()
    rows = args.rows
    row0 = rank * rows

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
    last_result = {}

    def step(profile: bool):
        nonlocal kernel_ms, kernel_launches, kernel_rows
        agg = Aggregator([key_t], spec)
        if profile:
            agg.profile(True)
        pred = ops.Predicate.compare(predc, ">", 0.5)
        if distributed:
            d = DistributedAggregator(agg, st)
            d.update([key], [None, val], pred)
            raw = d.finish()
            inner = d.agg
        else:
            agg.update([key], [None, val], pred, st)
            raw = agg.result_raw(st)
            inner = agg
        if profile and not distributed:
            ms, ln, rw = agg.profile_read(1)
            kernel_ms += ms
            kernel_launches += ln
            kernel_rows += rw
        if raw is not None:
            last_result["groups"] = int(len(raw[2]))
            last_result["count"] = int(raw[2].sum())
        last_result["path"] = inner.last_path
        inner.close()

    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.vk_event_create(C.byref(e0))
    lib.vk_event_create(C.byref(e1))
    for _ in range(args.warmup):
        step(False)
    # Python's cyclic collector walks every tracked object of the process (hundreds of thousands once
    # torch is imported: a 50-100 ms pause that lands in a random step).  Nothing in a step creates
    # cycles; collect now and keep the collector out of the timed regions.
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.prepare()
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

    # ---- end to end through the public host API (pinned host table, rank-local) ----
    # The call a user makes: Table.sql(<the query>) on a host Arrow table (pinned buffers).
    e2e_rows = args.e2e_rows
    host_cols = {n: vb.pinned_array(datagen.host_column(n, row0, e2e_rows)) for n in ("i0", "f0", "f1")}
    host_table = pa.table({n: pa.array(a) for n, a in host_cols.items()})
    user_table = vb.Table.from_arrow(host_table)
    user_table.sql(QUERY)  # warm-up
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        out = user_table.sql(QUERY).to_arrow()
    st.sync()
    stats = user_table.last_stats
    assert stats.get("streamed") and stats.get("agg_path") == 1, stats
    e2e_t = (time.perf_counter() - t0) / e2e_steps
    if distributed:
        t = torch.tensor([e2e_t], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    e2e_value = e2e_rows * world / e2e_t
    assert out.num_rows == 1000

    gc.enable()
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return {}

    peak, peak_src = _peaks()
    result = {
        "metric": "rows/sec filter->hash-agg", "value": value, "unit": "rows/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "northstar_filter_hashagg_T8", "query": QUERY, "rows_per_gpu": rows,
                   "table": "T8: 8 columns int64/float64 resident in HBM, 3 touched",
                   "l2_policy": "inputs (24 GB touched per step) larger than L2; no flush",
                   "groups": last_result.get("groups"), "selected_rows": last_result.get("count"),
                   "agg_path": last_result.get("path"), "parallelism": f"row-range x{world}"},
        "gpu_launches": int(launches1 - launches0),
        "step_wall_ms": step_wall_ms,  # host wall clock of each timed step (every step ends with a D2H read)
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "rows/s", "rows_per_step": e2e_rows * world,
                "h2d_bytes_per_step": int(stats.get("h2d_bytes", 0)) * world,
                "d2h_bytes_per_step": int(stats.get("d2h_bytes", 0)), "ms_per_step": e2e_t * 1e3,
                "api": "vinum_b200.Table.from_arrow(host pyarrow.Table, pinned buffers).sql(QUERY)"},
    }
    if not distributed and kernel_launches:
        per_launch_ms = kernel_ms / kernel_launches
        per_launch_bytes = ALG_BYTES_PER_ROW * (kernel_rows / kernel_launches)
        achieved = per_launch_bytes / (per_launch_ms / 1e3) / 1e9
        tpr = _traffic_per_row()
        result["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                              "frac": achieved / peak,
                              "traffic": None if tpr is None else tpr * (kernel_rows / kernel_launches),
                              "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per row of one "
                                                "captured launch (profiles/r01_traffic.json) x rows per launch",
                              "algorithmic_bytes_per_launch": per_launch_bytes, "peak_source": peak_src,
                              "kernel": "agg_fast_kernel", "launches": kernel_launches,
                              "avg_launch_ms": per_launch_ms,
                              "algorithmic_bytes_per_row": ALG_BYTES_PER_ROW}
    else:
        achieved = ALG_BYTES_PER_ROW * rows / (ms_per_step / 1e3) / 1e9
        result["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                              "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                              "kernel": "agg_fast_kernel (per-GPU step time incl. exchange)"}
    if world == 1:
        result["cpu_baseline"] = cpu_baseline_single(args.cpu_rows)
    if distributed:
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
    ap.add_argument("--ref-rows", type=int, default=int(os.environ.get("VK_BENCH_REF_ROWS", 10_000_000)),
                    help="reference arm: rows per process per step")
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm: processes (0 = all host cores)")
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
