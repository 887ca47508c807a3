"""A/B harness for the fused kernel: for each env-knob setting, 10 fresh north-star steps;
prints kernel-only GB/s (CUDA-event spans around every agg_fast launch) and step ms.
usage: agg_ab.py "K=V,K2=V2" "K=V" ...   (empty string = defaults)"""
import os, sys, time, statistics, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyarrow as pa
import vinum_b200 as vb
from vinum_b200 import _lib as L, datagen, ops
n = int(os.environ.get("VK_BENCH_ROWS", 1_000_000_000))
vb.lib.vk_set_device(0)
st = vb.default_stream()
cols = datagen.device_table(["i0", "f0", "f1"], 0, n, stream=st)
k32 = datagen.device_column("k32", 0, n, stream=st)
st.sync()
spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())]
KNOBS = ["VINUM_B200_AGG_PF", "VINUM_B200_AGG_WARPS", "VINUM_B200_AGG_DIRECT", "VINUM_B200_LIST_LOG2", "VINUM_B200_AGG_LOG2S"]
def run(setting, query):
    for k in KNOBS: os.environ.pop(k, None)
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("="); os.environ["VINUM_B200_" + k if not k.startswith("VINUM") else k] = v
    gbs, steps = [], []
    for it in range(12):
        if query == "northstar":
            key, pred, bpr = cols.column("i0"), ops.Predicate.compare(cols.column("f0"), ">", 0.5), 24
        else:
            key, pred, bpr = k32, None, 12
        t0 = time.perf_counter()
        agg = vb.Aggregator([key.arrow_type], spec)
        agg.profile(True)
        agg.update([key], [None, cols.column("f1")], pred, st)
        raw = agg.result_raw(st)
        t1 = time.perf_counter()
        ms, ln, rows = agg.profile_read(1)
        agg.close()
        if it >= 2:
            gbs.append(rows * bpr / ms / 1e6); steps.append((t1 - t0) * 1e3)
        assert len(raw[2]) == 1000
    print(f"{query:9s} [{setting:28s}] kernel GB/s med={statistics.median(gbs):7.1f} max={max(gbs):7.1f} min={min(gbs):7.1f} | "
          f"step ms med={statistics.median(steps):.3f} min={min(steps):.3f} launches/step={ln} "
          f"steps={[round(x, 2) for x in steps]}", flush=True)
for s in sys.argv[1:] or [""]:
    for q in ("northstar", "c3"):
        run(s, q)
