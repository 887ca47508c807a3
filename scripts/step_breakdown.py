"""Host-side wall-clock breakdown of one north-star step (create / update / result / close).
usage: step_breakdown.py [N_ROWS]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
import pyarrow as pa
import vinum_b200 as vb
from vinum_b200 import _lib as L, datagen, ops
vb.lib.vk_set_device(0)
st = vb.default_stream()
cols = datagen.device_table(["i0", "f0", "f1"], 0, n, stream=st)
st.sync()
spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())]
for it in range(6):
    t = [time.perf_counter()]
    agg = vb.Aggregator([pa.int64()], spec)
    t.append(time.perf_counter())
    pred = ops.Predicate.compare(cols.column("f0"), ">", 0.5)
    agg.update([cols.column("i0")], [None, cols.column("f1")], pred, st)
    t.append(time.perf_counter())
    st.sync()
    t.append(time.perf_counter())
    raw = agg.result_raw(st)
    t.append(time.perf_counter())
    agg.close()
    t.append(time.perf_counter())
    names = ["create", "update(host)", "sync", "result_raw", "close"]
    print(it, " ".join(f"{nm}={1e3 * (b - a):.3f}ms" for nm, a, b in zip(names, t, t[1:])), f"total={1e3 * (t[-1] - t[0]):.3f}ms",
          flush=True)
