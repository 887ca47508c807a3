"""A/B of the fused aggregate's options on ONE box: whole-operator ms (create + update + packed result) for
C3 (int32 key, no predicate), the north-star query (dense keys) and the same on non-dense keys (dictionary).
    python scripts/agg_ab.py [ROWS]"""
import ctypes as C, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyarrow as pa
import vinum_b200 as vb
from vinum_b200 import _lib as L, datagen, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
vb.lib.vk_set_device(0)
st = vb.default_stream()
lib = vb.lib
k32 = datagen.device_column("k32", 0, n, stream=st)
i0 = datagen.device_column("i0", 0, n, stream=st)
hk = ops.arith("*", i0, 2654435761, st)
f0 = datagen.device_column("f0", 0, n, stream=st)
f1 = datagen.device_column("f1", 0, n, stream=st)
spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())]
e0, e1 = C.c_void_p(), C.c_void_p()
lib.vk_event_create(C.byref(e0)); lib.vk_event_create(C.byref(e1))


def run(key, kt, pred_col, thr=0.5, reps=5):
    ts = []
    for i in range(reps + 2):
        lib.vk_event_record(e0, st.ptr)
        agg = vb.Aggregator([kt], spec)
        agg.update([key], [None, f1], ops.Predicate.compare(pred_col, ">", thr) if pred_col is not None else None, st)
        raw = agg.result_raw(st)
        lib.vk_event_record(e1, st.ptr); lib.vk_event_sync(e1)
        ms = C.c_float(); lib.vk_event_elapsed_ms(e0, e1, C.byref(ms))
        assert len(raw[2]) == 1000 and agg.last_path == 1
        agg.close()
        if i >= 2:
            ts.append(ms.value)
    return round(statistics.median(ts), 3)


out = {}
for name, opts in (("default", {}), ("hot_off", {"AGG_HOT": 0}), ("hot_always", {"AGG_HOT": 2})):
    with vb.options(**opts):
        out[name] = {"c3": run(k32, pa.int32(), None), "northstar": run(i0, pa.int64(), f0), "northstar_hash": run(hk, pa.int64(), f0),
                     "northstar_sel09": run(i0, pa.int64(), f0, 0.1)}
    print(json.dumps({name: out[name]}), flush=True)
