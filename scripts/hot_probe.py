"""Skewed keys on the global-table kernels: 1e6 groups, every second row on ONE key; update ms with the hot-slot merge
(option AGG_HOT: 1 = chosen from the learning launch's measure, 0 = off) for the global-table kernel and the partitioned plan.
    python scripts/hot_probe.py [ROWS]"""
import ctypes as C, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa
import vinum_b200 as vb
from vinum_b200 import _lib as L, datagen, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
vb.lib.vk_set_device(0)
st = vb.default_stream()
lib = vb.lib
i3 = datagen.device_column("i3", 0, n, stream=st)
i2 = datagen.device_column("i2", 0, n, stream=st)
f0 = datagen.device_column("f0", 0, n, stream=st)
f1 = datagen.device_column("f1", 0, n, stream=st)
key = ops.arith("*", i3, ops.arith("%", i2, 2, st), st)      # even rows -> key 0, odd rows keep one of 1e6 keys
spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())]
e0, e1 = C.c_void_p(), C.c_void_p()
lib.vk_event_create(C.byref(e0)); lib.vk_event_create(C.byref(e1))
want0 = None
for name, opts in (("auto", {}), ("hot_off", {"AGG_HOT": 0}), ("partitioned_auto", {"AGG_PARTITION": 2}),
                   ("partitioned_hot_off", {"AGG_PARTITION": 2, "AGG_HOT": 0}), ("uniform_auto", None), ("uniform_hot_always", None)):
    k = key
    if opts is None:
        k, opts = i3, ({} if name == "uniform_auto" else {"AGG_HOT": 2, "AGG_NOFAST": 1})
    ts = []
    for rep in range(3):
        with vb.options(**opts):
            agg = vb.Aggregator([pa.int64()], spec)
        lib.vk_event_record(e0, st.ptr)
        agg.update([k], [None, f1], ops.Predicate.compare(f0, ">", 0.5), st)
        lib.vk_event_record(e1, st.ptr); lib.vk_event_sync(e1)
        ms = C.c_float(); lib.vk_event_elapsed_ms(e0, e1, C.byref(ms))
        raw = agg.result_raw(st)
        path = agg.last_path
        agg.close()
        if rep:
            ts.append(ms.value)
    keys, cnt = raw[0][0].view(np.int64), raw[2].astype(np.int64)
    c0 = int(cnt[keys == 0][0]) if k is key else None
    if k is key and want0 is None:
        want0 = c0
    print(json.dumps({"case": name, "update_ms": round(min(ts), 3), "path": path, "groups": len(cnt), "rows_on_hot_key": c0,
                      "same_as_first": (c0 == want0) if k is key else None}), flush=True)
