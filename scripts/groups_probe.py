"""High-cardinality GROUP BY on ONE box: whole-operator ms (create + update + packed result; `_update` = up to the end of the update's kernels) of the north-star
query with the key folded to G groups (G beyond what the shared-memory tables hold goes to the global table),
as chosen automatically, with the partitioned plan forced (scatter into buckets that are slices of the table + slice-by-slice
update, option AGG_PARTITION), with the four-rows-per-thread global-table kernel and with the one-row-per-thread one.  Counts are
checked against np.bincount for the first setting of every G.
    python scripts/groups_probe.py [ROWS]"""
import ctypes as C, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa
import vinum_b200 as vb
from vinum_b200 import _lib as L, datagen, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
vb.lib.vk_set_device(0)
st = vb.default_stream()
lib = vb.lib
i3 = datagen.device_column("i3", 0, n, stream=st)      # uniform in [0, 1e6)
i1 = datagen.device_column("i1", 0, n, stream=st)
f0 = datagen.device_column("f0", 0, n, stream=st)
f1 = datagen.device_column("f1", 0, n, stream=st)
spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())]
e0, e1, em = C.c_void_p(), C.c_void_p(), C.c_void_p()
lib.vk_event_create(C.byref(e0)); lib.vk_event_create(C.byref(e1)); lib.vk_event_create(C.byref(em))


def run(key, reps=3):
    ts, us, raw, path = [], [], None, None
    for i in range(reps + 1):
        lib.vk_event_record(e0, st.ptr)
        agg = vb.Aggregator([pa.int64()], spec)
        agg.update([key], [None, f1], ops.Predicate.compare(f0, ">", 0.5), st)
        lib.vk_event_record(em, st.ptr)
        raw = agg.result_raw(st)
        lib.vk_event_record(e1, st.ptr); lib.vk_event_sync(e1)
        ms = C.c_float(); lib.vk_event_elapsed_ms(e0, e1, C.byref(ms))
        ms_u = C.c_float(); lib.vk_event_elapsed_ms(e0, em, C.byref(ms_u))
        path = agg.last_path
        agg.close()
        if i >= 1:
            ts.append(ms.value)
            us.append(ms_u.value)
    return round(statistics.median(ts), 3), raw, path, round(statistics.median(us), 3)


host_f0 = host_i3 = None
for groups in (10_000, 100_000, 1_000_000, 1 << 23):
    if groups <= 1_000_000:
        key = i3 if groups == 1_000_000 else ops.arith("%", i3, groups, st)
    else:
        key = ops.arith("&", i1, groups - 1, st)
    line = {"groups": groups}
    for name, opts in (("auto", {}), ("partitioned", {"AGG_PARTITION": 2}), ("wide", {"AGG_PARTITION": 0}), ("one_row", {"AGG_PARTITION": 0, "AGG_WIDE": 0})):
        with vb.options(**opts):
            ms, raw, path, ms_update = run(key)
        line[name] = ms
        line[name + "_update"] = ms_update
        line[name + "_path"] = path
        if name in ("auto", "partitioned") and n <= 200_000_000:
            if host_f0 is None:
                host_f0 = datagen.host_column("f0", 0, n) > 0.5
            hk = key.to_numpy(st)[host_f0]
            want = np.bincount(hk, minlength=groups)
            keys, cnt = raw[0][0].view(np.int64), raw[2].astype(np.int64)
            got = np.zeros(groups, np.int64); got[keys] = cnt
            line[name + "_ok"] = bool(np.array_equal(got, want))
    print(json.dumps(line), flush=True)
