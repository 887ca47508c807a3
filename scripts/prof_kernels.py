"""Workloads for ncu captures (one kernel family per invocation, BASELINE sizes):
    python scripts/prof_kernels.py northstar|hash|groups1e6|groups1e5|groups8m|c3|filter|sort|arith|compare|onegroup|topk [ROWS]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyarrow as pa
import vinum_b200 as vb
from vinum_b200 import _lib as L, datagen, ops
what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else (1_000_000_000 if what in ("northstar", "hash", "c3") else 100_000_000)
if what in ("groups1e6", "groups1e5", "groups8m") and len(sys.argv) <= 2:
    n = 200_000_000
vb.lib.vk_set_device(0)
st = vb.default_stream()
spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())]
if what in ("northstar", "hash", "c3", "groups1e6", "groups1e5", "groups8m"):
    key = datagen.device_column("k32" if what == "c3" else ("i3" if what in ("groups1e6", "groups1e5") else "i0"), 0, n, stream=st)
    if what == "hash":
        key = ops.arith("*", key, 2654435761, st)
    if what == "groups8m":
        key = ops.arith("&", datagen.device_column("i1", 0, n, stream=st), (1 << 23) - 1, st)
    if what == "groups1e5":
        key = ops.arith("%", key, 100000, st)
    f1 = datagen.device_column("f1", 0, n, stream=st)
    f0 = datagen.device_column("f0", 0, n, stream=st) if what != "c3" else None
    for _ in range(2):
        agg = vb.Aggregator([pa.int32() if what == "c3" else pa.int64()], spec)
        agg.update([key], [None, f1], ops.Predicate.compare(f0, ">", 0.5) if f0 is not None else None, st)
        print(what, agg.num_groups(st), "path", agg.last_path)
        agg.close()
elif what == "filter":
    t = datagen.device_table(["i1", "i2", "f0", "f1"], 0, n, stream=st)
    for _ in range(3):
        out = ops.filter_batch(t, ops.Predicate.compare(t.column("f0"), ">", 0.5), st)
    print("filter rows", out.num_rows)
elif what == "sort":
    f3 = datagen.device_column("f3", 0, n, stream=st)
    for _ in range(2):
        idx, srt = ops.sort_indices_keys([f3], [L.DESC], st)
    st.sync()
elif what in ("arith", "compare"):
    a, b = datagen.device_column("f0", 0, n, stream=st), datagen.device_column("f1", 0, n, stream=st)
    for _ in range(3):
        r = ops.arith("+", a, b, st) if what == "arith" else ops.compare(a, ">", 0.5, st)
    st.sync()
elif what == "onegroup":
    a, b = datagen.device_column("f0", 0, n, stream=st), datagen.device_column("f1", 0, n, stream=st)
    for _ in range(3):
        agg = vb.Aggregator([], spec)
        agg.update([], [None, b], ops.Predicate.compare(a, ">", 0.5), st)
        agg.result_raw(st)
        agg.close()
elif what == "topk":
    f3 = datagen.device_column("f3", 0, n, stream=st)
    for _ in range(2):
        ops.sort_top([f3], [L.DESC], 1000, st)
    st.sync()
