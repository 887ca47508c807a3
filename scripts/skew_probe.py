"""Fused filter -> hash aggregate under key skew: Grows/s of
    SELECT k, COUNT(*), SUM(v) FROM t WHERE p > 0.5 GROUP BY k
for uniform keys, Zipf(1.0)-like, 90 % one key, 100 % one key -- every result checked against np.bincount.
    python scripts/skew_probe.py [ROWS]"""
import ctypes as C, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa
import vinum_b200 as vb
from vinum_b200 import _lib as L, datagen, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
vb.lib.vk_set_device(0)
st = vb.default_stream()
lib = vb.lib
rng = np.random.default_rng(7)
p_d = datagen.device_column("f0", 0, n, stream=st)
v_d = datagen.device_column("f1", 0, n, stream=st)
p_h = datagen.host_column("f0", 0, n) > 0.5
v_h = datagen.host_column("f1", 0, n)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.vk_event_create(C.byref(e0)); lib.vk_event_create(C.byref(e1))
spec = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())]


def keys(kind):
    if kind == "uniform":
        return rng.integers(0, 1000, n)
    if kind == "zipf1.0":   # P(k) ~ 1/k over 1000 keys
        w = 1.0 / np.arange(1, 1001)
        return rng.choice(1000, size=n, p=w / w.sum()).astype(np.int64)
    if kind == "hot90":
        k = rng.integers(0, 1000, n)
        k[rng.random(n) < 0.9] = 500
        return k
    return np.full(n, 500, dtype=np.int64)


out = {}
for kind in ("uniform", "zipf1.0", "hot90", "hot100"):
    k_h = keys(kind)
    k_d = vb.DeviceColumn.from_numpy(k_h, st)
    want_c = np.bincount(k_h[p_h], minlength=1000)
    want_s = np.bincount(k_h[p_h], weights=v_h[p_h], minlength=1000)
    row = {}
    for name, opts in (("default", {}), ("no_hot_step", {"AGG_HOT": 0}), ("hot_step_always", {"AGG_HOT": 2}), ("global_table", {"AGG_NOFAST": 1})):
        ts = []
        with vb.options(**opts):
            for i in range(5):
                lib.vk_event_record(e0, st.ptr)
                agg = vb.Aggregator([pa.int64()], spec)
                agg.update([k_d], [None, v_d], ops.Predicate.compare(p_d, ">", 0.5), st)
                raw = agg.result_raw(st)
                lib.vk_event_record(e1, st.ptr); lib.vk_event_sync(e1)
                ms = C.c_float(); lib.vk_event_elapsed_ms(e0, e1, C.byref(ms))
                agg.close()
                if i >= 2:
                    ts.append(ms.value)
        kk = raw[0][0].view(np.int64)
        got_c = np.zeros(1000, dtype=np.int64); got_c[kk] = raw[2]
        got_s = np.zeros(1000); got_s[kk] = raw[3][1].view(np.float64)
        ok = bool(np.array_equal(got_c, want_c) and np.allclose(got_s, want_s, rtol=1e-6, atol=1e-6))
        row[name] = {"ms": round(statistics.median(ts), 3), "Grows_s": round(n / statistics.median(ts) / 1e6, 1), "ok": ok}
    out[kind] = row
    print(json.dumps({kind: row}), flush=True)
