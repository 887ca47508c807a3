"""Development probe run on the GPU box: quick parity + timing of every kernel family.
Not part of the test-suite (tests/ holds the real parity tests); it exists so that one
gpurun call answers many questions.  Usage: python scripts/gpu_check.py [sections...]
"""
import ctypes as C
import json
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vinum_b200 as vb  # noqa: E402
from vinum_b200 import _lib as L, datagen, ops  # noqa: E402
from vinum_b200.aggregate import Aggregator  # noqa: E402
import pyarrow as pa  # noqa: E402

lib = vb.lib
OUT = {}


def mark(msg):
    print('#', msg, flush=True)


def timed(fn, st, reps=5, warm=2):
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.vk_event_create(C.byref(e0)); lib.vk_event_create(C.byref(e1))
    for _ in range(warm):
        fn()
    st.sync()
    ts = []
    for _ in range(reps):
        lib.vk_event_record(e0, st.ptr)
        fn()
        lib.vk_event_record(e1, st.ptr)
        lib.vk_event_sync(e1)
        ms = C.c_float()
        lib.vk_event_elapsed_ms(e0, e1, C.byref(ms))
        ts.append(ms.value)
    return min(ts), float(np.median(ts))


def section(name):
    def deco(fn):
        def run():
            t = time.time()
            try:
                res = fn()
                OUT[name] = res
                print(json.dumps({name: res}), flush=True)
            except Exception as e:  # noqa: BLE001
                traceback.print_exc()
                OUT[name] = {"error": repr(e)}
                print(json.dumps({name: OUT[name]}), flush=True)
            print(f"# {name} took {time.time() - t:.1f}s", flush=True)
        run.__name__ = name
        return run
    return deco


@section("datagen")
def s_datagen():
    st = vb.default_stream()
    res = {}
    for name in datagen.KINDS:
        dev = datagen.device_column(name, 12345, 1_000_003, stream=st).to_numpy(st)
        host = datagen.host_column(name, 12345, 1_000_003)
        res[name] = bool(np.array_equal(dev.view(np.uint8), host.view(np.uint8)))
    return res


@section("filter")
def s_filter():
    st = vb.default_stream()
    res = {}
    n = 1_000_003
    names = ["i1", "i2", "f0", "f1"]
    batch = datagen.device_table(names, 0, n, stream=st)
    host = {k: datagen.host_column(k, 0, n) for k in names}
    mark("filter small cmp")
    out = ops.filter_batch(batch, ops.Predicate.compare(batch.column("f0"), ">", 0.5), st)
    mark("filter small done")
    m = host["f0"] > 0.5
    ok = out.num_rows == int(m.sum())
    for k in names:
        ok = ok and np.array_equal(out.column(k).to_numpy(st), host[k][m])
    res["cmp_parity"] = bool(ok)
    mark("compare")
    mask = ops.compare(batch.column("f0"), ">", 0.5, st)
    res["mask_parity"] = bool(np.array_equal(mask.to_numpy(st).astype(bool), m))
    out2 = ops.filter_batch(batch, ops.Predicate.from_mask(mask), st)
    res["mask_filter_parity"] = bool(out2.num_rows == int(m.sum()) and
                                     np.array_equal(out2.column("i2").to_numpy(st), host["i2"][m]))
    # timing at C2 size
    mark("filter c2")
    n = 100_000_000
    batch = datagen.device_table(names, 0, n, stream=st)
    vp = ops.Predicate.compare(batch.column("f0"), ">", 0.5).vk()
    vcols = (L.VkColumn * 4)(*[c.vk() for c in batch.columns])
    outs = [vb.DeviceColumn.empty(n, c.dtype, stream=st) for c in batch.columns]
    od = (C.c_void_p * 4)(*[o.data_ptr for o in outs])
    ov = (C.c_void_p * 4)()
    scratch = vb.DeviceBuffer(lib.vk_filter_scratch_bytes(n), st)
    rows = vb.DeviceBuffer(8, st)

    def run():
        lib.vk_filter(C.byref(vp), n, vcols, 4, od, ov, C.c_void_p(rows.ptr), C.c_void_p(scratch.ptr), st.ptr)
    best, med = timed(run, st)
    sel = int(rows.to_numpy(np.int64, 1, st)[0])
    byts = n * 32 + sel * 32
    res["c2_ms_best"] = best
    res["c2_ms_med"] = med
    res["c2_rows_out"] = sel
    res["c2_GBps"] = byts / best / 1e6
    # mask-only compare timing
    mk = vb.DeviceColumn.empty(n, L.BOOL8, stream=st)
    a = batch.column("f0").vk()
    s = L.make_scalar(0.5)
    best, med = timed(lambda: lib.vk_compare_scalar(C.byref(a), L.GT, C.byref(s), C.c_void_p(mk.data_ptr), st.ptr), st)
    res["compare_ms"] = best
    res["compare_GBps"] = n * 9 / best / 1e6
    return res


CFGS = [(1, 0), (0, 0), (1, 8), (0, 8), (1, 4)]
if os.environ.get("VK_CFGS"):
    CFGS = [tuple(int(x) for x in c.split(":")) for c in os.environ["VK_CFGS"].split(",")]


def np_groupby(keys, vals, mask):
    k = keys[mask]
    v = vals[mask]
    uk, inv = np.unique(k, return_inverse=True)
    cnt = np.bincount(inv, minlength=len(uk))
    sm = np.bincount(inv, weights=v, minlength=len(uk))
    return uk, cnt, sm


def run_agg(keycol, valcol, predcol, n, st, cfg, reps=3):
    """cfg = (direct policy, warps or 0) -> env knobs read by vk_agg_create."""
    direct, warps = cfg
    vb.set_option("AGG_DIRECT", direct)
    vb.set_option("AGG_WARPS", warps)
    pred = ops.Predicate.compare(predcol, ">", 0.5) if predcol is not None else None

    def once():
        agg = Aggregator([keycol.arrow_type], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())])
        agg.update([keycol], [None, valcol], pred, st)
        return agg
    agg = once()
    keys, kv, cnt, lo, hi, valid = agg.result_raw(st)
    path = agg.last_path
    agg.close()
    # timing: create+update only (result is tiny)
    ts = []
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.vk_event_create(C.byref(e0)); lib.vk_event_create(C.byref(e1))
    for _ in range(reps):
        a = Aggregator([keycol.arrow_type], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())])
        # warm table allocation
        st.sync()
        lib.vk_event_record(e0, st.ptr)
        a.update([keycol], [None, valcol], pred, st)
        lib.vk_event_record(e1, st.ptr)
        lib.vk_event_sync(e1)
        ms = C.c_float()
        lib.vk_event_elapsed_ms(e0, e1, C.byref(ms))
        ts.append(ms.value)
        a.close()
    return keys[0], cnt, lo[1].view(np.float64), path, min(ts)


@section("agg_parity")
def s_agg_parity():
    st = vb.default_stream()
    res = {}
    n = 3_000_017
    cols = datagen.device_table(["i0", "k32", "f0", "f1", "i3"], 0, n, stream=st)
    h = {k: datagen.host_column(k, 0, n) for k in ["i0", "k32", "f0", "f1", "i3"]}
    m = h["f0"] > 0.5
    for keyname in ["i0", "k32", "i3"]:
        uk, cnt, sm = np_groupby(h[keyname], h["f1"], m)
        for cfg in CFGS:
            for _once in (0,):
                strat = "d%d_w%d" % cfg
                mark(f"agg parity {keyname} cfg {strat}")
                k, c, s, path, ms = run_agg(cols.column(keyname), cols.column("f1"), cols.column("f0"), n, st, cfg, reps=1)
                kk = k.view(np.int64)
                order = np.argsort(kk)
                ok = (len(kk) == len(uk) and np.array_equal(kk[order], uk.astype(np.int64)) and
                      np.array_equal(c[order], cnt.astype(np.uint64)) and
                      np.allclose(s[order], sm, rtol=1e-9, atol=1e-6))
                res[f"{keyname}_{strat}"] = {"ok": bool(ok), "groups": int(len(kk)), "path": path, "ms": ms}
    # no predicate, general path via nulls-free int64 key but MIN/MAX funcs
    mark("agg general")
    agg = Aggregator([pa.int64()], [(L.AGG_COUNT_STAR, None), (L.AGG_MIN, pa.float64()), (L.AGG_MAX, pa.int64()),
                                    (L.AGG_SUM, pa.int64()), (L.AGG_AVG, pa.int64())])
    agg.update([cols.column("i0")], [None, cols.column("f1"), cols.column("i3"), cols.column("i3"), cols.column("i3")], None, st)
    keys, kv, cnt, lo, hi, valid = agg.result_raw(st)
    order = np.argsort(keys[0].view(np.int64))
    uk, inv = np.unique(h["i0"], return_inverse=True)
    mn = np.full(len(uk), np.inf); np.minimum.at(mn, inv, h["f1"])
    mx = np.full(len(uk), -1, dtype=np.int64); np.maximum.at(mx, inv, h["i3"])
    sm = np.bincount(inv, weights=h["i3"].astype(np.float64)).astype(np.int64)
    c = np.bincount(inv)
    res["general"] = {
        "path": agg.last_path,
        "min": bool(np.array_equal(lo[1].view(np.float64)[order], mn)),
        "max": bool(np.array_equal(lo[2].view(np.int64)[order], mx)),
        "sum": bool(np.array_equal(lo[3].view(np.int64)[order], sm)) and bool(np.all(hi[3] == 0)),
        "avg": bool(np.allclose(lo[4].view(np.float64)[order], sm / c, rtol=1e-12)),
        "cnt": bool(np.array_equal(cnt[order], c.astype(np.uint64))),
    }
    agg.close()
    return res


@section("agg_highcard")
def s_agg_highcard():
    st = vb.default_stream()
    res = {}
    n = int(os.environ.get("VK_HC_ROWS", 20_000_000))
    cols = datagen.device_table(["i1", "f1", "f0"], 0, n, stream=st)
    h = {k: datagen.host_column(k, 0, n) for k in ["i1", "f1", "f0"]}
    # ~n distinct keys: exercises replay list + table growth
    t0 = time.time()
    agg = Aggregator([pa.int64()], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())])
    agg.update([cols.column("i1")], [None, cols.column("f1")], None, st)
    g = agg.num_groups(st)
    res["update_s"] = time.time() - t0
    keys, kv, cnt, lo, hi, valid = agg.result_raw(st)
    uk, inv = np.unique(h["i1"], return_inverse=True)
    res["groups"] = int(g)
    res["groups_expected"] = int(len(uk))
    order = np.argsort(keys[0].view(np.int64))
    sm = np.bincount(inv, weights=h["f1"])
    res["keys_ok"] = bool(np.array_equal(keys[0].view(np.int64)[order], uk))
    res["sum_ok"] = bool(np.allclose(lo[1].view(np.float64)[order], sm, rtol=1e-9, atol=1e-9))
    res["cnt_ok"] = bool(np.array_equal(cnt[order], np.bincount(inv).astype(np.uint64)))
    res["path"] = agg.last_path
    agg.close()
    return res


@section("agg_bench")
def s_agg_bench():
    st = vb.default_stream()
    res = {}
    n = int(os.environ.get("VK_BENCH_ROWS", 1_000_000_000))
    cols = datagen.device_table(["i0", "f0", "f1"], 0, n, stream=st)
    k32 = datagen.device_column("k32", 0, n, stream=st)
    st.sync()
    for cfg in CFGS:
        name = "northstar_d%d_w%d" % cfg
        try:
            k, c, s, path, ms = run_agg(cols.column("i0"), cols.column("f1"), cols.column("f0"), n, st, cfg)
            res[name] = {"ms": ms, "GBps": n * 24 / ms / 1e6, "Grows": n / ms / 1e6,
                         "groups": int(len(k)), "path": path, "cnt": int(c.sum())}
        except Exception as e:  # noqa: BLE001
            res[name] = {"error": repr(e)}
        print(json.dumps({name: res[name]}), flush=True)
    for cfg in CFGS:
        name = "c3_d%d_w%d" % cfg
        k, c, s, path, ms = run_agg(k32, cols.column("f1"), None, n, st, cfg)
        res[name] = {"ms": ms, "GBps": n * 12 / ms / 1e6, "Grows": n / ms / 1e6, "groups": int(len(k)), "path": path}
        print(json.dumps({name: res[name]}), flush=True)
    return res


@section("sort")
def s_sort():
    st = vb.default_stream()
    res = {}
    n = 2_000_003
    f3 = datagen.device_column("f3", 0, n, stream=st)
    i0 = datagen.device_column("i0", 0, n, stream=st)
    hf3 = datagen.host_column("f3", 0, n)
    hi0 = datagen.host_column("i0", 0, n)
    mark("sort f3 desc")
    idx = ops.sort_indices([f3], [L.DESC], st).to_numpy(st)
    mark("sort f3 done")
    ref = np.argsort(-hf3, kind="stable")
    res["f3_desc"] = bool(np.array_equal(idx, ref))
    idx = ops.sort_indices([i0], [L.ASC], st).to_numpy(st)
    res["i0_asc_stable"] = bool(np.array_equal(idx, np.argsort(hi0, kind="stable")))
    idx = ops.sort_indices([i0, f3], [L.DESC, L.ASC], st).to_numpy(st)
    ref = np.lexsort((hf3, -hi0))
    res["multi"] = bool(np.array_equal(idx, ref))
    n = 100_000_000
    f3 = datagen.device_column("f3", 0, n, stream=st)
    out = vb.DeviceColumn.empty(n, L.I64, stream=st)
    scratch = vb.DeviceBuffer(lib.vk_sort_scratch_bytes(n), st)
    v = (L.VkColumn * 1)(f3.vk())
    o = (C.c_int32 * 1)(L.DESC)
    best, med = timed(lambda: lib.vk_sort_indices(v, o, 1, n, C.c_void_p(out.data_ptr), C.c_void_p(scratch.ptr), st.ptr), st, reps=3, warm=1)
    res["c4_ms"] = best
    res["c4_Mrows_s"] = n / best / 1e3
    sk = vb.DeviceColumn.empty(n, L.F64, stream=st)
    best, med = timed(lambda: lib.vk_sort_indices_keys(v, o, 1, n, C.c_void_p(out.data_ptr), C.c_void_p(sk.data_ptr),
                                                       C.c_void_p(scratch.ptr), st.ptr), st, reps=3, warm=1)
    res["c4_with_sorted_key_ms"] = best
    tk = vb.DeviceColumn.empty(n, L.F64, stream=st)
    fv = f3.vk()
    best, med = timed(lambda: lib.vk_take(C.byref(fv), C.c_void_p(out.data_ptr), n, C.c_void_p(tk.data_ptr), None, st.ptr), st, reps=3, warm=1)
    res["take_ms"] = best
    srt = tk.to_numpy(st)
    res["c4_sorted"] = bool(np.all(srt[:-1] >= srt[1:]))
    res["c4_sorted_key_equals_take"] = bool(np.array_equal(sk.to_numpy(st).view(np.uint64), srt.view(np.uint64)))
    return res


@section("arith")
def s_arith():
    st = vb.default_stream()
    res = {}
    n = 1_000_003
    names = ["i1", "i3", "f0", "f1"]
    dev = {k: datagen.device_column(k, 0, n, stream=st) for k in names}
    host = {k: datagen.host_column(k, 0, n) for k in names}
    cases = [("+", "f0", "f1"), ("*", "f1", 2.5), ("/", "i1", "i3"), ("-", 1.0, "f0"), ("%", "i1", 7),
             ("*", "i1", "i3"), ("neg", "f1", None), ("~", "i3", None), ("&", "i1", "i3")]
    ok = True
    with np.errstate(all="ignore"):
        for op, a, b in cases:
            da, db = dev.get(a, a) if isinstance(a, str) else a, dev.get(b, b) if isinstance(b, str) else b
            ha, hb = host.get(a, a) if isinstance(a, str) else a, host.get(b, b) if isinstance(b, str) else b
            got = ops.arith(op, da, db, st).to_numpy(st)
            uf = {"+": np.add, "-": np.subtract, "*": np.multiply, "/": np.divide, "%": np.mod, "&": np.bitwise_and,
                  "neg": np.negative, "~": np.invert}[op]
            want = uf(ha) if hb is None else uf(ha, hb)
            same = got.dtype == want.dtype and np.array_equal(got.view(np.uint8), np.ascontiguousarray(want).view(np.uint8))
            res[f"{a}{op}{b}"] = bool(same)
            ok = ok and same
    res["parity"] = bool(ok)
    n = 100_000_000
    f0 = datagen.device_column("f0", 0, n, stream=st)
    f1 = datagen.device_column("f1", 0, n, stream=st)
    out = vb.DeviceColumn.empty(n, L.F64, stream=st)
    a, b = f0.vk(), f1.vk()
    best, med = timed(lambda: lib.vk_arith(L.ADD, C.byref(a), None, C.byref(b), None, n, L.F64, C.c_void_p(out.data_ptr), st.ptr), st)
    res["add_f64_ms"] = best
    res["add_f64_GBps"] = n * 24 / best / 1e6
    s = L.make_scalar(2.5)
    best, med = timed(lambda: lib.vk_arith(L.MUL, C.byref(a), None, None, C.byref(s), n, L.F64, C.c_void_p(out.data_ptr), st.ptr), st)
    res["mul_scalar_ms"] = best
    res["mul_scalar_GBps"] = n * 16 / best / 1e6
    return res


@section("onegroup")
def s_onegroup():
    """Un-grouped reduction: COUNT(*), SUM(f64), SUM(int64) -> 128-bit, MIN, MAX with and
    without a fused predicate; parity on an odd row count, timing at 1e8 rows."""
    st = vb.default_stream()
    res = {}
    # two aggregates of <= 4 value functions each, so that the fused single-launch kernel
    # (VINUM_B200_ONEGROUP_FUSED=1, at most 4 functions) is eligible for both
    funcs_a = [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64()), (L.AGG_SUM, pa.int64()),
               (L.AGG_MIN, pa.int64()), (L.AGG_MAX, pa.float64())]
    funcs_b = [(L.AGG_COUNT, pa.float64())]

    def run(n, with_pred):
        dev = datagen.device_table(["f0", "f1", "i1"], 0, n, stream=st)
        pred = ops.Predicate.compare(dev.column("f0"), ">", 0.5) if with_pred else None
        got = []
        for funcs, vals in ((funcs_a, [None, dev.column("f1"), dev.column("i1"), dev.column("i1"), dev.column("f0")]),
                            (funcs_b, [dev.column("f1")])):
            agg = Aggregator([], funcs)
            agg.update([], vals, pred, st)
            _, aggs = agg.result_arrays(st)
            agg.close()
            got += [a[0].as_py() for a in aggs]
        return dev, pred, got
    ok = True
    for with_pred in (False, True):
        n = 1_000_003
        _, _, got = run(n, with_pred)
        f0, f1, i1 = (datagen.host_column(k, 0, n) for k in ("f0", "f1", "i1"))
        m = f0 > 0.5 if with_pred else np.ones(n, bool)
        want = [int(m.sum()), float(f1[m].sum()), int(i1[m].astype(object).sum()), int(i1[m].min()), float(f0[m].max()), int(m.sum())]
        same = got[0] == want[0] and abs(got[1] - want[1]) <= 1e-9 * max(1.0, abs(want[1])) * 1e3 and \
            int(got[2]) == want[2] and got[3] == want[3] and got[4] == want[4] and got[5] == want[5]
        res[f"parity_pred{int(with_pred)}"] = bool(same)
        if not same:
            res[f"got_pred{int(with_pred)}"] = [str(x) for x in got]
            res[f"want_pred{int(with_pred)}"] = [str(x) for x in want]
        ok = ok and same
    res["parity"] = bool(ok)
    n = 100_000_000
    dev = datagen.device_table(["f0", "f1"], 0, n, stream=st)
    pred = ops.Predicate.compare(dev.column("f0"), ">", 0.5)

    def once():
        a = Aggregator([], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())])
        a.update([], [None, dev.column("f1")], pred, st)
        return a
    aggs = []
    best, med = timed(lambda: aggs.append(once()), st)
    for a in aggs:
        a.close()
    res["count_sum_pred_ms"] = best
    res["count_sum_pred_GBps_16B_row"] = n * 16 / best / 1e6
    return res


@section("topk")
def s_topk():
    """ORDER BY f3 DESC LIMIT k at C4 size: radix select + sort of the candidates vs the full sort."""
    st = vb.default_stream()
    res = {}
    n = 100_000_000
    f3 = datagen.device_column("f3", 0, n, stream=st)
    full = ops.sort_indices([f3], [L.DESC], st)
    for k in (10, 1000, 100_000):
        top = ops.sort_top([f3], [L.DESC], k, st)
        res[f"k{k}_parity"] = bool(np.array_equal(top.to_numpy(st), full.slice(0, k).to_numpy(st)))
        cand = ops.topk_candidates(f3, L.DESC, k, st)
        res[f"k{k}_candidates"] = None if cand is None else cand.length
        best, med = timed(lambda: ops.sort_top([f3], [L.DESC], k, st), st, reps=3, warm=1)
        res[f"k{k}_ms"] = best
    best, med = timed(lambda: ops.sort_indices([f3], [L.DESC], st), st, reps=3, warm=1)
    res["full_sort_ms"] = best
    return res


SECTIONS = {f.__name__: f for f in [s_datagen, s_filter, s_agg_parity, s_agg_highcard, s_agg_bench, s_sort, s_arith,
                                    s_onegroup, s_topk]}

if __name__ == "__main__":
    lib.vk_set_device(0)
    print(json.dumps({"device": vb.device_info(0)}), flush=True)
    args = sys.argv[1:]
    for a in [a for a in args if "=" in a]:   # NAME=VALUE: a kernel-selection option (vk_set_option)
        name, value = a.split("=", 1)
        vb.set_option(name, int(value))
        print(json.dumps({"option": {name: int(value)}}), flush=True)
    want = [a for a in args if "=" not in a] or list(SECTIONS)
    for name in want:
        SECTIONS[name]()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gpu_check.json", "w") as f:
        json.dump(OUT, f, indent=1)
