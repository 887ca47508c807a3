"""Pageable host -> device ingest: rows/s of `Table.from_arrow(pageable).sql(QUERY)` (the e2e of bench.py)
for several bounce-copy piece sizes; run once per worker-thread count (read when the pool starts):
    VINUM_B200_INGEST_THREADS=8 python scripts/ingest_probe.py [ROWS]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa
import vinum_b200 as vb
from vinum_b200 import datagen
Q = "SELECT i0, COUNT(*), SUM(f1) FROM t WHERE f0 > 0.5 GROUP BY i0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
vb.lib.vk_set_device(0)
cols = {c: datagen.host_column(c, 0, n) for c in ("i0", "f0", "f1")}
t = pa.table({c: pa.array(a) for c, a in cols.items()})
out = {"threads": int(vb.lib.raw.vk_ingest_threads()), "rows": n}
for kb in (256, 512, 1024, 2048, 4096):
    vb.set_option("INGEST_PIECE_KB", kb)
    for _ in range(2):
        vb.Table.from_arrow(t).sql(Q)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        r = vb.Table.from_arrow(t).sql(Q).to_arrow()
        ts.append(time.perf_counter() - t0)
    assert r.num_rows == 1000
    out[f"piece_{kb}KB_Grows_s"] = round(n / min(ts) / 1e9, 3)
vb.set_option("INGEST_STAGED", 0)
t0 = time.perf_counter(); vb.Table.from_arrow(t).sql(Q); out["driver_staged_Grows_s"] = round(n / (time.perf_counter() - t0) / 1e9, 3)
print(json.dumps(out), flush=True)
