"""Tiny driver for ncu: one fused filter->aggregate update. usage: prof_agg.py DIRECT [N_ROWS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VINUM_B200_AGG_DIRECT"] = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000_000
import pyarrow as pa
import vinum_b200 as vb
from vinum_b200 import _lib as L, datagen, ops
vb.lib.vk_set_device(0)
st = vb.default_stream()
cols = datagen.device_table(["i0", "f0", "f1"], 0, n, stream=st)
st.sync()
for _ in range(2):
    agg = vb.Aggregator([pa.int64()], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())])
    agg.update([cols.column("i0")], [None, cols.column("f1")], ops.Predicate.compare(cols.column("f0"), ">", 0.5), st)
    print(agg.num_groups(st))
    agg.close()
