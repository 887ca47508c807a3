"""ncu driver for C3: GROUP BY k32 (int32 key, 1000 groups) COUNT(*), SUM(f1), no predicate. usage: prof_c3.py [N_ROWS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
import pyarrow as pa
import vinum_b200 as vb
from vinum_b200 import _lib as L, datagen
vb.lib.vk_set_device(0)
st = vb.default_stream()
k32 = datagen.device_column("k32", 0, n, stream=st)
f1 = datagen.device_column("f1", 0, n, stream=st)
st.sync()
for _ in range(2):
    agg = vb.Aggregator([pa.int32()], [(L.AGG_COUNT_STAR, None), (L.AGG_SUM, pa.float64())])
    agg.update([k32], [None, f1], None, st)
    print(agg.num_groups(st))
    agg.close()
