"""Summarise `ncu --page source --csv` output: top SASS lines by stall samples, with the
stall-reason breakdown.  usage: ncu_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hdr_i + 1:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
tot_inst = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
print("total samples", tot, "total warp-instructions", tot_inst)
agg = {}
for r in data:
    for h in stall_cols:
        agg[h] = agg.get(h, 0) + int(r[col[h]] or 0)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:n]
for i in sorted(idx):
    r = data[i]
    st = {h[6:]: int(r[col[h]] or 0) for h in stall_cols if int(r[col[h]] or 0)}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{i:5d} {int(r[col['# Samples']]):6d} exec={r[col['Instructions Executed']]:>9} thr={r[col['Avg. Threads Executed']]:>5} {r[col['Source']].strip()[:70]:70s} {top}")
