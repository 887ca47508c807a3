"""Markdown summary of an `ncu --set full` report: headline metrics, stall-reason shares and
the hottest SASS lines.  usage: ncu_summary.py report.ncu-rep > profiles/xxx.md"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpc__cycles_elapsed.avg.per_second"]
print(f"# ncu --set full: `{rep.split('/')[-1]}`\n")
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    print(f"## `{d['Kernel Name']}`  grid {d['Grid Size']} block {d['Block Size']}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d: print(f"| {k} | {d[k]} | {u[k]} |")
    st = {h[len('smsp__pcsamp_warps_issue_stalled_'):]: int(v) for h, v in zip(hdr, vals)
          if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and v.isdigit()}
    tot = sum(st.values()) or 1
    print("\nWarp stall samples: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]) + "\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
stall_cols = [h for h in rows[hi] if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print(f"## Hottest SASS lines ({tot} samples)\n\n| # | samples | warp-insts | SASS | top stalls |\n|---|---|---|---|---|")
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:25]
for i in sorted(idx):
    r = data[i]
    st = {h[6:]: int(r[col[h]] or 0) for h in stall_cols if int(r[col[h]] or 0)}
    top = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(f"| {i} | {r[col['# Samples']]} | {r[col['Instructions Executed']]} | `{r[col['Source']].strip()[:60]}` | {top} |")
