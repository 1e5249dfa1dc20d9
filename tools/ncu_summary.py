#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics, stall reasons and the hottest source lines."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"  {w:70s} {vals[i]:>18s} {units[i]}")
print("stall reasons (warps per issue-active cycle):")
st = [(float(vals[i]), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for v, h in sorted(st, reverse=True)[:8]:
    print(f"  {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio',''):28s} {v:8.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if rows:
    if rows[0] and rows[0][0] == 'Kernel Name': rows = rows[1:]
    h = rows[0]
    try:
        ci = h.index("Source"); si = [i for i, x in enumerate(h) if x.startswith("Warp Stall Sampling (All")][0]
        top = sorted(((float(r[si] or 0), r[ci]) for r in rows[1:] if len(r) > si and r[si].replace('.', '').isdigit()), reverse=True)[:14]
        tot = sum(float(r[si] or 0) for r in rows[1:] if len(r) > si and r[si].replace('.', '').isdigit())
        print("hottest SASS (stall samples, % of all):")
        for v, s in top:
            print(f"  {100*v/max(tot,1):5.1f}%  {s[:110]}")
    except Exception as e:
        print("source page parse failed:", e, h[:12])
