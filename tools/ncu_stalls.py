#!/usr/bin/env python
"""Summarise an ncu report: headline metrics + top stalled SASS instructions.  usage: ncu_stalls.py report.ncu-rep [kernel-id]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "launch__occupancy_per_block_size", "sm__maximum_warps_per_active_cycle_pct", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
for r in rows[2:]:
    print("=== kernel", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if any(h.endswith(w) or h == w for w in want):
            print(f"  {h:90s} {units[i]:12s} {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-id", sys.argv[2]] if len(sys.argv) > 2 and sys.argv[2] not in ("", "-") else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
data = []
for r in rows:
    if "Source" in r and "# Samples" in r:
        h = r
        continue
    if h and len(r) == len(h):
        try:
            data.append((int(r[h.index("# Samples")]), r))
        except ValueError:
            pass
tot = sum(s for s, _ in data) or 1
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
agg = {}
for s, r in data:
    for i in stall_cols:
        if r[i]:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print("total samples", tot, "stall mix:", {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for s, r in sorted(data, key=lambda t: -t[0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    st = {h[i]: int(r[i]) for i in stall_cols if r[i] and int(r[i]) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*s/tot:5.1f}% ex={r[h.index('Instructions Executed')]:>9} {r[h.index('Source')][:80]:80s} {top}")
