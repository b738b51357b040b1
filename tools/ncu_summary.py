#!/usr/bin/env python
"""Headline metrics of every kernel in an ncu report as JSON: usage ncu_summary.py report.ncu-rep [...] > out.json"""
import csv, io, json, subprocess, sys
WANT = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers_per_thread",
        "launch__grid_size": "grid", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct", "sm__cycles_elapsed.max": "sm_cycles",
        "smsp__inst_executed.sum": "warp_instructions", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
        "launch__shared_mem_per_block_dynamic": "smem_dynamic", "launch__shared_mem_per_block_static": "smem_static",
        "launch__cluster_size": "cluster_size"}
out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ks = []
    for r in rows[2:]:
        k = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in WANT:
                try:
                    k[WANT[h]] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    pass
        ks.append(k)
    out[rep.split("/")[-1]] = ks
print(json.dumps(out, indent=1))
