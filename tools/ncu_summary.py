#!/usr/bin/env python3
"""Key rows of an `ncu --set full` report as text (what profiles/*_ncu_full_*.txt hold): `ncu -i REP --page raw --csv`
filtered to the metrics the design notes argue with.  Usage: python tools/ncu_summary.py report.ncu-rep [kernel-regex] > out.txt"""
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "Kernel Name", "Block Size", "Grid Size",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "launch__cluster_dim_x", "launch__cluster_max_active", "launch__occupancy_limit_shared_mem",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if pat and not pat.search(name):
        continue
    for i, h in enumerate(hdr):
        base = h.split(".TriageCompute.")[-1]
        if h in KEEP or base in KEEP:
            print(f"{h} [{units[i]}] = {r[i]}")
    print()
