#!/usr/bin/env python
"""Key metrics of an ncu report as text (the format of profiles/*_summary.txt).

usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_summary.py
"""
import csv
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct")
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("# kernel:", d.get("Kernel Name", "?"), " grid", d.get("Grid Size", "?"), " block", d.get("Block Size", "?"))
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in KEEP) and "per_issue_active" in h or any(h == k or (h.startswith(k) and k.endswith(("limit", "throughput"))) for k in KEEP):
            print("%-90s %-16s %s" % (h, units[i], r[i]))
