#!/usr/bin/env python
"""Print the headline counters of every kernel in an .ncu-rep (development tool).

usage: ncu_summary.py report.ncu-rep [kernel_substring]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if pat not in name:
        continue
    print("---", name)
    for i, h in enumerate(hdr):
        if h in want or (h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h):
            print("  %-70s %-10s %s" % (h, units[i], r[i]))
