#!/usr/bin/env python
"""Pick the judged metrics out of `ncu --set full` reports (read here, without a GPU) into one small CSV.

    python tools/ncu_select.py out.csv report1.ncu-rep [report2.ncu-rep ...]
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "sm__inst_executed.sum",
        "smsp__inst_executed.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_dim_x", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_membar",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_dispatch_stall", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_sleeping",
        "smsp__pcsamp_warps_issue_stalled_imc_miss", "smsp__pcsamp_warps_issue_stalled_drain"]

cols, names = [], []
hdr_units = {}
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        names.append(r[idx["Kernel Name"]].split("(")[0])
        cols.append({h: r[idx[h]] for h in WANT if h in idx})
        hdr_units.update({h: units[idx[h]] for h in WANT if h in idx})
with open(sys.argv[1], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + names)
    for h in WANT:
        if h in hdr_units:
            w.writerow([h, hdr_units[h]] + [c.get(h, "") for c in cols])
print(open(sys.argv[1]).read())
