#!/usr/bin/env python
"""Summarise an .ncu-rep: key metrics per kernel + top source lines by stall samples.
usage: python tools/ncu_summary.py report.ncu-rep [--source N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nsrc = int(sys.argv[sys.argv.index("--source") + 1]) if "--source" in sys.argv else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct"]
units = rows[1]
for r in rows[2:]:
    print("==", r[idx["Kernel Name"]][:70])
    for k in want:
        if k in idx and r[idx[k]] not in ("", "n/a"):
            print("   %-78s %s %s" % (k, r[idx[k]], units[idx[k]]))
if nsrc:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    print(src[:200])
