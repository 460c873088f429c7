"""Text summary (per captured launch) of an ncu --set full report: duration, DRAM traffic, DMMA /
FP64 pipe utilisation, occupancy, registers, spills.  usage: ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_register_spilling",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
ik = h.index("Kernel Name")
for r in rows[2:]:
    print(r[ik])
    for k in want:
        if k in h:
            i = h.index(k)
            print(f"    {k:82s} {r[i]:>14s} {units[i]}")
    print()
