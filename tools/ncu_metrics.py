"""Selected raw metrics of an .ncu-rep (ncu --set full) as the text blocks kept under profiles/.
usage: python tools/ncu_metrics.py <report.ncu-rep> [kernel-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]
only = sys.argv[2] if len(sys.argv) > 2 else ""
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "lts__t_sectors_srcunit_tex_aperture_sysmem_op_read.sum"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if only and only not in r[ki]:
        continue
    print("----")
    print(f"{'Kernel Name':<90s} {r[ki]} ")
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:<90s} {r[i]:>15s} {units[i]}")
