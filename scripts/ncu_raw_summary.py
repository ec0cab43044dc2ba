"""Pick the metrics that matter out of `ncu -i X.ncu-rep --page raw --csv` (stdin), one block per kernel launch."""
import csv, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct"]
rows = list(csv.reader(sys.stdin))
hdr = rows[0]; units = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    print("----", r[ix["Kernel Name"]][:60])
    for k in KEEP:
        if k in ix:
            print(f"  {k:76s} {r[ix[k]]:>16s} {units[ix[k]]}")
