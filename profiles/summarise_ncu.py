"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of counters the roofline needs."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"]


def main(path):
    rows = list(csv.reader(open(path)))
    H, U = rows[0], rows[1]
    ki = H.index("Kernel Name")
    for r in rows[2:]:
        if len(r) < len(H):
            continue
        print(f"### {r[ki].split('(')[0]}  (launch id {r[0]})")
        for k in KEYS:
            if k in H:
                i = H.index(k)
                print(f"  {k:90s} {r[i]:>18s} {U[i]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
