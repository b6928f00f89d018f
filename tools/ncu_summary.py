#!/usr/bin/env python3
"""Condense an `ncu --page raw --csv` export into the handful of numbers the roofline discussion needs."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        print("kernel:", d["Kernel Name"][:60], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>18s} {u.get(k, '')}")


if __name__ == "__main__":
    main(sys.argv[1])
