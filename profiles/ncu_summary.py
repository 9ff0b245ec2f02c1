"""Summarise an ncu --set full report (raw CSV page) into the metrics the roofline needs.
Usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_summary.py raw.csv"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__inst_executed.sum", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "smsp__inst_executed_pipe_fp64",
        "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_lsu",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled", "local_load", "local_store", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "derived__smsp__inst_executed_op_local", "smsp__inst_executed_op_local"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[4][:60], "grid", r[8], "block", r[7])
        for i, h in enumerate(hdr):
            if any(k in h for k in KEYS):
                print(f"  {h:100s} {units[i]:14s} {r[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
