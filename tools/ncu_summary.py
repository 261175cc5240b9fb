#!/usr/bin/env python
"""Key metrics of one kernel from an .ncu-rep (read on the CPU box): python tools/ncu_summary.py rep.ncu-rep > profiles/x.csv"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "sm__cycles_elapsed.max"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    w = csv.writer(sys.stdout)
    for v in rows[2:]:
        for i, k in enumerate(h):
            stall = "warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")
            if k in KEYS or stall:
                w.writerow([k, u[i], v[i]])


if __name__ == "__main__":
    main()
