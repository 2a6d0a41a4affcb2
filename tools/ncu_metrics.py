#!/usr/bin/env python3
"""Print selected metrics from an .ncu-rep (ncu -i ... --page raw --csv), one block per kernel launch.

    python tools/ncu_metrics.py report.ncu-rep [substring ...]
"""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_allocated",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "memory_l1_wavefronts_shared_ideal", "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled",
           "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sectors_op_atom",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    for r in rows[2:]:
        print("==", r[name_col][:90] if name_col is not None else "")
        for h, u, v in zip(hdr, units, r):
            if any(p in h for p in pats):
                print(f"  {h} [{u}] = {v}")


if __name__ == "__main__":
    main()
