#!/usr/bin/env python
"""One-kernel summary of an ncu capture for profiles/: the raw-page metrics that the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep "header line" > profiles/rNN_react_kernel_X_ncu_full.txt
"""
import csv
import io
import subprocess
import sys

KEEP = ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gcc__cache_requests_type_instruction",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate", "launch__block_size", "launch__grid_size",
        "launch__occupancy", "launch__registers_per_thread", "launch__shared_mem", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate", "sm__cycles_active.avg", "sm__icc_request_hit_rate", "sm__inst_executed.sum",
        "sm__inst_executed_pipe_fp64", "sm__inst_issued.avg.per_cycle_active", "sm__instruction_throughput",
        "sm__pipe_fp64_cycles_active", "sm__throughput", "sm__warps_active", "smsp__average_warp", "smsp__inst_executed.sum",
        "smsp__issue_active", "smsp__pcsamp_warps_issue_stalled", "smsp__sass_thread_inst_executed_op_d",
        "smsp__thread_inst_executed_per_inst_executed", "smsp__warp_issue_stalled", "smsp__warps_eligible",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "sm__sass_inst_executed_op_shared")


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units, vals = rows[0], rows[1], rows[2]
    if len(sys.argv) > 2:
        print("# " + sys.argv[2])
    for n, u, v in zip(names, units, vals):
        if any(n.startswith(k) for k in KEEP):
            print("%s [%s] = %s" % (n, u, v))


if __name__ == "__main__":
    main()
