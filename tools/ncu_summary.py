#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the few lines that matter for this path.
usage: python tools/ncu_summary.py <rep> [<rep> ...] > profiles/xxx.txt"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sectors_srcunit_tex_op_red.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_src_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]


def summarize(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("=" * 100)
        print(rep)
        print("kernel:", d.get("Kernel Name", ("", "?"))[1][:150])
        for k in KEEP:
            if k in d:
                print(f"  {k:85s} {d[k][1]:>16s} {d[k][0]}")
        stalls = []
        for k, (u, v) in d.items():
            if "average_warp_latency_issue_stalled" in k or ("warp_issue_stalled" in k and k.endswith("per_warp_active.pct")):
                try:
                    stalls.append((float(v), k))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  top warp stall reasons:")
        for v, k in stalls[:8]:
            print(f"    {v:10.3f}  {k}")


if __name__ == "__main__":
    for r in sys.argv[1:]:
        summarize(r)
