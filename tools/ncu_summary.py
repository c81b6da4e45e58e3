#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into the per-kernel text blocks kept under profiles/.

    python tools/ncu_summary.py gpurun_out/pat_tma_full.ncu-rep > profiles/rXX_ncu_full_<kernel>_summary.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def find(name):
        if name in col:
            return col[name]
        for h, i in col.items():
            if h.endswith("." + name) or h.endswith(name):
                return i
        return None
    for r in rows[2:]:
        print("----")
        print(f"  {'Kernel Name':<90} {r[col['Kernel Name']][:100]}")
        print(f"  {'Grid Size':<90} {r[col['Grid Size']]}")
        print(f"  {'Block Size':<90} {r[col['Block Size']]}")
        for m in METRICS:
            i = find(m)
            if i is None or r[i] == "":
                continue
            print(f"  {m:<90} {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
