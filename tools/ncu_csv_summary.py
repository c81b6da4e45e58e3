#!/usr/bin/env python
"""Summarise `ncu -i report --page raw --csv` output (kept on the GPU box's scratch; only the CSV travels) into the
per-kernel text blocks kept under profiles/.

    python tools/ncu_csv_summary.py gpurun_out/r2e_ncu_v0_raw.csv [kernel-name-substring ...] > profiles/r02_..._summary.txt
"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2:]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [k for k in hdr if "issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k]
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if want and not any(w in name for w in want):
            continue
        print("----")
        print(f"  Kernel Name    {name[:150]}")
        print(f"  Grid / Block   {r[col['launch__grid_size']]} x {r[col['launch__block_size']]}")
        for k in KEYS:
            if k in col:
                print(f"  {k:<88} {r[col[k]]} {rows[1][col[k]]}")
        for k in stalls:
            try:
                if abs(float(r[col[k]])) < 0.25:
                    continue
            except ValueError:
                continue
            print(f"  {k:<88} {r[col[k]]}")


if __name__ == "__main__":
    main()
