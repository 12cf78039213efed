#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export (one profiled launch) into the handful of metrics DESIGN.md argues from.
usage: python tools/ncu_select.py <raw.csv> <out.txt> "<what was profiled>" """
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(raw, out, what):
    rows = list(csv.reader(open(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    with open(out, "w") as o:
        o.write(f"# selected raw metrics of ONE launch (ncu --set full --clock-control none): {what}\n")
        for i, h in enumerate(hdr):
            if h in KEYS or h.startswith("smsp__average_warps_issue_stalled"):
                o.write(f"{h:95s} {r[i]:>20s} {units[i]}\n")


if __name__ == "__main__":
    main(*sys.argv[1:4])
