#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small text file for profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "local_load", "local_store",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "sm__cycles_elapsed.max"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {rep}: selected metrics per profiled launch (ncu --set full --clock-control none)\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"\n== kernel {name}\n")
            for i, h in enumerate(hdr):
                if any(h == k or h.startswith(k) for k in KEYS):
                    f.write(f"{h:80s} {r[i]:>18s} {units[i]}\n")
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
