#!/usr/bin/env python
"""Summarise .ncu-rep captures (gpurun_out/) into the text files kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/r2ncu_c4.ncu-rep [more.ncu-rep ...] > profiles/r2_ncu_direct.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print("# %s: no data" % path)
            continue
        hdr, units = rows[0], rows[1]
        print("# %s" % path.split("/")[-1])
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print("--- %s" % d.get("Kernel Name", "?")[:110])
            for m in METRICS:
                if m in d:
                    print("%s [%s] = %s" % (m, units[hdr.index(m)], d[m]))


if __name__ == "__main__":
    main()
