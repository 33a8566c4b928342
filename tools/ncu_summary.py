"""Summarise an `ncu --set full` report: one row per profiled launch with the counters DESIGN.md cites.

usage: python tools/ncu_summary.py gpurun_out/x/prof.ncu-rep > profiles/rNN_ncu_full_<name>.summary.csv
"""
import csv
import io
import subprocess
import sys

COLS = [
    "ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__inst_executed.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = [head.index(c) if c in head else -1 for c in COLS]
    w = csv.writer(sys.stdout)
    w.writerow(COLS)
    w.writerow([units[i] if i >= 0 else "" for i in idx])
    for r in body:
        w.writerow([r[i] if i >= 0 else "" for i in idx])


if __name__ == "__main__":
    main(sys.argv[1])
