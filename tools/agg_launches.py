"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and grid."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        if unit in ("nsecond", "ns"):
            v /= 1000
        elif unit in ("msecond", "ms"):
            v *= 1000
        key = (name, row.get("Grid Size", ""))
        agg[key][0] += 1
        agg[key][1] += v
        tot += v
    print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d} avg={t / n:9.2f} us  {k[0][:64]} grid={k[1]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
