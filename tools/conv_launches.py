"""Per-launch time and DRAM traffic of the conv-stack kernels from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` log (tools/gpu scripts).  python tools/conv_launches.py launches.csv"""
import collections
import csv
import re
import sys

UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for r in csv.DictReader(rows):
    k = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]), r["Grid Size"])
    agg.setdefault(k, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
tot = 0.0
for k, m in agg.items():
    t, rd, wr = m["gpu__time_duration.sum"], m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
    us = t[0] / 1000 if t[1] in ("ns", "nsecond") else t[0]
    mb = rd[0] * UNIT.get(rd[1], 1) + wr[0] * UNIT.get(wr[1], 1)
    tot += us
    print(f"{k[1]:34s} {k[2]:16s} {us:9.1f} us  DRAM {mb:8.1f} MB  {mb / us * 1e-3:6.2f} TB/s")
print(f"total {tot / 1e3:.2f} ms")
