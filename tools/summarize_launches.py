"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel/grid."""
import collections
import csv
import re
import sys

for path in sys.argv[1:]:
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        a = agg.setdefault((k, r["Grid Size"], r["Block Size"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{path}: {sum(a[0] for a in agg.values())} launches, total {tot:.1f} us")
    for k, a in agg.items():
        print(f"  {k[0][:52]:52s} grid {k[1]:>14s} blk {k[2]:>11s} n={a[0]:4d} avg {a[1]/a[0]:8.2f} us sum {a[1]:9.1f} ({100*a[1]/tot:4.1f}%)")
