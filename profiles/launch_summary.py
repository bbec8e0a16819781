"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
    python profiles/launch_summary.py gpurun_out/r01_launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    if "<" in name and not name.startswith("mga::"):
        name = name.split("<")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:8d} {us:12.1f} {us / n:10.1f} {100 * us / tot:6.1f}%")
print(f"{'total':70s} {len(rows):8d} {tot:12.1f}")
