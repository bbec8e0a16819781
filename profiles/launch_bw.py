"""Per-kernel time and DRAM / L2 traffic from an ncu CSV with gpu__time_duration.sum, dram__bytes_*.sum, lts__t_bytes.sum:
    python profiles/launch_bw.py <csv> [skip_first_n_launches]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
per = collections.OrderedDict()
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
launch = {}
for r in rows:
    lid = int(r[0])
    if lid < skip:
        continue
    name = r[4].split("(")[0].replace("void ", "")
    if "<" in name and not name.startswith("mga::"):
        name = name.split("<")[0]
    d = launch.setdefault(lid, {"name": name})
    d[r[12]] = float(r[14].replace(",", "")) * UNIT.get(r[13], 1)
for d in launch.values():
    a = per.setdefault(d["name"], collections.Counter())
    a["n"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0)
    a["dram"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    a["l2"] += d.get("lts__t_bytes.sum", 0)
tot = sum(a["us"] for a in per.values())
print(f"{'kernel':48s} {'n':>5s} {'total us':>10s} {'avg us':>9s} {'share':>6s} {'DRAM GB/s':>10s} {'L2 GB/s':>9s} {'DRAM MB/launch':>14s}")
for k, a in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{k[:48]:48s} {a['n']:5d} {a['us']:10.1f} {a['us'] / a['n']:9.1f} {100 * a['us'] / tot:5.1f}% "
          f"{a['dram'] / a['us'] / 1e3:10.0f} {a['l2'] / a['us'] / 1e3:9.0f} {a['dram'] / a['n'] / 1e6:14.1f}")
print(f"{'total':48s} {sum(a['n'] for a in per.values()):5d} {tot:10.1f}")
