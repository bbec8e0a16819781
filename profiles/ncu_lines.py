"""Per-source-line stall samples / shared-memory wavefronts of one kernel in an .ncu-rep (source page, needs -lineinfo):
    python profiles/ncu_lines.py <rep> [top_n] [kernel-substring]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Function Name":
        cur["fn"] = r[1]
    elif r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and r and r[0].isdigit():
        cur["rows"].append(r)
seen = set()
def num(v):
    try:
        return int(v.replace(",", ""))
    except ValueError:
        return 0


for b in blocks:
    if want not in b.get("fn", ""):
        continue
    h = b["hdr"]
    ci = {n: h.index(n) for n in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal") if n in h}
    tot = sum(num(r[ci["# Samples"]]) for r in b["rows"])
    key = (b["fn"], b["file"])
    if tot == 0 or key in seen:
        continue
    seen.add(key)
    print(f"== {b['fn'][:90]} | {b['file'].split('/')[-1]} | samples {tot}")
    rs = sorted(b["rows"], key=lambda r: -num(r[ci["# Samples"]]))[:top]
    for r in rs:
        print(f"{num(r[ci['# Samples']]):7d} {100.0 * num(r[ci['# Samples']]) / tot:5.1f}%  inst {r[ci['Instructions Executed']]:>9s}  wf {r[ci['L1 Wavefronts Shared']]:>9s}/{r[ci['L1 Wavefronts Shared Ideal']]:>9s}  L{r[0]:>4s}: {r[1].strip()[:110]}")
