"""profiles/<tag>_resident_counters.json (read by bench.py for roofline.traffic / the shared-memory roofline) from an ncu
report:    python profiles/refresh_counters.py gpurun_out/<rep>.ncu-rep <batch> [out.json]"""
import csv
import json
import subprocess
import sys

rep, batch = sys.argv[1], int(sys.argv[2])
dest = sys.argv[3] if len(sys.argv) > 3 else "profiles/r02_resident_counters.json"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
d, u = dict(zip(hdr, rows[2])), dict(zip(hdr, units))


def val(k):
    return float(d[k].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[k], 1)


j = {"source": f"{rep} (ncu --set full --clock-control none, profiles/profile_step.py "
               f"--mode resident --batch {batch}, 2nd launch)",
     "kernel": d["Kernel Name"], "batch": batch,
     "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
     "smem_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
     "smem_bank_conflicts": val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
     "smem_pipe_pct_of_peak": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
     "warp_instructions": val("smsp__inst_executed.sum"),
     "duration_us": val("gpu__time_duration.sum") * (1e3 if u["gpu__time_duration.sum"] == "ms" else 1)}
json.dump(j, open(dest, "w"), indent=1)
print(j)
