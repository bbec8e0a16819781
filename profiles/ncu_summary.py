"""Print the key counters of an .ncu-rep (read here, no GPU needed): python profiles/ncu_summary.py <rep> [kernel-row]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    for k in KEYS:
        if k in d:
            print(f"{k:90s} {d[k]:>18s} {units[hdr.index(k)]}")
    stalls = sorted(((float(v.replace(',', '')), h) for h, v in d.items()
                     if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and v),
                    reverse=True)
    for v, h in stalls[:8]:
        print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:8.3f} warps/issue")
