"""Summarise an .ncu-rep (read on the CPU box): python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
out = []
seen = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    if "Function Name" in idx and "<" not in name:
        name = r[idx["Function Name"]]
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > 1:
        continue
    out.append("## %s" % name)
    for k in KEYS:
        if k in idx:
            out.append("- %s = %s %s" % (k, r[idx[k]], units[idx[k]]))
    st = []
    for h, i in idx.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    out.append("- warps stalled per issued instruction, by reason (smsp__average_warps_issue_stalled_*_per_issue_active): "
               + ", ".join("%s %.2f" % (h, v) for v, h in sorted(st, reverse=True)[:8]))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("# ncu summary of %s\n\n" % rep + text + "\n")
