"""profiles/r02_counters.json from .ncu-rep files (read on the CPU box):

    python scripts/ncu_counters.py <source_sha> <smem.ncu-rep> <replicas> <sweeps> [<hbm.ncu-rep> <n>]

smem report: launches of scripts/inst_count_smem.py (the LAST sweep_smem_kernel<float, extras> launch is the bench's step: amplitude
move + measurement); hbm report: scripts/profile_target.py hbm (first sweep_hbm_kernel<float,2,32,false> launch).  The hash is the
one `python -c "from cylinder_b200.build import source_sha; print(source_sha())"` printed in the SAME gpurun call."""
import csv, json, os, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]

def num(x):
    return float(x.replace(",", ""))

sha, smem_rep, replicas, sweeps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out = {"source_sha": sha}
ks = [r for r in raw(smem_rep) if "sweep_smem_kernel<float, 1" in r["Kernel Name"]]
k = ks[-1]
stalls = sorted(((num(v), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for n, v in k.items()
                 if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and v), reverse=True)[:6]
dur = num(k["gpu__time_duration.sum"]) * 1e-6           # base unit: ns
out["sweep_smem_kernel<float,extras>"] = {
    "replicas": replicas, "sweeps": sweeps, "flags": "ADAPT_SIGMA | AMP_MOVES | MEASURE", "report": os.path.basename(smem_rep),
    "smsp__inst_executed.sum": num(k["smsp__inst_executed.sum"]), "gpu__time_duration_ms": dur,
    "issue_active_pct": num(k["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
    "registers_per_thread": num(k["launch__registers_per_thread"]),
    "top_stalls": ["%s %.2f" % (n, v) for v, n in stalls]}
if len(sys.argv) > 6:
    hbm_rep, n = sys.argv[5], int(sys.argv[6])
    hs = [r for r in raw(hbm_rep) if "sweep_hbm_kernel<float, 2, 32, 0>" in r["Kernel Name"] or "sweep_hbm_kernel<float, (int)2, (int)32, (bool)0>" in r["Kernel Name"]]
    h = hs[len(hs) // 2]
    out["sweep_hbm_kernel<float,2,32,false>"] = {
        "n": n, "report": os.path.basename(hbm_rep), "launch": "sweep %d of %d profiled" % (len(hs) // 2, len(hs)),
        "dram__bytes_read.sum": num(h["dram__bytes_read.sum"]), "dram__bytes_write.sum": num(h["dram__bytes_write.sum"]),
        "gpu__time_duration_us": num(h["gpu__time_duration.sum"]) * 1e-3, "smsp__inst_executed.sum": num(h["smsp__inst_executed.sum"]),
        "issue_active_pct": num(h["smsp__issue_active.avg.pct_of_peak_sustained_active"])}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_counters.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
