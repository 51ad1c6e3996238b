"""Static view of the hot loops of a kernel: every backward branch of the function whose body contains a marker opcode
(default MUFU.SIN = a site move), with the body's instruction count and opcode histogram.

    python scripts/sass_loops.py cylinder_b200/build/k_sweep_smem.o IfLb1ELb0E [MARKER]
"""
import os, re, subprocess, sys, tempfile
from collections import Counter
obj, pat = sys.argv[1], sys.argv[2]
marker = sys.argv[3] if len(sys.argv) > 3 else "MUFU.SIN"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
fn, cur = {}, None
for l in dis:
    if l.startswith(".text."):
        cur = l[6:].rstrip(":"); fn[cur] = []
    elif cur is not None:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", l)
        if m:
            fn[cur].append(("i", m.group(2).strip()))
        else:
            m = re.match(r"(\.L_x_\d+):", l)
            if m:
                fn[cur].append(("l", m.group(1)))
for name, items in fn.items():
    if pat not in name:
        continue
    ins, labels = [], {}
    for kind, v in items:
        if kind == "l":
            labels[v] = len(ins)
        else:
            ins.append(v)
    print("%s: %d instructions" % (name[:60], len(ins)))
    for i, s in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s+)*`\((\.L_x_\d+)\)", s)
        if m and m.group(1) in labels and labels[m.group(1)] <= i:
            lo = labels[m.group(1)]
            body = ins[lo:i + 1]
            if any(marker in b for b in body) and len(body) < 400:
                ops = Counter(re.sub(r"^@!?U?P\d+\s+", "", b).split()[0].split(".")[0] for b in body)
                print("  loop %5d..%5d: %3d instructions  %s" % (lo, i, len(body), " ".join("%s:%d" % kv for kv in ops.most_common(14))))
