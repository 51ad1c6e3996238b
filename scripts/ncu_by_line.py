"""Warp-instructions executed and stall samples per SOURCE LINE of one kernel of an .ncu-rep (read on the CPU box).

    python scripts/ncu_by_line.py gpurun_out/prof.ncu-rep cylinder_b200/build/k_sweep_smem.o 'sweep_smem_kernel<float, (bool)1' [launch#] [top]

The report's source page lists SASS only; the line table comes from `nvdisasm -g` of the same object (the build is
deterministic: the instruction sequences are compared opcode by opcode before anything is attributed)."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, obj, pat = sys.argv[1], sys.argv[2], sys.argv[3]
nth = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
kernels, cur = [], None
for row in csv.reader(raw):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        kernels.append(cur)
    elif row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
sel = [k for k in kernels if pat in k["name"]]
k = sel[nth]
h = {n: i for i, n in enumerate(k["hdr"])}

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()


def mangled_matches(sym, name):
    # compare template arguments loosely: type letter and the two bools
    m = re.search(r"sweep_\w+_kernel|[a-z_]+_kernel", name)
    base = m.group(0) if m else name.split("<")[0].split()[-1]
    if base not in sym:
        return False
    t = "f" if "<float" in name else ("d" if "<double" in name else "")
    bools = re.findall(r"\(bool\)(\d)", name)
    ints = re.findall(r"\(int\)(\d+)", name)
    want = "I" + t + "".join("Lb%sE" % b for b in bools) if bools and not ints else None
    return want is None or want in sym


funcs, curf, line = {}, None, None
for l in dis:
    if l.startswith(".text."):
        curf = l[6:].rstrip(":")
        funcs[curf] = []
        line = None
    elif "//## File" in l:
        m = re.search(r'File "([^"]+)", line (\d+)', l)
        line = (os.path.basename(m.group(1)), int(m.group(2)))
    elif curf and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        ins = re.sub(r"/\*[0-9a-f]+\*/", "", l).strip().rstrip(";").strip()
        funcs[curf].append((line, ins))
cands = [f for f in funcs if mangled_matches(f, k["name"]) and len(funcs[f]) == len(k["rows"])]
assert cands, "no function of %s has %d instructions" % (obj, len(k["rows"]))
fn = funcs[cands[0]]
for (ln, ins), row in zip(fn, k["rows"]):
    a, b = ins.split()[0].lstrip("@!PU0123456789 "), row[h["Source"]].split()
    op = [w for w in ins.split() if not w.startswith("@")][0]
    op2 = [w for w in b if not w.startswith("@")][0]
    assert op == op2, (ins, row[h["Source"]])

by = defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0]
for (ln, ins), row in zip(fn, k["rows"]):
    ie, te, smp = int(row[h["Instructions Executed"]]), int(row[h["Thread Instructions Executed"]]), int(row[h["# Samples"]])
    e = by[ln]
    e[0] += ie; e[1] += te; e[2] += smp; e[3] += 1
    tot[0] += ie; tot[1] += te; tot[2] += smp
print("kernel %s\n%d SASS instructions, %.4g warp-instructions executed, %.4g thread-instructions (%.1f active lanes), %d samples"
      % (k["name"][:90], len(fn), tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2]))
print("%-22s %6s %8s %8s %6s %8s" % ("file:line", "#sass", "warp-inst", "share", "lanes", "samples"))
for ln, e in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %6d %8.3g %7.2f%% %6.1f %7.2f%%" % ("%s:%d" % ln if ln else "?", e[3], e[0], 100.0 * e[0] / tot[0], e[1] / max(e[0], 1),
                                                  100.0 * e[2] / max(tot[2], 1)))
# by file
byf = defaultdict(lambda: [0, 0])
for ln, e in by.items():
    byf[ln[0] if ln else "?"][0] += e[0]
    byf[ln[0] if ln else "?"][1] += e[2]
print()
for f, e in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print("%-22s %8.3g %7.2f%%  samples %6.2f%%" % (f, e[0], 100.0 * e[0] / tot[0], 100.0 * e[1] / max(tot[2], 1)))
# by line ranges of the main file (phases), if given as env CYL_RANGES="name:lo-hi,..."
rng = os.environ.get("CYL_RANGES")
if rng:
    print()
    main = os.path.basename(obj).replace(".o", ".cu")
    for part in rng.split(","):
        name, r = part.split(":")
        lo, hi = map(int, r.split("-"))
        s = sum(e[0] for ln, e in by.items() if ln and ln[0] == main and lo <= ln[1] <= hi)
        sm = sum(e[2] for ln, e in by.items() if ln and ln[0] == main and lo <= ln[1] <= hi)
        print("%-16s lines %4d-%4d %8.3g %7.2f%%  samples %6.2f%%" % (name, lo, hi, s, 100.0 * s / tot[0], 100.0 * sm / max(tot[2], 1)))

# straight-line stretches (equal execution count) in address order: the loop-level view
if os.environ.get("CYL_BLOCKS"):
    print()
    minshare = float(os.environ.get("CYL_BLOCKS"))
    i, n = 0, len(fn)
    while i < n:
        ie = int(k["rows"][i][h["Instructions Executed"]])
        j, te, smp, lines = i, 0, 0, defaultdict(int)
        while j < n and int(k["rows"][j][h["Instructions Executed"]]) == ie:
            te += int(k["rows"][j][h["Thread Instructions Executed"]]); smp += int(k["rows"][j][h["# Samples"]])
            lines[fn[j][0]] += 1
            j += 1
        w = ie * (j - i)
        if 100.0 * w / tot[0] >= minshare:
            tl = sorted(lines.items(), key=lambda kv: -kv[1])[:6]
            print("sass %5d..%5d  n=%4d exec=%9.3g warp-inst=%9.3g %6.2f%% lanes %4.1f samples %5.2f%%  %s"
                  % (i, j - 1, j - i, ie, w, 100.0 * w / tot[0], te / max(w, 1), 100.0 * smp / max(tot[2], 1),
                     " ".join("%s:%d(%d)" % (a[0].replace("k_sweep_", "").replace("cyl_", ""), a[1], c) if a else "?" for a, c in tl)))
        i = j

# CYL_DUMP="lo-hi": the SASS of that index range with its source lines
if os.environ.get("CYL_DUMP"):
    lo, hi = map(int, os.environ["CYL_DUMP"].split("-"))
    for i in range(lo, hi + 1):
        ln, ins = fn[i]
        print("%5d %-18s %s" % (i, "%s:%d" % (ln[0].replace("k_sweep_", "").replace("cyl_", ""), ln[1]) if ln else "?", ins))

# CYL_CLASSES=1: warp-instructions grouped by execution count (which loop level a stretch belongs to)
if os.environ.get("CYL_CLASSES"):
    cl = defaultdict(lambda: [0, 0])
    for row in k["rows"]:
        ie = int(row[h["Instructions Executed"]])
        cl[ie][0] += 1; cl[ie][1] += ie
    print()
    for ie, (n, w) in sorted(cl.items(), key=lambda kv: -kv[1][1])[:40]:
        print("exec=%10.4g  n=%5d  warp-inst=%9.3g %6.2f%%" % (ie, n, w, 100.0 * w / tot[0]))

# CYL_EXEC="count": dump every instruction executed exactly `count` times (e.g. once per warp per sweep), with source lines
if os.environ.get("CYL_EXEC"):
    want = float(os.environ["CYL_EXEC"])
    print()
    for i, ((ln, ins), row) in enumerate(zip(fn, k["rows"])):
        ie = int(row[h["Instructions Executed"]])
        if abs(ie - want) <= 0.002 * want:
            print("%5d %-18s %s" % (i, "%s:%d" % (ln[0].replace("k_sweep_", "").replace("cyl_", ""), ln[1]) if ln else "?", ins))
