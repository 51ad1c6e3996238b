"""Opcode histogram of every kernel of the shared library's objects (static SASS): python scripts/sass_summary.py > profiles/r02_sass_summary.md
The opcodes that prove the Blackwell/Hopper data-movement and packed-math paths are listed first (UTMALDG = TMA tensor load,
UBLKCP = bulk copy, SYNCS = mbarrier, FFMA2/FMUL2/FADD2 = packed fp32)."""
import os, re, subprocess, sys, tempfile
from collections import Counter
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objdir = os.path.join(HERE, "cylinder_b200", "build")
KEY = ["UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU", "IMAD", "LOP3", "FFMA", "DFMA", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "REDUX", "ATOMS", "RED", "LDL", "STL"]
print("# SASS opcode summary of the sm_100a kernels (static counts per kernel)\n")
print("`cuobjdump -sass` of `cylinder_b200/build/*.o`, kernel sources %s.  No `tcgen05`/`UTCMMA` anywhere: the path has no contraction.\n" %
      subprocess.run([sys.executable, "-c", "from cylinder_b200.build import source_sha; print(source_sha())"], cwd=HERE, capture_output=True, text=True).stdout.strip())
print("| object | kernel | instructions | " + " | ".join(KEY) + " |")
print("|---|---|---:|" + "---:|" * len(KEY))
for o in sorted(os.listdir(objdir)):
    if not o.endswith(".o"):
        continue
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(objdir, o)], capture_output=True, text=True).stdout
    cur, ops = None, {}
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1); ops[cur] = Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", l)
        if m and cur:
            ops[cur][m.group(1)] += 1
    for k, c in ops.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        if sum(c.values()) < 40:
            continue
        print("| %s | `%s` | %d | %s |" % (o, name[:70], sum(c.values()), " | ".join(str(c.get(x, 0)) for x in KEY)))
