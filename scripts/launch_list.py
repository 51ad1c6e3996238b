"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table: python scripts/launch_list.py in.csv out.md "command" """
import collections
import csv
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.defaultdict(float), collections.defaultdict(int)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0][:90]
    tot[name] += v / 1e3
    cnt[name] += 1
total = sum(tot.values())
out = ["# ncu launch list of `%s`" % cmd, "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` (per-launch times are cold-cache and serialised: "
       "compare shares). Raw list: `%s`." % src.split("/")[-1], "", "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    out.append("| `%s` | %d | %.1f | %.1f %% |" % (name, cnt[name], t, 100 * t / total))
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
