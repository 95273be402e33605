"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: count / total time / share per kernel.
usage: launch_summary.py launches.csv > summary.txt"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "")[:72]
    tot[name] += us; cnt[name] += 1
all_us = sum(tot.values())
print("%-72s %7s %12s %7s" % ("kernel", "count", "total us", "share"))
for k, v in tot.most_common(40):
    print("%-72s %7d %12.1f %6.1f%%" % (k, cnt[k], v, 100 * v / all_us))
step = {k: v for k, v in tot.items() if any(s in k for s in ("harm_hw", "wn_lane", "fund_tile"))}
if step:
    s = sum(step.values())
    print("\nwithin the headline step (hpf_solve, structured strategy: fund_tile -> wn_lane -> harm_hw):")
    for k, v in sorted(step.items(), key=lambda kv: -kv[1]):
        print("  %-60s %5.1f%% of the step's kernel time (%d launches)" % (k, 100 * v / s, cnt[k]))
