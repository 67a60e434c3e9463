"""Summarise an ncu launch list (csv from --metrics gpu__time_duration.sum): per kernel count, total and mean duration.  usage: launch_summary.py file.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r and "Metric Value" in r:
        hdr = i; break
if hdr is None:
    print("no header found"); sys.exit(1)
h = rows[hdr]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    name = re.sub(r"\(.*", "", r[kn]); name = re.sub(r"^void ", "", name)
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    u = r[mu]
    v_us = v / 1e3 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1e3 if u in ("msecond", "ms") else v)
    c, t = acc.get(name, (0, 0.0)); acc[name] = (c + 1, t + v_us)
tot = sum(t for _, t in acc.values())
print("%-90s %7s %12s %10s %6s" % ("kernel", "count", "total_us", "mean_us", "share"))
for name, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-90s %7d %12.1f %10.2f %5.1f%%" % (name[:90], c, t, t / c, 100 * t / tot))
print("total %.1f us" % tot)
