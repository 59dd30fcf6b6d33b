"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv):
python tools/ncu_launch_summary.py X.csv "<command that was profiled>" [out.txt]"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik])[:70]
    v = float(r[iv].replace(",", ""))
    ms = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu] in ("us", "usecond") else v
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + ms)
total = sum(t for _, t in agg.values())
out = ["ncu launch list of `%s` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)" % sys.argv[2],
       "%d launches captured, total %.3f ms" % (sum(c for c, _ in agg.values()), total),
       "%-72s %6s %10s %7s" % ("kernel", "count", "ms", "share")]
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-72s %6d %10.3f %6.1f%%" % (name, c, t, 100 * t / total))
for key in ("scan_regtab", "scan_hmma", "finalize_candidates"):
    s = sum(t for n, (_, t) in agg.items() if key in n)
    if s:
        out.append("%s share of all captured GPU time: %.1f%%" % (key, 100 * s / total))
text = "\n".join(out)
print(text)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(text + "\n")
