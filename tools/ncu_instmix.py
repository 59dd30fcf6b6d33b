"""Instruction mix per kernel from an .ncu-rep captured with --import-source on (read here, no GPU needed):
python tools/ncu_instmix.py file.ncu-rep [out.txt]
Executed warp instructions per opcode, normalised per shift-register step (scan_regtab: FADD2 / (WB-1)) or per tile of
8 positions (scan_hmma: HMMA / (KS*NB))."""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
out = []
seen = set()
for b in blocks:
    sig = (b["name"], len(b["rows"]), tuple(b["rows"][1][:6]) if len(b["rows"]) > 1 else ())
    if sig in seen:  # the source page lists a launch once per view
        continue
    seen.add(sig)
    hdr = b["rows"][0]
    data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
    isrc, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    cnt, smp = Counter(), Counter()
    for r in data:
        toks = [t for t in r[isrc].split() if not t.startswith("@")]
        if not toks:
            continue
        op = toks[0].split(".")[0]
        cnt[op] += int(r[ie])
        smp[op] += int(r[isamp])
    tot, tots = sum(cnt.values()), sum(smp.values())
    out.append("kernel: " + b["name"])
    out.append("executed warp instructions: %d   stall samples: %d" % (tot, tots))
    per, unit = None, "-"
    m = re.search(r"scan_regtab_kernel<\(int\)(\d+)", b["name"])
    if m and cnt.get("FADD2"):
        per, unit = cnt["FADD2"] / (int(m.group(1)) - 1), "step"
    m = re.search(r"scan_hmma_kernel<\(int\)(\d+), \(int\)(\d+)", b["name"])
    if m and cnt.get("HMMA"):
        per, unit = cnt["HMMA"] / (int(m.group(1)) * int(m.group(2))), "tile"
    if per:
        out.append("warp %ss: %d   instructions per %s: %.2f" % (unit, per, unit, tot / per))
    out.append("%-10s %14s %10s %10s" % ("opcode", "executed", "per " + unit, "% samples"))
    for op, n in cnt.most_common(20):
        out.append("%-10s %14d %10s %9.1f%%" % (op, n, "%.2f" % (n / per) if per else "-", 100.0 * smp[op] / max(tots, 1)))
    out.append("")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
