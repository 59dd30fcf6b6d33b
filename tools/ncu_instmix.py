"""Instruction mix of one kernel from an .ncu-rep captured with --import-source on (read here, no GPU needed):
python tools/ncu_instmix.py file.ncu-rep [out.txt]   -- executed warp instructions per opcode, per FADD2-normalised step"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kernel, hdr, data = rows[0][1], rows[1], rows[2:]
isrc, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
cnt, smp = Counter(), Counter()
for r in data:
    toks = [t for t in r[isrc].split() if not t.startswith("@")]
    op = toks[0].split(".")[0]
    cnt[op] += int(r[ie])
    smp[op] += int(r[isamp])
tot, tots = sum(cnt.values()), sum(smp.values())
out = ["kernel: " + kernel, "executed warp instructions: %d   stall samples: %d" % (tot, tots)]
per = None
if cnt.get("FADD2"):
    # a shift-register step of width WB issues WB-1 FADD2 (the first element is a register copy)
    import re
    m = re.search(r"scan_regtab_kernel<\(int\)(\d+)", kernel)
    if m:
        per = cnt["FADD2"] / (int(m.group(1)) - 1)
        out.append("warp steps (FADD2 / (WB-1)): %d   instructions per step: %.2f" % (per, tot / per))
out.append("%-10s %14s %10s %10s" % ("opcode", "executed", "per step", "% samples"))
for op, n in cnt.most_common(24):
    out.append("%-10s %14d %10s %9.1f%%" % (op, n, "%.2f" % (n / per) if per else "-", 100.0 * smp[op] / max(tots, 1)))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
