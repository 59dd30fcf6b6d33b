"""Per-width scan time of every engine next to the default hybrid: N PWMs of ONE width x 4 Mb of
random ACGT, both strands, threshold 0.7, hits + counts, scan kernels only (tools/bench_widths.py in a subprocess per
engine so that the environment switches are read fresh).

    python tools/bench_widths_matrix.py [widths] [n_pwms,n_pwms,...]  >  profiles/r02_bench_widths.txt

Rows: ms per scan (lower is better).  'hybrid' is what a caller gets with no switches set; the last column is
hybrid / best single engine (<= 1.05 means the split rule picked the fastest engine at that width)."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
widths = sys.argv[1] if len(sys.argv) > 1 else "6,7,8,9,10,11,12,13,14,16,18,20,24"
sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "256,1024").split(",")]
ENGINES = [
    ("regtab", {"SMEAGOL_B200_ENGINE": "regtab", "SMEAGOL_B200_TC_MIN_WIDTH": "0"}),
    ("hmma", {"SMEAGOL_B200_ENGINE": "hmma", "SMEAGOL_B200_TC_MIN_WIDTH": "0"}),
    ("tc5", {"SMEAGOL_B200_ENGINE": "tc5", "SMEAGOL_B200_TC_MIN_WIDTH": "0"}),
    ("kmer", {"SMEAGOL_B200_KMER_MAX_WIDTH": "12"}),  # the k-mer index for every width it supports, tcgen05 above
    ("hybrid", {}),
]


def run(env_extra, nm):
    env = dict(os.environ)
    for k in ("SMEAGOL_B200_ENGINE", "SMEAGOL_B200_TC_MIN_WIDTH", "SMEAGOL_B200_KMER_MAX_WIDTH", "SMEAGOL_B200_KMER",
              "SMEAGOL_B200_TC_KIND"):
        env.pop(k, None)
    env.update(env_extra)
    out = subprocess.run([sys.executable, os.path.join(HERE, "bench_widths.py"), widths, "0.7", str(nm)], env=env,
                         capture_output=True, text=True)
    res = {}
    for line in out.stdout.splitlines():
        m = re.match(r"W=\s*(\d+).*?([\d.]+) ms", line)
        if m:
            res[int(m.group(1))] = float(m.group(2))
    if not res:
        sys.stderr.write(out.stdout[-2000:] + out.stderr[-2000:])
    return res


for nm in sizes:
    table = {name: run(env, nm) for name, env in ENGINES}
    ws = sorted(table["hybrid"])
    print("%d PWMs of one width x 4,000,000 bases x 2 strands, thr 0.7 -- ms per scan" % nm)
    print("%-8s" % "W" + "".join("%9d" % w for w in ws))
    for name, _ in ENGINES:
        print("%-8s" % name + "".join(("%9.3f" % table[name][w]) if w in table[name] else "%9s" % "-" for w in ws))
    ratio = []
    for w in ws:
        best = min(table[n][w] for n, _ in ENGINES[:-1] if w in table[n])
        ratio.append(table["hybrid"][w] / best)
    print("%-8s" % "hyb/best" + "".join("%9.2f" % r for r in ratio))
    print()
