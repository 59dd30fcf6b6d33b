"""Per-width throughput of the fused scan kernel: 256 PWMs of one width x 4 Mb random ACGT, both strands.
Prints T lookup-add/s (algorithmic: 2*W per base x motif) and the fraction of the measured FP32 add peak."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smeagol_b200 import _lib  # noqa: E402
from smeagol_b200 import device as dev  # noqa: E402
from smeagol_b200 import synth  # noqa: E402
from smeagol_b200.models import PWMModel  # noqa: E402

widths = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(6, 25))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
nm = int(sys.argv[3]) if len(sys.argv) > 3 else 256
L = 4_000_000
seq = synth.single_sequence(L, 11)
batch = dev.SeqBatch([seq])
out_rows = np.array([[0, 1]], dtype=np.int32)
for W in widths:
    pw = synth.synth_pwms(nm, widths=[W] * nm)
    model = PWMModel(pw)
    plan = dev.ScanPlan(model.device_set, batch, model.thresholds(frac), 3, out_rows, 2, engine=os.environ.get("SMEAGOL_B200_ENGINE"))
    plan.launch()
    res = plan.finish(sort=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        plan.counters.zero_()
        plan.counts.zero_()
        e0.record()
        plan.launch_kmer_build()  # the per-call k-mer index (no-op for the other engines) is part of the scan
        plan.launch_scan()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    lookups = L * 2.0 * W * nm
    print("W=%2d groups=%d buckets=%d  %7.3f ms  %6.2f T lookup-add/s  %5.1f%% of FP32 peak  %6.1f Gbase*motif/s  hit density %.2e" % (
        W, model.device_set.n_groups, model.device_set.n_buckets,
        best, lookups / best / 1e9, lookups / best / 1e9 / 36.9 * 100, L * nm / best / 1e6,
        res.stats["n_hits_total"] / (2.0 * L * nm)), flush=True)
    model.device_set.close()
