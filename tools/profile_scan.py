"""Minimal driver for ncu captures: build the c2 workload (or a smaller one) and run the fused scan a few times."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smeagol_b200 import device as dev  # noqa: E402
from smeagol_b200 import synth  # noqa: E402
from smeagol_b200.models import PWMModel  # noqa: E402

n_genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pw = synth.synth_pwms(1000, 7, 12)
genomes = synth.genome_set(1000, 10_000_000, 5000, 15000, seed=2)[:n_genomes]
model = PWMModel(pw)
n = len(genomes)
batch = dev.SeqBatch(genomes)
out_rows = np.stack([2 * np.arange(n), 2 * np.arange(n) + 1], axis=1).astype(np.int32)
plan = dev.ScanPlan(model.device_set, batch, model.thresholds(0.7), 3, out_rows, 2 * n)
for _ in range(reps):
    plan.launch()
    res = plan.finish(sort=True)
torch.cuda.synchronize()
print("hits", res.n_hits, res.stats)
