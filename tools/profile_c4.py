"""Driver for ncu captures of the hybrid engine: a slice of config c4 (JASPAR-sized PWMs, W 6-24) on one sequence."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smeagol_b200 import device as dev  # noqa: E402
from smeagol_b200 import synth  # noqa: E402
from smeagol_b200.models import PWMModel  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
M = 1000
model = PWMModel(synth.synth_pwms(M, widths=synth.jaspar_widths(M, seed=4)))
seq = synth.single_sequence(L, 4)
batch = dev.SeqBatch([seq])
out_rows = np.array([[0, 1]], dtype=np.int32)
plan = dev.ScanPlan(model.device_set, batch, model.thresholds(0.7), 3, out_rows, 2)
for _ in range(reps):
    plan.launch()
    res = plan.finish(sort=False)
torch.cuda.synchronize()
print("hits", res.n_hits, res.stats)
