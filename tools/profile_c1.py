"""Host-side profile of the small-input path (config c1): where do the milliseconds of one scan_sequences() call go?"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smeagol_b200 as sb  # noqa: E402
from smeagol_b200 import synth  # noqa: E402

pw = synth.synth_pwms(100, 7, 12)
model = sb.models.PWMModel(pw)
seq = synth.single_sequence(30_000, 1)
rec = [sb.SeqRecord(sb.Seq(seq.tobytes().decode()), id="virus", name="virus")]
for _ in range(3):
    out = sb.scan.scan_sequences(rec, model, 0.7, sense="+", rcomp="both", outputs=["sites", "counts"], score=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    out = sb.scan.scan_sequences(rec, model, 0.7, sense="+", rcomp="both", outputs=["sites", "counts"], score=True)
torch.cuda.synchronize()
print("ms per call", (time.perf_counter() - t0) / 20 * 1e3, "sites", len(out["sites"]))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    out = sb.scan.scan_sequences(rec, model, 0.7, sense="+", rcomp="both", outputs=["sites", "counts"], score=True)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
