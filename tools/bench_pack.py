"""HBM roofline of the 2-bit packer: 248 Mb of ASCII already in HBM -> packed + mask (+ codes only where needed).
Algorithmic bytes per base: 1 read + 0.25 + 0.125 written."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smeagol_b200 import _lib  # noqa: E402
from smeagol_b200 import device as dev  # noqa: E402
from smeagol_b200 import synth  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 248_000_000
seq = synth.single_sequence(L, 4)
lib = _lib.load()
device = torch.device("cuda", 0)
rows = np.zeros(1, dtype=dev.ROW_DTYPE)
words = ((L + dev.ROW_PAD_BASES + 63) // 64) * 4
rows["g_len"], rows["n_avail"], rows["n_own"] = L, L, L
d_rows = dev._to_dev(rows, device)
d_src = torch.from_numpy(seq).to(device)
d_off = torch.zeros(1, dtype=torch.int64, device=device)
packed = torch.empty(words, dtype=torch.int32, device=device)
amask = torch.empty(words, dtype=torch.int16, device=device)
codes = torch.empty(words * 16, dtype=torch.uint8, device=device)
status = torch.tensor([-1, 0], dtype=torch.int64, device=device)
best = 1e9
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.smg_pack(dev._ptr(d_src), 0, dev._ptr(d_rows), dev._ptr(d_off), 1, words, dev._ptr(packed), dev._ptr(amask),
                            dev._ptr(codes), dev._ptr(status), dev._stream()), "smg_pack")
    e1.record()
    torch.cuda.synchronize()
    if i:
        best = min(best, e0.elapsed_time(e1))
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
gbs = L * 1.375 / best / 1e6
print(json.dumps({"kernel": "pack_kernel", "bases": L, "ms": best, "algorithmic_bytes": int(L * 1.375), "achieved_gbs": gbs,
                  "peak_gbs": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"], "gbases_per_s": L / best / 1e6}))
