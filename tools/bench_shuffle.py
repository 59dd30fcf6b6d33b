"""Background generation for config c3: 1,000 dinucleotide-preserving shuffles of a 30 kb genome.

    python tools/bench_shuffle.py [out.json]

Device: smg_kmer_shuffle (all copies in one launch) timed with CUDA events, and shuffle + pack into a SeqBatch end to
end (wall clock).  CPU: the restatement of deeplift.dinuc_shuffle's algorithm (oracle/shuffle_oracle.py) on a bounded
sample of copies, one core, extrapolated.  SURVEY section 6 quotes 11.1 s for 250 shuffles in the reference."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smeagol_b200 as sb  # noqa: E402
from oracle import shuffle_oracle as sho  # noqa: E402
from smeagol_b200 import synth  # noqa: E402
from smeagol_b200 import utils as su  # noqa: E402

genome = synth.single_sequence(30_000, 3).tobytes().decode()
rec = [sb.SeqRecord(sb.Seq(genome), id="virus", name="virus")]
out = {"workload": "1,000 shuffles of a 30,000-base genome (config c3's backgrounds)", "runs": []}
for k in (2, 1):
    for _ in range(2):
        su.shuffle_batch_device(rec, 1000, k, seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ids, names, batch, _ = su.shuffle_batch_device(rec, 1000, k, seed=1)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    # kernel alone
    from smeagol_b200 import _lib
    from smeagol_b200 import device as dev
    from smeagol_b200.encode import integer_encoding_dict
    lut = np.full(256, 255, dtype=np.uint8)
    for ch, code in integer_encoding_dict.items():
        lut[ord(ch)] = code
    tok = lut[np.frombuffer(genome.encode(), dtype=np.uint8)]
    succ, off = su._successor_buckets(tok)
    d_tok, d_succ, d_off = (torch.from_numpy(x).cuda() for x in (tok.copy(), succ, off))
    d_out = torch.empty(1000 * 30_000, dtype=torch.uint8, device="cuda")
    lib = _lib.load()
    evs = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.smg_kmer_shuffle(dev._ptr(d_tok), 30_000, dev._ptr(d_succ), dev._ptr(d_off), 1000, k, 1, None, dev._ptr(d_out),
                                        30_000, dev._stream()), "shuffle")
        e1.record()
        torch.cuda.synchronize()
        evs.append(e0.elapsed_time(e1) * 1e-3)
    rs = np.random.RandomState(1)
    n_cpu = 20
    t0 = time.perf_counter()
    for _ in range(n_cpu):
        if k == 2:
            sho.dinuc_shuffle(genome, rs)
        else:
            "".join(np.array(list(genome))[rs.permutation(len(genome))])
    t_cpu = (time.perf_counter() - t0) / n_cpu * 1000
    out["runs"].append({"simK": k, "device_kernel_s": float(np.median(evs)), "device_shuffle_and_pack_s": t_e2e,
                        "bases_per_s_kernel": 30e6 / float(np.median(evs)),
                        "cpu_restatement_s_for_1000": t_cpu, "cpu_sample_copies": n_cpu, "cpu_cores": 1,
                        "speedup_vs_cpu_restatement": t_cpu / t_e2e})
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
