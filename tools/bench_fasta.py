"""FASTA ingest: device parser + packer against the host parser, on synthetic files in memory.

    python tools/bench_fasta.py [n_bases] [out.json]

Workloads: one record of n_bases (60-column lines, config c4's shape) and 1,000 records of n_bases / 1,000 (config
c2's shape).  Device time: CUDA events around smg_fasta_index + sort of the header offsets + smg_fasta_compact +
smg_pack with the file's bytes already in HBM (algorithmic traffic: read the text twice for index and count, once
for the compaction, write + read the compacted bases, write 0.375 B/base packed ~= 5.4 B per input byte);
end to end: wall clock from the host byte buffer to a ready SeqBatch.  Host: smeagol_b200.io.read_fasta (the port of
the reference's reader) + SeqBatch from its records."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smeagol_b200 import device as dev  # noqa: E402
from smeagol_b200 import io as sio  # noqa: E402

n_bases = int(sys.argv[1]) if len(sys.argv) > 1 else 248_000_000
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}


def make_fasta(n_records, total):
    rng = np.random.default_rng(7)
    per = total // n_records
    seq = rng.integers(0, 4, size=per, dtype=np.uint8)
    body = np.frombuffer(b"ACGT", dtype=np.uint8)[seq]
    lines = per // 60
    wrapped = np.full((lines, 61), ord("\n"), dtype=np.uint8)
    wrapped[:, :60] = body[:lines * 60].reshape(lines, 60)
    rec = wrapped.tobytes() + body[lines * 60:].tobytes() + b"\n"
    return b"".join((">seq%d synthetic record\n" % i).encode() + rec for i in range(n_records))


out = {"n_bases": n_bases, "hbm_peak_gbs": peaks["hbm_gbs"], "workloads": []}
for n_records in (1, 1000):
    data = make_fasta(n_records, n_bases)
    n = len(data)
    # ---- device, end to end (host bytes -> SeqBatch), and device-resident stages
    torch.cuda.synchronize()
    t_e2e = []
    for _ in range(3):
        t0 = time.perf_counter()
        fb = sio.read_fasta_device(None, data=data)
        torch.cuda.synchronize()
        t_e2e.append(time.perf_counter() - t0)
    # device-resident: repeat the kernels on a resident copy
    import ctypes
    from smeagol_b200 import _lib
    lib = _lib.load()
    host = np.frombuffer(data, dtype=np.uint8)
    d_src = torch.from_numpy(host.copy()).cuda()
    d_bytes = torch.empty_like(d_src)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for _ in range(5):
        d_bytes.copy_(d_src)
        torch.cuda.synchronize()
        ev[0].record()
        d_hdr = torch.empty(1 << 16, dtype=torch.int64, device="cuda")
        d_nh = torch.zeros(1, dtype=torch.int64, device="cuda")
        _lib.check(lib.smg_fasta_index(dev._ptr(d_bytes), n, dev._ptr(d_hdr), 1 << 16, dev._ptr(d_nh), dev._stream()), "index")
        n_hdr = int(d_nh.item())
        hs = torch.sort(d_hdr[:n_hdr]).values.contiguous()
        d_end = torch.empty(n_hdr, dtype=torch.int64, device="cuda")
        d_len = torch.empty(n_hdr, dtype=torch.int64, device="cuda")
        d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
        d_nout = torch.zeros(1, dtype=torch.int64, device="cuda")
        tb = ctypes.c_uint64(0)
        args = (dev._ptr(d_bytes), n, dev._ptr(hs), n_hdr, dev._ptr(d_end), dev._ptr(d_len), dev._ptr(d_out), dev._ptr(d_nout), 0)
        _lib.check(lib.smg_fasta_compact(*args, None, ctypes.byref(tb), dev._stream()), "size")
        d_tmp = torch.empty(max(int(tb.value), 16), dtype=torch.uint8, device="cuda")
        _lib.check(lib.smg_fasta_compact(*args, dev._ptr(d_tmp), ctypes.byref(tb), dev._stream()), "compact")
        lens = d_len.cpu().numpy().astype(np.int64)
        src_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        b = dev.SeqBatch(None, flat=(d_out, src_off, lens))
        ev[1].record()
        torch.cuda.synchronize()
        times.append(ev[0].elapsed_time(ev[1]) * 1e-3)
    t_dev = float(np.median(times))
    traffic = 3 * n + 2 * int(lens.sum()) + 0.375 * int(lens.sum())
    # ---- host parser (bounded: at most 32 MB of the file, extrapolated linearly)
    cut = data if n <= (32 << 20) else data[:data.rfind(b"\n", 0, 32 << 20) + 1]
    path = os.path.join(tempfile.mkdtemp(), "x.fa")
    with open(path, "wb") as fh:
        fh.write(cut)
    t0 = time.perf_counter()
    recs = sio.read_fasta(path)
    hb = dev.SeqBatch([str(r.seq) for r in recs])
    torch.cuda.synchronize()
    t_host = (time.perf_counter() - t0) * (n / len(cut))
    out["workloads"].append({"records": n_records, "file_bytes": n, "device_resident_s": t_dev,
                             "device_resident_gbs_of_file": n / t_dev / 1e9,
                             "algorithmic_traffic_bytes": int(traffic), "achieved_gbs": traffic / t_dev / 1e9,
                             "frac_of_hbm_peak": traffic / t_dev / 1e9 / peaks["hbm_gbs"],
                             "end_to_end_from_host_bytes_s": float(np.median(t_e2e)),
                             "host_parser_s": t_host, "host_parser_sample_bytes": len(cut),
                             "speedup_end_to_end": t_host / float(np.median(t_e2e))})
    del fb, d_src, d_bytes, d_out, b, hb
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
