"""FASTA read/write used by the scan path (reference: smeagol/io.py:16-67).  PWM file readers and database
downloaders of the reference are out of scope.

read_fasta         host parser, list of SeqRecord (the reference's return type)
read_fasta_device  the same file parsed on the GPU straight into a packed SeqBatch: the bytes of
                   the file are uploaded once, header lines are located and line breaks removed by CUDA kernels
                   (csrc/fasta.cu), and smg_pack encodes the result -- no Python string per record or per line.
"""
import gzip
from functools import partial
from mimetypes import guess_type

import numpy as np

from .seqrecord import Seq, SeqRecord


def read_fasta(file):
    """Read sequences from a fasta or fasta.gz file -> list of SeqRecord (io.py:16-41)."""
    records = []
    encoding = guess_type(file)[1]
    _open = partial(gzip.open, mode="rt") if encoding == "gzip" else open
    name, desc, chunks = None, "", []
    with _open(file) as handle:
        for line in handle:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    records.append(SeqRecord(Seq("".join(chunks)), id=name, name=name, description=desc))
                desc = line[1:]
                name = desc.split()[0] if desc.split() else ""
                chunks = []
            elif name is not None:  # Biopython: lines right-stripped, then blanks and carriage returns removed
                chunks.append(line.strip().replace(" ", "").replace("\r", ""))
    if name is not None:
        records.append(SeqRecord(Seq("".join(chunks)), id=name, name=name, description=desc))
    print("Read " + str(len(records)) + " records from " + file)
    return records


def write_fasta(records, file, gz=True):
    """Write SeqRecords to a fasta or fasta.gz file (io.py:44-67)."""
    _open = partial(gzip.open, mode="wt") if file.endswith(".gz") else partial(open, mode="wt")
    with _open(file) as handle:
        for record in records:
            handle.write(">%s %s\n" % (record.id, getattr(record, "description", "")))
            s = str(record.seq)
            for i in range(0, len(s), 60):
                handle.write(s[i:i + 60] + "\n")
    print("Wrote " + str(len(records)) + " sequences to " + file)


def _header_fields(line):
    """'>id rest of the description' -> (id, name, description) as Biopython's FastaIterator sets them."""
    desc = line[1:].rstrip("\r\n")
    tok = desc.split()
    name = tok[0] if tok else ""
    return name, name, desc


class FastaBatch:
    """Result of read_fasta_device: per-record ids / names / descriptions / lengths and the packed device batch."""

    def __init__(self, ids, names, descriptions, lengths, batch, ascii_dev, src_off):
        self.ids, self.names, self.descriptions = ids, names, descriptions
        self.lengths, self.batch = lengths, batch
        self._ascii, self._src_off = ascii_dev, src_off

    def __len__(self):
        return len(self.ids)

    def records(self):
        """SeqRecords for callers of the reference API (scan_sequences, enrich_in_genome, ...): their sequences stay on
        the device (seqrecord.DeviceSeq) and are read back only when their text is asked for."""
        from .seqrecord import DeviceSeq

        return [SeqRecord(DeviceSeq(self._ascii, o, n), id=i, name=m, description=d)
                for i, m, d, o, n in zip(self.ids, self.names, self.descriptions, self._src_off, self.lengths)]


def read_fasta_device(file, device=None, upper=False, data=None):
    """Parse a fasta / fasta.gz file on the GPU into a packed SeqBatch (same records as read_fasta).

    file: path (or None with data= the file's bytes); upper: map lower-case bases to upper case on the fly (the
    reference raises KeyError for lower case, encode.py:66 -- soft-masked genomes need upper=True).
    Returns a FastaBatch; .batch feeds smeagol_b200.device.scan / ScanPlan directly."""
    import ctypes

    import torch

    from . import _lib
    from . import device as dev

    if data is None:
        if guess_type(file)[1] == "gzip":
            with gzip.open(file, "rb") as fh:
                data = fh.read()
        else:
            data = np.fromfile(file, dtype=np.uint8)
    host = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    n = int(host.shape[0])
    if n >= 2 ** 31:
        raise ValueError("read_fasta_device handles files below 2 GiB per call; split the file by record")
    d = dev._require_cuda(device)
    lib = _lib.load()
    with torch.cuda.device(d):
        d_bytes = torch.empty(max(n, 16), dtype=torch.uint8, device=d)
        if n:
            import warnings
            with warnings.catch_warnings():  # bytes objects give read-only arrays; the tensor is only read
                warnings.simplefilter("ignore")
                d_bytes[:n].copy_(torch.from_numpy(host), non_blocking=False)
        cap = 1 << 16
        while True:
            d_hdr = torch.empty(cap, dtype=torch.int64, device=d)
            d_nh = torch.zeros(1, dtype=torch.int64, device=d)
            _lib.check(lib.smg_fasta_index(dev._ptr(d_bytes), n, dev._ptr(d_hdr), cap, dev._ptr(d_nh), dev._stream()),
                       "smg_fasta_index")
            n_hdr = int(d_nh.item())
            if n_hdr <= cap:
                break
            cap = n_hdr
        if n_hdr == 0:
            raise ValueError("no FASTA record in " + str(file))
        d_hdr = torch.sort(d_hdr[:n_hdr]).values.contiguous()
        d_end = torch.empty(n_hdr, dtype=torch.int64, device=d)
        d_len = torch.empty(n_hdr, dtype=torch.int64, device=d)
        d_out = torch.empty(max(n, 16), dtype=torch.uint8, device=d)
        d_nout = torch.zeros(1, dtype=torch.int64, device=d)
        tb = ctypes.c_uint64(0)
        args = (dev._ptr(d_bytes), n, dev._ptr(d_hdr), n_hdr, dev._ptr(d_end), dev._ptr(d_len), dev._ptr(d_out), dev._ptr(d_nout),
                1 if upper else 0)
        _lib.check(lib.smg_fasta_compact(*args, None, ctypes.byref(tb), dev._stream()), "smg_fasta_compact(size)")
        d_tmp = torch.empty(max(int(tb.value), 16), dtype=torch.uint8, device=d)
        _lib.check(lib.smg_fasta_compact(*args, dev._ptr(d_tmp), ctypes.byref(tb), dev._stream()), "smg_fasta_compact")
        starts = d_hdr.cpu().numpy()
        ends = d_end.cpu().numpy()
        lens = d_len.cpu().numpy().astype(np.int64)
        total = int(d_nout.item())
    assert total == int(lens.sum()), "FASTA compaction and per-record counts disagree"
    ids, names, descs = [], [], []
    for a, b in zip(starts, ends):
        i, m, ds = _header_fields(host[a:b].tobytes().decode("utf-8", errors="replace"))
        ids.append(i)
        names.append(m)
        descs.append(ds)
    src_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    batch = dev.SeqBatch(None, flat=(d_out[:max(total, 16)], src_off, lens), device=d)
    print("Read " + str(n_hdr) + " records from " + str(file))
    return FastaBatch(ids, names, descs, lens, batch, d_out[:total], src_off)
