"""Host glue kept from the reference's smeagol/utils.py for the scan path callers."""
import numpy as np

from .io import read_fasta
from .seqrecord import DeviceSeq, Seq, SeqRecord, is_record


def _equals(x, y, eps=1e-4):
    """Tolerance helper of the reference tests (utils.py:14-33): abs diff < eps * size."""
    if type(x) == type(y) == np.ndarray:
        assert x.shape == y.shape
        return np.all(abs(x - y) < (eps * x.size))
    elif type(x) == type(y) == list:
        assert len(x) == len(y)
        return np.all(abs(np.asarray(x) - np.asarray(y)) < (eps * len(x)))
    else:
        return abs(x - y) < eps


def get_Seq(seq):
    """str / Seq / SeqRecord -> Seq (utils.py:86-106); TypeError otherwise."""
    if type(seq) == str:
        return Seq(seq)
    if isinstance(seq, Seq):
        return seq
    if is_record(seq):
        return seq.seq if isinstance(seq.seq, Seq) else Seq(str(seq.seq))
    raise TypeError("Input sequence must be a string, Seqrecord or Seq object.")


def seq_str(seq, keep_device=False):
    """Forward string of a str / Seq / SeqRecord (keep_device: a device-resident sequence is returned as it is)."""
    if type(seq) == str:
        return seq
    if keep_device and isinstance(seq, DeviceSeq):
        return seq
    if keep_device and is_record(seq) and isinstance(seq.seq, DeviceSeq):
        return seq.seq
    if isinstance(seq, DeviceSeq):
        return str(seq)
    if is_record(seq):
        return str(seq.seq)
    if isinstance(seq, Seq):
        return str(seq)
    raise TypeError("Input sequence must be a string, Seqrecord or Seq object.")


def read_bg_seqs(file, records, simN):
    """Read pre-generated background sequences and attach the originating record's name (utils.py:109-132)."""
    bg = read_fasta(file)
    assert len(bg) == len(records) * simN
    for i, bg_record in enumerate(bg):
        matching_record = records[i // simN]
        assert len(bg_record) == len(matching_record), "length of shuffled sequence does not match original sequence."
        bg_record.name = matching_record.name
    return bg


def _get_tiling_windows_over_record(record, width, shift):
    """Tiling windows over one sequence (utils.py:135-155): id, start, end."""
    import pandas as pd

    starts = np.arange(0, len(record), shift)
    return pd.DataFrame({"id": np.tile(record.id, len(starts)), "start": starts, "end": np.minimum(starts + width, len(record))})


def get_tiling_windows_over_genome(genome, width, shift=None):
    """Tiling windows over every sequence of a genome (utils.py:158-180); shift defaults to width."""
    import pandas as pd

    if shift is None:
        shift = width
    if len(genome) == 1:
        return _get_tiling_windows_over_record(genome[0], width, shift)
    return pd.concat([_get_tiling_windows_over_record(record, width, shift) for record in genome])


def get_site_seq(site, seqs):
    """Sequence of a binding site, sliced from the FORWARD record (utils.py:183-195)."""
    if type(seqs) == str:
        seqs = read_fasta(seqs)
    seq = [x for x in seqs if x.name == site["name"]][0]
    return str(seq.seq[site.start:site.end])


def _successor_buckets(tok):
    """Successors tok[i+1] bucketed by tok[i] in original order + bucket offsets (17 symbols): the per-symbol lists
    that the dinucleotide shuffle permutes (deeplift.dinuc_shuffle's shuf_next_inds, as symbols instead of indices)."""
    order = np.argsort(tok[:-1], kind="stable")
    succ = np.ascontiguousarray(tok[1:][order])
    off = np.zeros(18, dtype=np.int64)
    off[1:] = np.cumsum(np.bincount(tok[:-1], minlength=17))
    return succ, off


def shuffle_batch_device(records, simN, simK=2, seed=1, device=None):
    """simN k-mer preserving shuffles of every record, generated and packed on the GPU (csrc/shuffle.cu): the copies
    never exist on the host.  Returns (ids, names, SeqBatch, ascii) with the reference's naming (utils.py:68-71):
    id 'background_seq_i', name = the record's id; rows are ordered record by record, copy by copy.
    Seeded and deterministic, but not the numpy / Python random stream of the reference."""
    import torch

    from . import _lib
    from . import device as dev
    from .encode import integer_encoding_dict

    if simK not in (1, 2):
        raise ValueError("simK must be 1 or 2.")
    d = dev._require_cuda(device)
    lib = _lib.load()
    lut = np.full(256, 255, dtype=np.uint8)
    for ch, code in integer_encoding_dict.items():
        lut[ord(ch)] = code
    ids, names, lens, outs = [], [], [], []
    with torch.cuda.device(d):
        for r, record in enumerate(records):
            raw = np.frombuffer(str(record.seq).encode("ascii"), dtype=np.uint8)
            tok = lut[raw]
            if tok.size and tok.max() == 255:
                raise KeyError(chr(int(raw[int(np.argmax(tok == 255))])))
            L = int(tok.shape[0])
            stride = (L + 15) // 16 * 16
            d_out = torch.zeros(max(simN * stride, 16), dtype=torch.uint8, device=d)
            if L > 0 and simN > 0:
                d_tok = torch.from_numpy(tok.copy()).to(d)
                if simK == 2 and L > 1:
                    succ, off = _successor_buckets(tok)
                    d_succ, d_off = torch.from_numpy(succ).to(d), torch.from_numpy(off).to(d)
                else:
                    d_succ = d_off = None
                d_scr = torch.empty(simN * L, dtype=torch.uint8, device=d) if L > 200 * 1024 else None
                k = simK if L > 1 else 1
                _lib.check(lib.smg_kmer_shuffle(dev._ptr(d_tok), L, dev._ptr(d_succ), dev._ptr(d_off), simN, k,
                                                (int(seed) * 0x9E3779B1 + r) & (2 ** 64 - 1), dev._ptr(d_scr), dev._ptr(d_out), stride,
                                                dev._stream()), "smg_kmer_shuffle")
            outs.append((d_out, stride, L))
            ids += ["background_seq_{0:d}".format(i) for i in range(simN)]
            names += [record.id] * simN
            lens += [L] * simN
        flat = torch.cat([o for o, _, _ in outs]) if len(outs) > 1 else outs[0][0]
        src_off, base = [], 0
        for o, stride, L in outs:
            src_off += [base + i * stride for i in range(simN)]
            base += int(o.numel())
        batch = dev.SeqBatch(None, flat=(flat, np.asarray(src_off, dtype=np.int64), np.asarray(lens, dtype=np.int64)), device=d)
    return ids, names, batch, (flat, np.asarray(src_off, dtype=np.int64), np.asarray(lens, dtype=np.int64))


def shuffle_records(records, simN, simK=2, out_file=None, seed=1):
    """Shuffle sequences conserving k-mer frequencies, k = 1 or 2 (utils.py:36-83; same signature, ids and names).
    The copies are generated on the GPU (shuffle_batch_device); the returned SeqRecords hold device-resident sequences
    (seqrecord.DeviceSeq) that are read back only when a caller asks for their text.  simK = 2 is the algorithm of deeplift.dinuc_shuffle (every dinucleotide
    count, the first and the last base are preserved), with a different, seeded random stream."""
    ids, names, _, (flat, src_off, lens) = shuffle_batch_device(records, simN, simK, seed)
    # the copies stay on the device: .seq is a DeviceSeq (str() / len() / slicing read it back on demand), and
    # scan_sequences packs such records without a host round trip
    shuf_records = [SeqRecord(DeviceSeq(flat, o, n), id=i, name=m) for i, m, o, n in zip(ids, names, src_off, lens)]
    print("Shuffled {} sequence(s) {} times while conserving k-mer frequency for k = {}.".format(len(records), simN, simK))
    if out_file is not None:
        from .io import write_fasta

        write_fasta(shuf_records, out_file)
    return shuf_records
