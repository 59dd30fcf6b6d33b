"""Host glue kept from the reference's smeagol/utils.py for the scan path callers."""
import numpy as np

from .io import read_fasta
from .seqrecord import Seq, SeqRecord, is_record


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


def seq_str(seq):
    """Forward string of a str / Seq / SeqRecord."""
    if type(seq) == str:
        return seq
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


def get_site_seq(site, seqs):
    """Sequence of a binding site, sliced from the FORWARD record (utils.py:183-195)."""
    if type(seqs) == str:
        seqs = read_fasta(seqs)
    seq = [x for x in seqs if x.name == site["name"]][0]
    return str(seq.seq[site.start:site.end])


def shuffle_records(records, simN, simK=2, out_file=None, seed=1):
    """Background generation (utils.py:36-83) is outside the accelerated path ("next" rank 1):
    BASELINE config 3 uses fixed pre-generated shuffles.  Only the mononucleotide shuffle (simK=1), which needs no
    third-party code, is provided; the dinucleotide shuffle of the reference lives in deeplift."""
    import random

    if simK != 1:
        raise NotImplementedError("dinucleotide shuffling (deeplift.dinuc_shuffle) is out of scope; pass shuf_seqs")
    shuf_records = []
    for record in records:
        seq = str(record.seq)
        random.seed(seed)
        for i in range(simN):
            shuffled = list(seq)
            random.shuffle(shuffled)
            shuf_records.append(SeqRecord(Seq("".join(shuffled)), id="background_seq_{0:d}".format(i), name=record.id))
    return shuf_records
