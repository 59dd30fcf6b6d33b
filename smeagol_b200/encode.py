"""Sequence encoding -- API mirror of smeagol/encode.py.

The reference encodes every base with a Python dict lookup (encode.py:66) into float32 integer codes and stacks
(N, L) arrays.  Here the records are kept as forward strings; the device path (smeagol_b200.device.SeqBatch) uploads
the ASCII bytes and packs them to 2 bits/base + ambiguity mask on the GPU.  The reference-visible attributes
(``SeqEncoding.seqs`` etc.) are materialised lazily on the host with a vectorised table lookup, only when read."""
from collections import defaultdict

import numpy as np

from .io import read_fasta
from .utils import get_Seq, seq_str

# encode.py:11-41
one_hot_dict = {
    "Z": [0, 0, 0, 0],
    "A": [1, 0, 0, 0],
    "C": [0, 1, 0, 0],
    "G": [0, 0, 1, 0],
    "T": [0, 0, 0, 1],
    "U": [0, 0, 0, 1],
    "N": [1 / 4, 1 / 4, 1 / 4, 1 / 4],
    "W": [1 / 2, 0, 0, 1 / 2],
    "S": [0, 1 / 2, 1 / 2, 0],
    "M": [1 / 2, 1 / 2, 0, 0],
    "K": [0, 0, 1 / 2, 1 / 2],
    "R": [1 / 2, 0, 1 / 2, 0],
    "Y": [0, 1 / 2, 0, 1 / 2],
    "B": [0, 1 / 3, 1 / 3, 1 / 3],
    "D": [1 / 3, 0, 1 / 3, 1 / 3],
    "H": [1 / 3, 1 / 3, 0, 1 / 3],
    "V": [1 / 3, 1 / 3, 1 / 3, 0],
}
sense_complement_dict = {"+": "-", "-": "+"}
bases = list(one_hot_dict.keys())
base_one_hot = list(one_hot_dict.values())
integer_encoding_dict = {base: i for i, base in enumerate(bases)}

_LUT = np.full(256, 255, dtype=np.uint8)
for _b, _i in integer_encoding_dict.items():
    _LUT[ord(_b)] = _i
_ONE_HOT32 = np.array(base_one_hot, dtype=np.float32)


def _codes_u8(seq_string):
    raw = np.frombuffer(seq_string.encode("latin-1"), dtype=np.uint8)
    codes = _LUT[raw]
    bad = np.flatnonzero(codes == 255)
    if bad.size:
        raise KeyError(seq_string[int(bad[0])])  # the reference's dict lookup raises KeyError (encode.py:66)
    return codes


def integer_encode(seq, rc=False):
    """Encode a nucleic acid sequence as integers 0..16 (float32, shape (L,)) -- encode.py:47-68."""
    seq = get_Seq(seq)
    if rc:
        seq = seq.reverse_complement()
    return _codes_u8(str(seq)).astype(np.float32)


def one_hot_encode(seq, rc=False):
    """Soft one-hot encode a sequence -> float32 (L, 4) -- encode.py:71-90."""
    if rc:
        seq = get_Seq(seq).reverse_complement()
    return _ONE_HOT32[_codes_u8(seq_str(seq))]


class SeqEncoding:
    """A set of equal-length sequences of one sense (encode.py:93-176).

    Attributes (as the reference): len, ids, names, senses, seqs (N, L) integer codes with rows
    [fwd_0..fwd_{n-1}] (rcomp='none'), [rc_0..] ('only') or [fwd..., rc...] ('both').
    Extra (used by the device path): records (forward strings), rcomp, sense, n_records.
    """

    def __init__(self, records, rcomp="none", sense=None):
        if type(records) == str:  # the reference intends this (encode.py:124 compares against the string "str")
            records = read_fasta(records)
        self.records = [seq_str(r, keep_device=True) for r in records]  # str, or DeviceSeq for device-resident records
        self._check_equal_lens(self.records)
        self.len = len(self.records[0])
        self.n_records = len(self.records)
        self.rcomp = rcomp
        self.sense = sense
        ids = np.array([record.id for record in records])
        names = np.array([record.name for record in records])
        assert rcomp in ["none", "both", "only"], "rcomp should be 'none', 'only' or 'both'."
        n = self.n_records
        if rcomp == "none":
            self.senses = np.array([sense] * n)
        elif rcomp == "only":
            self.senses = np.array([sense_complement_dict[sense]] * n)
        else:
            self.senses = np.concatenate([[sense] * n, [sense_complement_dict[sense]] * n])
            ids = np.tile(ids, 2)
            names = np.tile(names, 2)
        self.ids = ids
        self.names = names
        self._seqs = None

    @property
    def seqs(self):
        """(N, L) float32 integer codes, materialised on first access (the scan itself never needs it)."""
        if self._seqs is None:
            rows = []
            if self.rcomp in ("none", "both"):
                rows += [integer_encode(str(s), rc=False) for s in self.records]
            if self.rcomp in ("only", "both"):
                rows += [integer_encode(str(s), rc=True) for s in self.records]
            self._seqs = np.vstack(rows)
        return self._seqs

    def _check_equal_lens(self, records):
        if len(records) > 1:
            lens = np.unique([len(r) for r in records])
            if len(lens) != 1:
                raise ValueError("Cannot encode - sequences have unequal length!")


class SeqGroups:
    """One or more groups of equal-length sequences (encode.py:179-218)."""

    def __init__(self, records, rcomp="none", sense=None, group_by=None):
        if type(records) == str:
            records = read_fasta(records)
        if (group_by is not None) and (len(records)) > 1:
            grouped = self._group_by(records, group_by)
            self.seqs = [SeqEncoding(group, sense=sense, rcomp=rcomp) for group in grouped]
        else:
            self.seqs = [SeqEncoding([record], sense=sense, rcomp=rcomp) for record in records]

    def _group_by(self, records, key):
        records_dict = defaultdict(list)
        for record in records:
            records_dict[getattr(record, key)].append(record)
        return records_dict.values()
