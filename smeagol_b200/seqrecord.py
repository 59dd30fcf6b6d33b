"""Sequence containers.  The reference uses Biopython's Seq / SeqRecord (smeagol/utils.py:86-106); Biopython is
used here when it is installed, otherwise these minimal stand-ins with the same attributes (.seq, .id, .name,
.description, len(), str(), slicing, reverse_complement()) are used.  Any object with .seq/.id/.name works."""

try:  # pragma: no cover - depends on the environment
    from Bio.Seq import Seq  # type: ignore
    from Bio.SeqRecord import SeqRecord  # type: ignore

    HAVE_BIOPYTHON = True
except Exception:  # Biopython absent
    HAVE_BIOPYTHON = False

    _DNA_COMP = str.maketrans("ACGTMRWSYKVHDBNacgtmrwsykvhdbn", "TGCAKYWSRMBDHVNtgcakywsrmbdhvn")
    _RNA_COMP = str.maketrans("ACGUMRWSYKVHDBNacgumrwsykvhdbn", "UGCAKYWSRMBDHVNugcakywsrmbdhvn")

    class Seq(object):
        def __init__(self, data):
            self._data = str(data)

        def __str__(self):
            return self._data

        def __repr__(self):
            return "Seq(%r)" % (self._data if len(self._data) < 60 else self._data[:57] + "...")

        def __len__(self):
            return len(self._data)

        def __iter__(self):
            return iter(self._data)

        def __getitem__(self, i):
            return Seq(self._data[i]) if isinstance(i, slice) else self._data[i]

        def __eq__(self, other):
            return str(self) == str(other)

        def __hash__(self):
            return hash(self._data)

        def reverse_complement(self):
            """Biopython 1.79 behaviour (the reference pins biopython>=1.79 and its tests/test_encode.py:21-23
            expects the RNA complement of "AGCAU"): U without T -> RNA alphabet, both -> ValueError."""
            d = self._data
            has_u = ("U" in d) or ("u" in d)
            has_t = ("T" in d) or ("t" in d)
            if has_u and has_t:
                raise ValueError("Mixed RNA/DNA found")
            return Seq(d.translate(_RNA_COMP if has_u else _DNA_COMP)[::-1])

    class SeqRecord(object):
        def __init__(self, seq, id="<unknown id>", name="<unknown name>", description="<unknown description>"):
            self.seq = seq
            self.id = id
            self.name = name
            self.description = description

        def __len__(self):
            return len(self.seq)

        def __repr__(self):
            return "SeqRecord(seq=%r, id=%r, name=%r)" % (self.seq, self.id, self.name)


class DeviceSeq(object):
    """A sequence that lives in a device ASCII buffer (produced by the GPU shuffle / FASTA ingest).  Behaves like a Seq
    for the reference API -- len(), str(), slicing, reverse_complement() read the bases back on demand -- while the scan
    path recognises it and packs it on the device without a host round trip (scan._device_scan_groups)."""

    def __init__(self, flat, offset, length):
        self.flat, self.offset, self.length = flat, int(offset), int(length)
        self._host = None

    def __len__(self):
        return self.length

    def __str__(self):
        if self._host is None:
            self._host = self.flat[self.offset:self.offset + self.length].cpu().numpy().tobytes().decode("ascii")
        return self._host

    def __repr__(self):
        return "DeviceSeq(length=%d, device=%s)" % (self.length, self.flat.device)

    def __iter__(self):
        return iter(str(self))

    def __getitem__(self, i):
        return Seq(str(self)[i]) if isinstance(i, slice) else str(self)[i]

    def __eq__(self, other):
        return str(self) == str(other)

    def __hash__(self):
        return hash((id(self.flat), self.offset, self.length))

    def reverse_complement(self):
        return Seq(str(self)).reverse_complement()


def is_record(x):
    return hasattr(x, "seq") and hasattr(x, "id")
