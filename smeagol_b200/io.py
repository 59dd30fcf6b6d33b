"""FASTA read/write used by the scan path (reference: smeagol/io.py:16-67).  PWM file readers and database
downloaders of the reference are out of scope."""
import gzip
from functools import partial
from mimetypes import guess_type

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
            elif name is not None:
                chunks.append(line.strip())
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
