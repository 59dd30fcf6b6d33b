"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Harness that imports the UNMODIFIED reference package from /root/reference in a container that lacks its
third-party dependencies (tensorflow, keras, biopython, deeplift, statsmodels).  The reference's own
`encode.py`, `models.py`, `scan.py`, `enrich.py`, `variant.py`, `utils.py` run verbatim; only the missing
third-party modules are replaced by minimal stand-ins:

* ``Bio.Seq.Seq`` / ``Bio.SeqRecord.SeqRecord`` / ``Bio.SeqIO``  -> tiny pure-python equivalents
* ``tensorflow.keras`` functional API (Input, ZeroPadding1D, Embedding, Reshape, Conv1D, Model) -> the same
  graph evaluated with ``torch.nn.functional.conv1d`` on CPU in fp32 (reference graph: models.py:50-68)
* ``deeplift.dinuc_shuffle`` / ``statsmodels.stats.multitest`` -> stubs (fdrcorrection restated: BH procedure)
* ``scipy.stats.binom_test`` (removed in SciPy >= 1.12; reference enrich.py:66) -> ``binomtest(...).pvalue``

It is used ONLY by ``tests/golden/make_golden.py`` (run in the build container, where /root/reference exists)
to generate the committed golden vectors.  /root/reference does not exist on the GPU box, so nothing in
tests/, smoke() or bench.py imports this file at run time.
"""
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


# ----------------------------------------------------------------------------- Bio stand-ins
_DNA_COMP = str.maketrans("ACGTMRWSYKVHDBNacgtmrwsykvhdbn", "TGCAKYWSRMBDHVNtgcakywsrmbdhvn")
_RNA_COMP = str.maketrans("ACGUMRWSYKVHDBNacgumrwsykvhdbn", "UGCAKYWSRMBDHVNugcakywsrmbdhvn")


class Seq:
    def __init__(self, data):
        self._data = str(data)

    def __str__(self):
        return self._data

    def __repr__(self):
        return "Seq(%r)" % self._data

    def __len__(self):
        return len(self._data)

    def __iter__(self):
        return iter(self._data)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return Seq(self._data[i])
        return self._data[i]

    def __eq__(self, other):
        return str(self) == str(other)

    def __hash__(self):
        return hash(self._data)

    def reverse_complement(self):
        # Biopython 1.79 semantics (requirements.txt pins >=1.79; tests/test_encode.py:21-23 expects the RNA
        # complement for "AGCAU"): U without T -> RNA table, otherwise DNA table.
        d = self._data
        if ("U" in d or "u" in d) and not ("T" in d or "t" in d):
            return Seq(d.translate(_RNA_COMP)[::-1])
        if ("U" in d or "u" in d) and ("T" in d or "t" in d):
            raise ValueError("Mixed RNA/DNA found")
        return Seq(d.translate(_DNA_COMP)[::-1])


class SeqRecord:
    def __init__(self, seq, id="<unknown id>", name="<unknown name>", description="<unknown description>"):
        self.seq = seq
        self.id = id
        self.name = name
        self.description = description

    def __len__(self):
        return len(self.seq)


def _fasta_parse(handle, fmt):
    assert fmt == "fasta"
    name, chunks, desc = None, [], ""
    for line in handle:
        line = line.rstrip("\n").rstrip("\r")
        if line.startswith(">"):
            if name is not None:
                yield SeqRecord(Seq("".join(chunks)), id=name, name=name, description=desc)
            desc = line[1:]
            name = desc.split()[0] if desc.split() else ""
            chunks = []
        elif name is not None:
            chunks.append(line.strip())
    if name is not None:
        yield SeqRecord(Seq("".join(chunks)), id=name, name=name, description=desc)


def _fasta_write(record, handle, fmt):
    handle.write(">%s %s\n%s\n" % (record.id, record.description, str(record.seq)))


# ----------------------------------------------------------------------------- keras stand-ins
class _Node:
    def __init__(self, layer, parent):
        self.layer, self.parent = layer, parent


class _Layer:
    def __init__(self):
        self._w = []

    def __call__(self, node):
        return _Node(self, node)

    def set_weights(self, ws):
        self._w = [np.asarray(w, dtype=np.float32) for w in ws]

    @property
    def weights(self):
        return self._w


class InputLayer(_Layer):
    pass


def Input(shape=None):
    return _Node(InputLayer(), None)


class ZeroPadding1D(_Layer):
    def __init__(self, padding):
        super().__init__()
        self.padding = padding

    def run(self, x):
        return np.pad(x, ((0, 0), (self.padding[0], self.padding[1]), (0, 0)))


class Embedding(_Layer):
    def __init__(self, input_dim, output_dim, input_length=None):
        super().__init__()

    def run(self, x):
        return self._w[0][x.astype(np.int64)]  # (N, L, 1, 4)


class Reshape(_Layer):
    def __init__(self, target):
        super().__init__()
        self.target = target

    def run(self, x):
        return x.reshape(x.shape[0], -1, self.target[1])


class Conv1D(_Layer):
    def __init__(self, filters, kernel_size, padding="valid", use_bias=False):
        super().__init__()
        assert padding == "valid" and not use_bias

    def run(self, x):
        import torch

        k = torch.from_numpy(np.ascontiguousarray(self._w[0]))  # (W, 4, C)
        xin = torch.from_numpy(np.ascontiguousarray(x)).permute(0, 2, 1)  # (N, 4, Lp)
        out = torch.nn.functional.conv1d(xin, k.permute(2, 1, 0).contiguous())  # (N, C, L)
        return out.permute(0, 2, 1).contiguous().numpy()


class Model:
    def __init__(self, inputs=None, outputs=None, name=None):
        chain, node = [], outputs
        while node is not None:
            chain.append(node.layer)
            node = node.parent
        self.layers = chain[::-1]

    def predict(self, x, **kw):
        x = np.asarray(x, dtype=np.float32)
        if x.ndim == 2:
            x = x[:, :, None]
        for layer in self.layers[1:]:
            x = layer.run(x)
        return x


def _fdrcorrection(pvals, alpha=0.05):
    """Benjamini-Hochberg, as statsmodels.stats.multitest.fdrcorrection(method='indep')."""
    p = np.asarray(pvals, dtype=float)
    n = len(p)
    order = np.argsort(p)
    ps = p[order]
    adj = ps * n / np.arange(1, n + 1)
    adj = np.minimum.accumulate(adj[::-1])[::-1]
    adj = np.minimum(adj, 1.0)
    out = np.empty(n)
    out[order] = adj
    return out <= alpha, out


def install():
    """Insert the stand-in modules into sys.modules and put the reference on sys.path."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    bio = mod("Bio")
    bio.Seq = mod("Bio.Seq", Seq=Seq)
    bio.SeqRecord = mod("Bio.SeqRecord", SeqRecord=SeqRecord)
    bio.SeqIO = mod("Bio.SeqIO", parse=_fasta_parse, write=_fasta_write)

    tf = mod("tensorflow", float32=np.float32, convert_to_tensor=lambda x, dtype=None: np.asarray(x, dtype=dtype))
    keras = mod("tensorflow.keras", Model=Model)
    layers = mod("tensorflow.keras.layers", Conv1D=Conv1D, Input=Input, Embedding=Embedding, Reshape=Reshape,
                 ZeroPadding1D=ZeroPadding1D)
    tf.keras = keras
    keras.layers = layers

    dl = mod("deeplift")
    dl.dinuc_shuffle = mod("deeplift.dinuc_shuffle", dinuc_shuffle=None)
    sm = mod("statsmodels")
    sm.stats = mod("statsmodels.stats")
    sm.stats.multitest = mod("statsmodels.stats.multitest", fdrcorrection=_fdrcorrection)

    import scipy.stats as st

    if not hasattr(st, "binom_test"):
        st.binom_test = lambda k, n, p, alternative="two-sided": st.binomtest(
            int(k), int(n), p, alternative=alternative).pvalue

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
