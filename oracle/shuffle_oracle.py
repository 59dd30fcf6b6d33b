"""TEST INFRASTRUCTURE ONLY (see oracle/README or DESIGN section 6): CPU restatement of the k-mer preserving shuffles
behind smeagol/utils.py:36-83.

simK = 2 calls deeplift.dinuc_shuffle, a third-party dependency that is NOT under /root/reference (requirements.txt:
`deeplift`, unpinned).  Its published algorithm is restated here: tokenise the characters; for every token collect
the indices that FOLLOW an occurrence of it; permute each such list except for its last element; re-walk from
position 0, always taking the next unused entry of the current token's list.  The result is an Eulerian path through
the same dinucleotide multigraph: every dinucleotide count, the first and the last character are preserved.
simK = 1 is random.shuffle of the characters.  Parity with the device kernel is on these properties and on the
sampling distribution (the random streams differ) -- "parity unpinned" at the byte level, by construction."""
import numpy as np


def dinuc_shuffle(seq, rng):
    """One dinucleotide-preserving shuffle of a str (restatement of deeplift.dinuc_shuffle for a string input)."""
    arr = np.frombuffer(seq.encode("ascii"), dtype=np.uint8)
    if arr.size < 2:
        return seq
    chars, tokens = np.unique(arr, return_inverse=True)
    nxt = [np.flatnonzero(tokens[:-1] == t) + 1 for t in range(len(chars))]
    for t in range(len(chars)):
        n = len(nxt[t])
        if n > 1:
            inds = np.arange(n)
            inds[:-1] = rng.permutation(n - 1)
            nxt[t] = nxt[t][inds]
    counters = [0] * len(chars)
    ind = 0
    out = np.empty_like(tokens)
    out[0] = tokens[0]
    for j in range(1, len(tokens)):
        t = tokens[ind]
        ind = nxt[t][counters[t]]
        counters[t] += 1
        out[j] = tokens[ind]
    return chars[out].tobytes().decode("ascii")


def kmer_counts(seq, k):
    """Counts of every k-mer of a str as a dict."""
    out = {}
    for i in range(len(seq) - k + 1):
        out[seq[i:i + k]] = out.get(seq[i:i + k], 0) + 1
    return out


def check_shuffle(orig, shuf, k):
    """Properties every simK = k shuffle of the reference has."""
    assert len(shuf) == len(orig)
    assert kmer_counts(shuf, 1) == kmer_counts(orig, 1)
    if k == 2 and len(orig) > 1:
        assert kmer_counts(shuf, 2) == kmer_counts(orig, 2)
        assert shuf[0] == orig[0] and shuf[-1] == orig[-1]
