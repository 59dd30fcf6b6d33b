"""Multi-GPU sharding of the scan path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The path shards into independent units -- whole sequences / backgrounds (configs c2, c3) or contiguous ranges of a
long sequence with a (W_max - 1)-base right halo (config c4) -- so there is NO data-path collective.  The only
exchange is ONE collective on the int32 count tables per scan: all_reduce(SUM) when ranks hold segments of the same
rows, all_gather when they hold disjoint rows.  Hits stay rank-local (sorted per rank; keys carry global
coordinates), a global site list is materialised only on request (gather_hits).
"""
import numpy as np
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(lens, world):
    """Assign whole sequences to ranks, balanced by bases (longest-first greedy).  Returns a list of index arrays
    (ascending within a rank, so per-rank output order follows the input order)."""
    lens = np.asarray(lens, dtype=np.int64)
    order = np.argsort(-lens, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    owner = np.zeros(len(lens), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += lens[i]
    return [np.flatnonzero(owner == r) for r in range(world)]


def shard_sequence(L, world, halo, align=16):
    """Split [0, L) into `world` contiguous ranges with `align`-aligned cut points.  Returns per rank
    (a, b, e): window starts [a, b) are owned by the rank, bases [a, e) must be resident (e = min(L, b + halo)).
    A window is owned by the range containing its START, so no boundary hit is duplicated or lost."""
    cuts = [0]
    for r in range(1, world):
        c = (L * r // world) // align * align
        cuts.append(max(c, cuts[-1]))
    cuts.append(L)
    return [(cuts[r], cuts[r + 1], min(L, cuts[r + 1] + halo)) for r in range(world)]


def segment_rows(a, b, e, L, seg_len=None, halo=0, align=16):
    """Cut a rank's range [a, b) into segments of at most seg_len owned starts (for H2D pipelining / int32 row
    lengths).  Returns (slices, segments): slices = [(lo, hi)] of the sequence to upload, segments =
    [(g_off, g_len, n_own)] for SeqBatch."""
    if seg_len is None or seg_len >= b - a:
        return [(a, e)], [(a, L, b - a)]
    seg_len = max(align, seg_len // align * align)
    slices, segs = [], []
    s = a
    while s < b:
        t = min(b, s + seg_len)
        slices.append((s, min(L, t + halo)))
        segs.append((s, L, t - s))
        s = t
    return slices, segs


def allreduce_counts(counts):
    """Sum the per-(row, motif, strand) count tables over ranks (ranks hold segments of the same rows)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def allgather_counts(counts):
    """Gather equally shaped count tables of disjoint rows: (rows, M, 2) -> (world, rows, M, 2)."""
    rank, world = world_info()
    if world == 1:
        return counts.unsqueeze(0)
    flat = torch.empty((world * counts.shape[0],) + tuple(counts.shape[1:]), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(flat, counts.contiguous())
    return flat.view((world,) + tuple(counts.shape))


def gather_hits(keys, scores, n_hits):
    """Variable-length gather of the rank-local (sorted) hit lists onto every rank; concatenation in rank order is
    globally sorted when ranks own increasing key ranges (contiguous sequence ranges / row blocks)."""
    rank, world = world_info()
    if world == 1:
        return keys[:n_hits], scores[:n_hits]
    dev = keys.device
    n = torch.tensor([n_hits], dtype=torch.int64, device=dev)
    ns = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    mx = max(max(ns), 1)
    kp = torch.zeros(mx, dtype=keys.dtype, device=dev)
    sp = torch.zeros(mx, dtype=scores.dtype, device=dev)
    kp[:n_hits] = keys[:n_hits]
    sp[:n_hits] = scores[:n_hits]
    kl = [torch.empty(mx, dtype=keys.dtype, device=dev) for _ in range(world)]
    sl = [torch.empty(mx, dtype=scores.dtype, device=dev) for _ in range(world)]
    dist.all_gather(kl, kp)
    dist.all_gather(sl, sp)
    return torch.cat([k[:c] for k, c in zip(kl, ns)]), torch.cat([s[:c] for s, c in zip(sl, ns)])


def long_sequence_batch(seq_u8, L, a, b, e, halo, seg_len=1 << 26):
    """SeqBatch of one rank's range [a, b) (+ halo up to e) of a long sequence, cut into segments of seg_len owned
    starts.  seq_u8: the full sequence as a uint8 ASCII numpy array, or a (pinned) torch uint8 tensor -- the tensor
    route uploads the rank's range straight from that buffer (no host-side gather: segment starts are 16-aligned)."""
    from . import device as dev

    slices, segs = segment_rows(a, b, e, L, seg_len=seg_len, halo=halo)
    if isinstance(seq_u8, torch.Tensor):
        host = seq_u8[a:e]
        src_off = np.array([lo - a for lo, _ in slices], dtype=np.int64)
        lens = np.array([hi - lo for lo, hi in slices], dtype=np.int64)
        return dev.SeqBatch(None, flat=(host, src_off, lens), segments=segs), segs
    return dev.SeqBatch([seq_u8[lo:hi] for lo, hi in slices], segments=segs), segs


def plan_long_sequence(seq_u8, model, frac_threshold, rcomp="both", seg_len=1 << 26, want_hits=True, chunk_len=None,
                       hit_cap=None, engine=None):
    """Upload + pack this rank's range of ONE long sequence and prepare its fused scan (config c4).  Returns
    (ScanPlan, SeqBatch); plan.launch() / plan.finish() may be called repeatedly (benchmarks)."""
    from . import device as dev

    rank, world = world_info()
    L = int(seq_u8.shape[0])
    halo = int(model.max_width) - 1
    a, b, e = shard_sequence(L, world, halo)[rank]
    batch, segs = long_sequence_batch(seq_u8, L, a, b, e, halo, seg_len)
    n = len(segs)
    out_rows = np.stack([np.zeros(n), np.ones(n)], axis=1).astype(np.int32)
    plan = dev.ScanPlan(model.device_set, batch, model.thresholds(frac_threshold), dev.STRANDS[rcomp], out_rows, 2,
                        want_hits=want_hits, want_counts=True, adjudicate=model.adjudicate, chunk_len=chunk_len,
                        hit_cap=hit_cap, engine=engine)
    return plan, batch


def reduce_long_counts(plan):
    """(segments, M, 2) per-rank count table -> global (1, M, 2) table: ONE all_reduce over NVLink."""
    counts = plan.counts.sum(dim=0, keepdim=True).to(torch.int32)
    return allreduce_counts(counts)


def scan_long_sequence(seq_u8, model, frac_threshold, rcomp="both", seg_len=1 << 26, want_hits=True, chunk_len=None,
                       hit_cap=None, sort=True):
    """Config-c4 style scan of ONE long sequence sharded over the ranks of the default process group.

    seq_u8: the full sequence as a uint8 ASCII array or pinned torch uint8 tensor (every rank passes the same buffer,
    or at least its own range).  Each rank uploads only its range + halo, scans it (both strands fused; RC coordinates
    use the global length), and the count table is all-reduced.  Returns (ScanResult with rank-local hits -- sorted,
    keys in global coordinates --, global counts tensor (1, M, 2))."""
    plan, _ = plan_long_sequence(seq_u8, model, frac_threshold, rcomp, seg_len, want_hits, chunk_len, hit_cap)
    plan.launch()
    res = plan.finish(sort=sort)
    model.last_scan_stats = res.stats
    return res, reduce_long_counts(plan)
