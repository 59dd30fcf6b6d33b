"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: sharding plans, halo ownership and the single
count-table collective.  The per-rank "scan" is the CPU oracle here (test infrastructure); the GPU version of the
same invariance is tests/test_gpu_parity.py::test_segments_halo_ownership_equals_whole."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import scan_oracle as so
from smeagol_b200 import dist as sdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pwms(rng, n):
    ws = []
    for _ in range(n):
        W = int(rng.integers(5, 13))
        ws.append(np.log2((rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04 / 0.25))
    return ws


def test_shard_rows_balanced_and_complete():
    rng = np.random.default_rng(0)
    lens = rng.integers(5000, 15000, size=1000)
    for world in (1, 2, 4, 8):
        parts = sdist.shard_rows(lens, world)
        allidx = np.sort(np.concatenate(parts))
        np.testing.assert_array_equal(allidx, np.arange(1000))
        loads = np.array([lens[p].sum() for p in parts])
        assert loads.max() - loads.min() <= lens.max()
        for p in parts:
            assert np.all(np.diff(p) > 0)


def test_shard_sequence_covers_without_overlap():
    for L, world in ((248_000_000, 8), (1000, 3), (17, 4), (16, 2)):
        sh = sdist.shard_sequence(L, world, halo=23)
        assert sh[0][0] == 0 and sh[-1][1] == L
        for (a, b, e), (a2, b2, e2) in zip(sh[:-1], sh[1:]):
            assert b == a2 and a % 16 == 0 and e == min(L, b + 23)
        slices, segs = sdist.segment_rows(*sh[0], L, seg_len=100, halo=23)
        assert sum(s[2] for s in segs) == sh[0][1] - sh[0][0]
        for (lo, hi), (g_off, g_len, n_own) in zip(slices, segs):
            assert lo == g_off and g_len == L and hi == min(L, lo + n_own + 23)


def _worker(rank, world, port, seq, ws, frac, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = len(seq)
        halo = max(w.shape[0] for w in ws) - 1
        a, b, e = sdist.shard_sequence(L, world, halo)[rank]
        # the rank's scan, restated with the oracle on its slice: windows owned iff start in [a, b)
        res = so.scan_rows([seq], ws, frac, rcomp="both", order="kernel")
        widths = np.array([w.shape[0] for w in ws])
        fwd_start = np.where(res["row"] == 0, res["pos"], L - res["pos"] - widths[res["motif"]])
        own = (fwd_start >= a) & (fwd_start < b)
        # every owned window must fit inside the resident bases [a, e)
        assert np.all(fwd_start[own] + widths[res["motif"]][own] <= e)
        counts = torch.zeros((1, len(ws), 2), dtype=torch.int32)
        np.add.at(counts.numpy(), (0, res["motif"][own], res["row"][own]), 1)
        total = sdist.allreduce_counts(counts.clone())
        gathered = sdist.allgather_counts(counts)
        keys = torch.from_numpy(((res["row"][own] << 40) | (res["pos"][own] << 8) | res["motif"][own]).astype(np.int64))
        scores = torch.from_numpy(res["score"][own])
        gk, gs = sdist.gather_hits(keys, scores, len(keys))
        if rank == 0:
            q.put((total.numpy(), gathered.numpy(), gk.numpy(), int(own.sum())))
    finally:
        dist.destroy_process_group()


def test_world2_counts_allreduce_equals_whole():
    rng = np.random.default_rng(5)
    ws = _pwms(rng, 12)
    seq = "".join(rng.choice(list("ACGT"), size=3000))
    frac = 0.6
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = os.pathsep.join([root, os.path.join(root, "tests"), os.environ.get("PYTHONPATH", "")])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seq, ws, frac, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        total, gathered, gk, n0 = q.get(timeout=180)  # never block forever on a failed worker
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    whole = so.scan_rows([seq], ws, frac, rcomp="both", order="kernel")
    exp = so.count_table(whole, 2, len(ws)).T[None, :, :]  # (1, M, 2)
    np.testing.assert_array_equal(total, exp)
    np.testing.assert_array_equal(gathered.sum(axis=0), exp)
    assert gathered.shape == (2, 1, len(ws), 2)
    assert len(gk) == len(whole["row"]) and len(np.unique(gk)) == len(gk)  # no boundary hit lost or duplicated
