#!/usr/bin/env python
"""Benchmarks of the other BASELINE.json configurations (bench.py holds the headline config c2).

    python bench_configs.py --config c1|c3|c4|c5 [--steps K] [--warmup W]          (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... bench_configs.py --config c3|c4   (STRONG scaling)

c1  30 kb genome x 100 PWMs (W 7-12), both strands, thr 0.7 -- latency of the public scan_sequences() call
c3  30 kb genome x 500 PWMs over 1,000 seeded dinucleotide-preserving shuffles (smg_kmer_shuffle), counts only, backgrounds sharded
    across ranks, one all_gather of the count tables
c4  248 Mb sequence x 1,000 JASPAR-sized PWMs (W 6-24, mean 12), contiguous ranges + halo across ranks, hits stay
    rank-local (count + checksum), one all_reduce of the count table
c5  1,000,000 SNVs x 500 PWMs on a 10 Mb reference, batched max ref/alt window score (variant restatement)
One JSON line per run, same keys as bench.py where they apply.  Synthetic data, generators in smeagol_b200/synth.py.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
FP32_ADD_PEAK_T = 36.9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c1", "c3", "c4", "c5"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (testing)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from smeagol_b200 import _lib, synth
    from smeagol_b200 import device as dev
    from smeagol_b200 import dist as sdist
    from smeagol_b200.models import PWMModel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def timed(step, units_total, extra):
        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        l0 = _lib.launch_count()
        for e0, e1 in evs:
            e0.record()
            step()
            e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        per_step = [a.elapsed_time(b) for a, b in evs]
        ms = float(np.median(per_step))  # median: a step that hits an allocator growth is an outlier, not the rate
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        line = {"metric": "Gbase*motif/s (both strands)", "value": units_total / (ms * 1e-3) / 1e9, "unit": "Gbase*motif/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "gpu_launches": int(_lib.launch_count() - l0), "ms_per_step_all_rank0": [round(x, 3) for x in per_step]}
        line.update(extra)
        return line

    if args.config == "c1":
        import smeagol_b200 as sb

        pw = synth.synth_pwms(100, 7, 12)
        model = PWMModel(pw)
        seq = synth.single_sequence(30_000, 1)
        rec = [sb.SeqRecord(sb.Seq(seq.tobytes().decode()), id="virus", name="virus")]
        out = None
        times = []
        for i in range(args.warmup + args.steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = sb.scan.scan_sequences(rec, model, 0.7, sense="+", rcomp="both", outputs=["sites", "counts"], score=True)
            torch.cuda.synchronize()
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        t = float(np.median(times))
        units = 30_000 * 100
        line = {"metric": "Gbase*motif/s (both strands)", "value": units / t / 1e9, "unit": "Gbase*motif/s", "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "c1: 30 kb x 100 PWMs, both strands, thr 0.7; wall time of the public scan_sequences() call "
                                       "incl. H2D, pack, scan, sort, D2H and DataFrame assembly (latency-bound, SURVEY H5)",
                           "n_sites": int(len(out["sites"]))}}
        print(json.dumps(line))
        return 0

    if args.config == "c3":
        n_bg = int(1000 * args.scale)
        pw = synth.synth_pwms(500, 7, 12)
        model = PWMModel(pw)
        genome = synth.single_sequence(30_000, 3)
        mine = sdist.shard_rows([30_000] * n_bg, world)[rank]
        n = len(mine)
        # seeded DINUCLEOTIDE-preserving shuffles -- generated and packed on the device by smg_kmer_shuffle
        # (utils.shuffle_batch_device, simK = 2), this rank's share with its own seed; they never exist on the host
        import smeagol_b200 as sb
        from smeagol_b200 import utils as sutils

        rec = [sb.SeqRecord(sb.Seq(genome.tobytes().decode()), id="virus", name="virus")]
        _, _, batch, _ = sutils.shuffle_batch_device(rec, n, simK=2, seed=3 + 1000 * rank, device=device)
        out_rows = np.stack([np.arange(n), n + np.arange(n)], axis=1).astype(np.int32)  # one group, rows [fwd.., rc..]
        plan = dev.ScanPlan(model.device_set, batch, model.thresholds(0.7), 3, out_rows, 2 * n, want_hits=False,
                            want_counts=True)
        plan.launch()
        res = plan.finish()

        def step():
            plan.launch_kmer_build()  # the per-call k-mer index belongs to the step
            plan.launch()
            if world > 1:
                sdist.allgather_counts(plan.counts)
        units = n_bg * 30_000 * 500
        lookups = n_bg * 30_000 * 2 * int(model.device_set.widths.astype(np.int64).sum())
        line = timed(step, units, {"config": {"workload": "c3: 30 kb genome x 500 PWMs x %d seeded dinucleotide-preserving shuffles (device-generated), counts only (no site list), "
                                                            "backgrounds sharded over ranks, one all_gather of count tables" % n_bg,
                                              "hits_counted_rank0": res.stats["n_hits_total"]}})
        line["roofline"] = {"bound": "fp32_pipe", "achieved": lookups / (line["ms_per_step"] * 1e-3) / 1e12 / world,
                            "peak": FP32_ADD_PEAK_T, "unit": "Tlookup-add/s per GPU (whole step)"}
        line["roofline"]["frac"] = line["roofline"]["achieved"] / FP32_ADD_PEAK_T
        if rank == 0:
            print(json.dumps(line))
    elif args.config == "c4":
        L = int(248_000_000 * args.scale)
        M = 1000
        pw = synth.synth_pwms(M, widths=synth.jaspar_widths(M, seed=4))
        model = PWMModel(pw)
        seq = synth.single_sequence(L, 4)
        halo = int(model.max_width) - 1
        a, b, e = sdist.shard_sequence(L, world, halo)[rank]
        slices, segs = sdist.segment_rows(a, b, e, L, seg_len=1 << 26, halo=halo)
        t0 = time.perf_counter()
        batch = dev.SeqBatch([seq[lo:hi] for lo, hi in slices], segments=segs)
        torch.cuda.synchronize()
        t_upload = time.perf_counter() - t0
        n = len(segs)
        out_rows = np.stack([np.zeros(n), np.ones(n)], axis=1).astype(np.int32)
        plan = dev.ScanPlan(model.device_set, batch, model.thresholds(0.7), 3, out_rows, 2, want_hits=True, want_counts=True,
                            engine=os.environ.get("SMEAGOL_B200_ENGINE"))
        plan.launch()
        res = plan.finish(sort=False)
        total_counts = torch.zeros((1, M, 2), dtype=torch.int32, device=device)

        def step():
            plan.launch_kmer_build()  # the per-call k-mer index belongs to the step
            plan.launch()
            total_counts.copy_(plan.counts.sum(dim=0, keepdim=True))
            sdist.allreduce_counts(total_counts)
        units = L * M
        lookups = L * 2 * int(model.device_set.widths.astype(np.int64).sum())
        line = timed(step, units, {"config": {"workload": "c4: %d bp x 1000 JASPAR-sized PWMs (W 6-24, mean %.1f), both strands, thr 0.7, "
                                                            "contiguous ranges + %d-base halo per rank, hits rank-local (unsorted), "
                                                            "one all_reduce of the count table" % (L, float(model.device_set.widths.mean()), halo),
                                              "hits_rank0": res.n_hits, "upload_pack_s_rank0": t_upload,
                                              "hit_checksum_rank0": res.key_checksum()[0] % (1 << 61)}})
        line["roofline"] = {"bound": "fp32_pipe", "achieved": lookups / (line["ms_per_step"] * 1e-3) / 1e12 / world,
                            "peak": FP32_ADD_PEAK_T, "unit": "Tlookup-add/s per GPU (whole step)"}
        line["roofline"]["frac"] = line["roofline"]["achieved"] / FP32_ADD_PEAK_T
        if rank == 0:
            print(json.dumps(line))
    else:  # c5
        from smeagol_b200 import variant

        n_snv = int(1_000_000 * args.scale)
        pw = synth.synth_pwms(500, 7, 12)
        model = PWMModel(pw)
        seq = synth.single_sequence(10_000_000, 5)
        pos, alt = synth.snvs(seq, n_snv, 12, seed=5)
        batch = dev.SeqBatch([seq])
        rows = np.zeros(n_snv, dtype=np.int32)

        d_rows, d_pos, d_alt = (torch.from_numpy(x).to(device) for x in (rows, pos.astype(np.int32), alt))

        def step():  # SNVs resident in HBM: the kernels alone
            variant.variant_window_scores_on_batch(batch, d_rows, d_pos, d_alt, model, to_host=False)

        def step_e2e():  # (pos, alt) uploaded from host arrays every step
            variant.variant_window_scores_on_batch(batch, rows, pos, alt, model, to_host=False)
        w = model.device_set.widths.astype(np.int64)
        adds = n_snv * int((2 * (w * w + w)).sum())  # 2 strands x (W windows x W terms + W substitutions)
        adds_executed = n_snv * int((2 * (w * w + 2 * w)).sum())  # delta form: W^2 + 2 W packed (two-strand) operations
        line = timed(step, n_snv * 500, {"metric": "SNV*motif/s", "unit": "G SNV*motif/s",
                                         "config": {"workload": "c5: %d SNVs x 500 PWMs on a 10 Mb reference, max ref/alt window score over "
                                                                "covering windows and both strands (fp32, delta form); SNVs resident in HBM, "
                                                                "4 GB of scores written per step" % n_snv}})
        line["value"] = n_snv * 500 / (line["ms_per_step"] * 1e-3) / 1e9
        e2e = timed(step_e2e, n_snv * 500, {})
        line["e2e"] = {"value": n_snv * 500 / (e2e["ms_per_step"] * 1e-3) / 1e9, "unit": "G SNV*motif/s",
                       "h2d_bytes_per_step": int(rows.nbytes + 4 * len(pos) + alt.nbytes), "d2h_bytes_per_step": 0,
                       "note": "scores stay on the device (4 GB per step)"}
        line["roofline"] = {"bound": "fp32_pipe", "achieved": adds / (line["ms_per_step"] * 1e-3) / 1e12, "peak": FP32_ADD_PEAK_T,
                            "unit": "Tadd/s", "frac": adds / (line["ms_per_step"] * 1e-3) / 1e12 / FP32_ADD_PEAK_T,
                            "algorithmic": "2*(W^2 + W) adds per (SNV, PWM) (SURVEY 8d)",
                            "executed_Tadd_per_s": adds_executed / (line["ms_per_step"] * 1e-3) / 1e12}
        if rank == 0:
            print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
