#!/usr/bin/env python
"""Benchmark of the PWM-scan hot path (BASELINE.json metric: Gbase x motif / s, both strands).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c4|c2] [--scale F]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...          (N > 1)

Workload (default, every N): BASELINE config c4, the one north_star's scaling target is quoted on -- ONE 248 Mb
synthetic chromosome x 1,000 JASPAR-sized PWMs (W 6-24, mean 11.8), both strands, threshold 0.7 of the maximum
score.  STRONG scaling: the same total work at every N; rank r scans a contiguous range of window starts plus a
(W_max - 1)-base right halo (smeagol_b200.dist.shard_sequence); hits stay rank-local, sorted, in global coordinates;
the single exchange is one NCCL all_reduce of the (1, M, 2) count table per step.  Config c2 (1,000 virus genomes x
1,000 PWMs, the configuration of round 1's bench line) rides along at N=1 under the key "c2"; `--config c2` runs it
alone (weak scaling: one shard per rank).

Timed regions
  value : the rank's packed range already resident in HBM; per step = reset counters/counts, the per-call index builds
          (k-mer index of the narrow PWMs, seed index of the wide ones), k-mer scan, seed scan + exact second pass (or
          the tcgen05 prefilter + second pass where the seed engine does not apply), fp64 adjudication, sort of the
          rank's hit list into the reference's (row, pos, motif) order, sum of the per-segment count tables and the
          all_reduce.  CUDA events per step on the launching stream; the working
          set (93 MB packed sequence + > 4 GB of hits at N=1) is larger than L2; max over ranks.
  e2e   : the same work through the public multi-GPU seam dist.scan_long_sequence() from a pinned HOST buffer: H2D
          of the rank's range, 2-bit pack kernel, scan, sort, all_reduce, then D2H of the count table and the hit
          counters.  The sorted site list stays on the device (3.9 GB at N=1: no reference-style DataFrame of 3e8
          rows exists on any host); wall clock around synchronised steps.
  roofline : the dominant kernel of the step, timed live with CUDA events around its C-ABI call: the k-mer scan
          (output-bound, against the HBM copy bandwidth) when both sides run on per-call indexes, else the tensor side
          (against the sustained bf16 peak); the step breakdown and SURVEY 8(d)'s lookup-add rate ride along.
  cpu_baseline / --impl reference : the reference pipeline's CPU port (oracle/ref_port.py: per-base Python encoding,
          dense conv, np.where, pandas) with all host threads on BASELINE.md section 3's c4 recipe: 100 kb chunks
          with a (W_max - 1)-base halo from a 5 Mb slice of the same chromosome (the reference cannot run c4 unchunked:
          its dense score tensor would be 2 TB), a bounded number of chunks per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gbase*motif/s (both strands)"
FP32_ADD_PEAK_T = 36.9   # T add/s, FADD2 microbenchmark on this pool's B200 (profiles/microbench_fp32_r01.txt)
THRESHOLD = 0.7
C4_LEN = 248_000_000
C4_WORKLOAD = ("c4: 248 Mb synthetic chromosome x 1000 JASPAR-sized PWMs (W 6-24), both strands, thr 0.7, contiguous "
               "ranges + halo over the ranks, sorted rank-local site lists + all-reduced count table")
C2_WORKLOAD = "c2: 1000 synthetic virus genomes (10 Mb) x 1000 PWMs (W 7-12), both strands, thr 0.7, sites + per-genome counts"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe): one streaming
    `nvidia-smi -lms 100` process, samples time-stamped on arrival; only those inside [t_start, t_end] are used."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.samples, self.proc = gpu_index, [], None
        self.t_start, self.t_end = None, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.perf_counter(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        def isnum(x):
            return x.replace(".", "", 1).isdigit()
        inside = [s for t, s in self.samples if len(s) > 8 and (self.t_start is None or self.t_start <= t <= self.t_end)]
        if not inside:  # region shorter than one sampling period: take the closest sample
            inside = [s for t, s in self.samples if len(s) > 8][-1:]
        sm = [float(s[1]) for s in inside if isnum(s[1])]
        mx = [float(s[2]) for s in inside if isnum(s[2])]
        pw = [float(s[3]) for s in inside if isnum(s[3])]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in inside:
            for n, v in zip(names, s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workloads
def build_c4(scale=1.0):
    from smeagol_b200 import synth

    L = int(C4_LEN * scale) // 16 * 16
    pw = synth.synth_pwms(1000, widths=synth.jaspar_widths(1000, seed=4))
    return pw, synth.single_sequence(L, 4)


def build_c2(rank=0):
    from smeagol_b200 import synth

    return synth.synth_pwms(1000, 7, 12), synth.genome_set(1000, 10_000_000, 5000, 15000, seed=2 + 1000 * rank)


def config_of(name, scale):
    """Identical in both arms (the driver compares them)."""
    wl = C4_WORKLOAD if name == "c4" else C2_WORKLOAD
    if scale != 1.0:
        wl += " [scaled x%g]" % scale
    return {"workload": wl, "scaling": "strong" if name == "c4" else "weak",
            "l2": "inputs + hit lists larger than L2 (c4) / L2 flushed between steps (c2); timing = CUDA events per step"}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_threads():
    import torch

    n = os.cpu_count() or 1
    torch.set_num_threads(n)  # under torchrun OMP_NUM_THREADS=1 would otherwise pin the arm to one core
    return torch.get_num_threads()


def cpu_reference_c4(pw, seq, n_chunks, first_chunk=0, chunk=100_000, slice_len=5_000_000):
    """the reference pipeline's CPU port on `n_chunks` chunks of 100 kb (+ halo) of the 5 Mb
    slice at the start of the chromosome.  Returns (Gbase*motif/s, details)."""
    from oracle import ref_port

    cores = cpu_threads()
    model = ref_port.PortModel(list(pw.Matrix_id), list(pw.weights))
    halo = model.max_width - 1
    n_slice = max(1, min(slice_len, len(seq)) // chunk)
    seqs, ids, owned = [], [], 0
    for i in range(n_chunks):
        c = (first_chunk + i) % n_slice
        lo, hi = c * chunk, min(len(seq), (c + 1) * chunk + halo)
        seqs.append(seq[lo:hi].tobytes().decode("latin-1"))
        ids.append("chr_chunk_%03d" % c)
        owned += min(len(seq), (c + 1) * chunk) - lo
    tm = {}
    t0 = time.perf_counter()
    out = ref_port.scan_sequences(seqs, ids, ids, model, THRESHOLD, sense="+", rcomp="both", outputs=("sites", "counts"),
                                  score=True, timings=tm)
    dt = time.perf_counter() - t0
    units = owned * len(pw)
    return units / dt / 1e9, {"seconds": dt, "units": units, "cores": cores, "n_sites": int(len(out.get("sites", []))),
                              "stage_split_s": {k: round(v, 3) for k, v in tm.items()},
                              "sample": "%d chunk(s) of 100 kb + %d-base halo from the first 5 Mb of the chromosome x %d PWMs "
                                        "(%.2f Gbase*motif), sites + counts DataFrames" % (n_chunks, halo, len(pw), units / 1e9)}


def cpu_reference_c2(pw, genomes, n_sample):
    from oracle import ref_port

    cores = cpu_threads()
    model = ref_port.PortModel(list(pw.Matrix_id), list(pw.weights))
    seqs = [g.tobytes().decode("latin-1") for g in genomes[:n_sample]]
    ids = ["genome_%04d" % i for i in range(n_sample)]
    units = sum(len(s) for s in seqs) * len(pw)
    tm = {}
    t0 = time.perf_counter()
    out = ref_port.scan_sequences(seqs, ids, ids, model, THRESHOLD, sense="+", rcomp="both", outputs=("sites", "counts"),
                                  score=True, timings=tm)
    dt = time.perf_counter() - t0
    return units / dt / 1e9, {"seconds": dt, "units": units, "cores": cores, "n_sites": int(len(out.get("sites", []))),
                              "stage_split_s": {k: round(v, 3) for k, v in tm.items()},
                              "sample": "first %d of 1000 genomes (%d bases) x %d PWMs, sites + counts DataFrames" % (
                                  n_sample, units // len(pw), len(pw))}


def run_reference_arm(args):
    """The reference's own CPU implementation of the path (port, all host threads), a bounded sample per step."""
    times, det = [], None
    if args.config == "c4":
        pw, seq = build_c4(args.scale)
        per_step = max(1, args.cpu_chunks_per_step)
        for it in range(args.warmup + args.steps):
            v, det = cpu_reference_c4(pw, seq, per_step, first_chunk=it * per_step)
            if it >= args.warmup:
                times.append((det["seconds"], det["units"]))
    else:
        pw, genomes = build_c2(0)
        for it in range(args.warmup + args.steps):
            v, det = cpu_reference_c2(pw, genomes, min(args.cpu_sample, 24))
            if it >= args.warmup:
                times.append((det["seconds"], det["units"]))
    secs = float(np.mean([t for t, _ in times]))
    value = float(np.mean([u for _, u in times])) / secs / 1e9
    line = {"metric": METRIC, "value": value, "unit": "Gbase*motif/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.config == "c4" else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": config_of(args.config, args.scale),
            "cpu_baseline": {"value": value, "unit": "Gbase*motif/s", "cores": det["cores"], "kind": "port",
                             "sample": det["sample"] + " per step", "stage_split_s": det["stage_split_s"]},
            "e2e": {"value": value, "unit": "Gbase*motif/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm helpers
def allreduce_max(x, world, device):
    import torch
    import torch.distributed as dist

    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def load_traffic(key):
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def tensor_side_roofline(dset, total_bases, tc_ms, step_ms, peaks, peaks_kind, tc_kind, traffic):
    """Roofline of the tensor-core prefilter + exact second pass (the dominant side of the scan)."""
    w = dset.widths.astype(np.int64)
    tc = (w >= max(dset.tc_min_width, dset.kmer_max_width + 1)) & (w <= 32)
    flops = total_bases * 2 * int((2 * 4 * w[tc]).sum())                  # 2 strands x 2 flop x 4W per base x motif
    flops_padded = total_bases * 2 * int((2 * 16 * ((w[tc] + 3) // 4)).sum())
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    tc_s = tc_ms * 1e-3
    kern = ("smg::scan_tc5_kernel (tcgen05.mma + TMEM) + finalize_candidates_kernel" if tc_kind == "tc5"
            else "smg::scan_hmma_kernel (mma.sync) + finalize_candidates_kernel")
    return {"bound": "tensor", "achieved": flops / tc_s / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": flops / tc_s / 1e12 / peak_tf, "traffic": traffic,
            "kernel": "%s: %d PWMs of width >= %d" % (kern, int(tc.sum()), max(dset.tc_min_width, dset.kmer_max_width + 1)),
            "kernel_ms": tc_ms, "kernel_share_of_step": tc_ms / step_ms,
            "algorithmic": "2 strands x 2 flop x 4*W per base x motif (one-hot contraction, SURVEY 8d); with K padded to the "
                           "16-row slices the tensor cores execute: %.1f TFLOP/s" % (flops_padded / tc_s / 1e12),
            "peak_source": "MEASURED_PEAKS.json dense bf16 (%s, %s)" % (
                "sustained" if "bf16_tflops_sustained" in peaks else "burst", peaks_kind)}


def index_scan_roofline(total_bases, n_hits, scan_ms, step_ms, peaks, peaks_kind, traffic, kmer_w, n_pwms):
    """Roofline of smg::scan_kmer_kernel, the dominant kernel when both sides of the width split run on per-call k-mer
    indexes.  The kernel does no per-cell arithmetic: it is OUTPUT-bound -- per launch it must read
    the packed sequence (2 bits + 1 mask bit per base) and write 12 bytes per hit (64-bit key + fp32 score)."""
    alg_bytes = total_bases * 0.375 + n_hits * 12.0
    peak = float(peaks.get("hbm_gbs_sustained", peaks.get("hbm_gbs", 6459.0)))
    t = scan_ms * 1e-3
    return {"bound": "hbm", "achieved": alg_bytes / t / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_bytes / t / 1e9 / peak,
            "traffic": traffic,
            "kernel": "smg::scan_kmer_kernel: %d PWMs of width <= %d through the per-call k-mer index" % (n_pwms, kmer_w),
            "kernel_ms": scan_ms, "kernel_share_of_step": scan_ms / step_ms,
            "algorithmic": "bytes per launch = 0.375 B per base (2-bit codes + ambiguity mask) + 12 B per emitted hit (%d hits); "
                           "the kernel is bound by dependent L2 look-ups (bitmap -> list head -> entry) and the hit write-out, "
                           "not by arithmetic: see fp32_pipe_view for the rate in SURVEY 8(d)'s lookup-add units" % n_hits,
            "peak_source": "MEASURED_PEAKS.json HBM copy bandwidth (%s)" % peaks_kind,
            "other_configuration": "SMEAGOL_B200_SEED=0 runs the wide PWMs through the tcgen05 prefilter instead: its tensor side "
                                   "reaches 0.81 of the sustained bf16 peak (1,166 TFLOP/s algorithmic) but the c4 step takes 37.0 ms "
                                   "instead of 22.1 (profiles/r02_bench_c4_n1_tcgen05.json) -- a lower roofline fraction on a much "
                                   "smaller amount of work"}


# ------------------------------------------------------------------------------------------------ c4 (headline)
def run_c4(args, rank, world, local_rank, device, peaks, peaks_kind, warmup):
    import torch
    import torch.distributed as dist

    from smeagol_b200 import _lib
    from smeagol_b200 import device as dev
    from smeagol_b200 import dist as sdist
    from smeagol_b200.models import PWMModel

    pw, seq = build_c4(args.scale)
    L, M = int(seq.shape[0]), len(pw)
    model = PWMModel(pw)
    dset = model.device_set
    units = L * M
    host = torch.from_numpy(seq).pin_memory()

    # ---- device-resident plan (the rank's range packed once, untimed)
    plan, batch = sdist.plan_long_sequence(host, model, THRESHOLD, engine=args.engine)
    dset = plan.pwmset  # the sibling set of this plan's engine split (k-mer index / register tables / tensor cores)
    plan.launch()
    res0 = plan.finish(sort=False)  # calibrates the hit buffer; the workload is deterministic
    n_hits = res0.n_hits
    batch_own = batch.total_own
    n_kmer_hits = 0
    if plan.engine == "auto" and plan.kmer_w > 0:  # hits written by the k-mer scan alone (its algorithmic output bytes)
        plan.counters.zero_()
        plan.counts.zero_()
        plan.launch_kmer_side()
        n_kmer_hits = int(plan.counters.cpu().numpy()[_lib.CTR_HITS_APPENDED])
    keys_out = torch.empty(max(n_hits, 1), dtype=torch.int64, device=device)
    scores_out = torch.empty(max(n_hits, 1), dtype=torch.float32, device=device)
    sort_tmp = torch.empty(16, dtype=torch.uint8, device=device)
    gcounts = torch.zeros((1, M, 2), dtype=torch.int32, device=device)

    def step(ev=None):
        nonlocal sort_tmp
        if ev:
            ev[0].record()
        plan.counters.zero_()
        plan.counts.zero_()
        if plan.engine == "auto":
            plan.launch_kmer_build()  # per-call indexes of (PWM set, thresholds): inside the step
            if ev:
                ev[6].record()
            plan.launch_kmer_side()
            if ev:
                ev[5].record()
            plan.launch_regtab_side()
            if ev:
                ev[1].record()
            plan.launch_tc_side()
            if ev:
                ev[2].record()
        else:
            plan.launch()
            if ev:
                ev[6].record()
                ev[5].record()
                ev[1].record()
                ev[2].record()
        if n_hits:
            _, _, sort_tmp = dev.sort_hits(plan.keys, plan.scores, n_hits, plan.key_bits, keys_out, scores_out, sort_tmp,
                                           return_temp=True, motif_bits=plan.motif_bits, total_positions=2 * batch_own)
        if ev:
            ev[3].record()
        torch.sum(plan.counts, dim=0, keepdim=True, out=gcounts)
        if world > 1:
            dist.all_reduce(gcounts, op=dist.ReduceOp.SUM)
        if ev:
            ev[4].record()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.t_start = time.perf_counter()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(7)] for _ in range(args.steps)]
    launches0 = _lib.launch_count()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        step(evs[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() - launches0
    step_ms = float(np.mean([e[0].elapsed_time(e[4]) for e in evs]))
    kmer_ms = float(np.mean([e[0].elapsed_time(e[5]) for e in evs]))
    build_ms = float(np.mean([e[0].elapsed_time(e[6]) for e in evs]))
    kmer_scan_ms = kmer_ms - build_ms
    regtab_ms = float(np.mean([e[5].elapsed_time(e[1]) for e in evs]))
    tc_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    sort_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
    coll_ms = float(np.mean([e[3].elapsed_time(e[4]) for e in evs]))
    c = plan.counters.cpu().numpy()
    assert int(c[_lib.CTR_HITS_APPENDED]) == n_hits, "non-deterministic hit count"
    ms_per_step = allreduce_max(step_ms, world, device)
    value = units / (ms_per_step * 1e-3) / 1e9

    # ---- global result check values (identical at every N): hits, key checksum, count-table checksum
    ksum = int(keys_out[:n_hits].sum().item()) if n_hits else 0
    glob = torch.tensor([n_hits, ksum & ((1 << 62) - 1)], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(glob, op=dist.ReduceOp.SUM)
    sorted_ok = bool((keys_out[1:n_hits] >= keys_out[:n_hits - 1]).all().item()) if n_hits > 1 else True
    count_sum = int(gcounts.sum().item())
    count_chk = int((gcounts.view(-1).to(torch.int64) * torch.arange(1, 2 * M + 1, device=device)).sum().item())
    del keys_out, scores_out, sort_tmp

    # ---- end to end through the public seam, from the pinned host buffer
    e2e_times = []
    counts_host = torch.empty((1, M, 2), dtype=torch.int32, pin_memory=True)
    h2d = int(batch.h2d_bytes + batch.n_rows * 40)
    plan_hit_cap = plan.hit_cap
    plan_kmer_w = plan.kmer_w
    plan_n_kmer_pwms = int((dset.widths <= plan.kmer_w).sum())
    del plan, batch
    torch.cuda.empty_cache()
    n_e2e = max(1, min(args.steps, 3))
    for i in range(1 + n_e2e):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res, cnt = sdist.scan_long_sequence(host, model, THRESHOLD, hit_cap=plan_hit_cap)
        counts_host.copy_(cnt, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= 1:
            e2e_times.append(dt)
        assert res.n_hits == n_hits
        del res, cnt
    e2e_t = allreduce_max(float(np.mean(e2e_times)), world, device)
    assert int(counts_host.sum()) == count_sum
    sampler.t_end = time.perf_counter()
    sampler.stop()
    sampler.join(timeout=2)
    clocks = sampler.summary()

    line = None
    if rank == 0:
        tc_kind = model.last_scan_stats.get("tc_kind") or "hmma"
        total_hits = int(glob[0].item())
        traffic = load_traffic("c4_n1_dominant_kernel_dram_bytes") if (world == 1 and args.scale == 1.0) else None
        sort_bytes = load_traffic("c4_n1_sort_dram_bytes") if (world == 1 and args.scale == 1.0) else None
        my_bases = L // world
        if tc_kind == "seed":
            roofline = index_scan_roofline(my_bases, n_kmer_hits, kmer_scan_ms, step_ms, peaks, peaks_kind,
                                           load_traffic("c4_n1_kmer_scan_dram_bytes") if (world == 1 and args.scale == 1.0) else None,
                                           plan_kmer_w, plan_n_kmer_pwms)
        else:
            roofline = tensor_side_roofline(dset, my_bases, tc_ms, step_ms, peaks, peaks_kind, tc_kind, traffic)
        w = dset.widths.astype(np.int64)
        lookups = my_bases * 2 * int(w.sum())
        roofline["step_breakdown_ms"] = {"index_builds": build_ms, "kmer_scan": kmer_scan_ms, "register_table_side": regtab_ms,
                                         ("seed_side" if tc_kind == "seed" else "tensor_core_side"): tc_ms,
                                         "sort": sort_ms, "count_sum_allreduce": coll_ms}
        roofline["sort_dram_bytes"] = sort_bytes
        roofline["algorithmic_dram_bytes"] = int(my_bases * 0.375 + n_hits * 12 + M * 8)
        scan_ms = kmer_ms + regtab_ms + tc_ms
        roofline["fp32_pipe_view"] = {"Tlookup_add_per_s": lookups / (scan_ms * 1e-3) / 1e12, "peak": FP32_ADD_PEAK_T,
                                      "frac": lookups / (scan_ms * 1e-3) / 1e12 / FP32_ADD_PEAK_T,
                                      "note": "2*W lookup-adds per base x motif (SURVEY 8d) over the scan kernels of rank 0, against "
                                              "the measured FP32 add peak; the prefilter skips most lookups, so > 1 is possible"}
        line = {"metric": METRIC, "value": value, "unit": "Gbase*motif/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of("c4", args.scale),
                "details": {"bases": L, "pwms": M, "hits_global": total_hits, "hit_density": total_hits / (2.0 * units),
                            "key_checksum_global_mod_2p62": int(glob[1].item()) & ((1 << 62) - 1),
                            "count_table_sum": count_sum, "count_table_checksum": count_chk,
                            "rank0_hits": n_hits, "rank0_sorted": sorted_ok, "engine": model.last_scan_stats.get("engine"),
                            "tc_kind": tc_kind, "tc_min_width": dset.tc_min_width,
                            "kmer_max_width": model.last_scan_stats.get("kmer_max_width"),
                            "kmer_index_entries": model.last_scan_stats.get("kmer_index_entries"),
                            "near_threshold_adjudicated_rank0": int(c[_lib.CTR_NEAR] + c[_lib.CTR_NEAR_TC]),
                            "collective": "all_reduce(count table (1, M, 2) int32) per step" if world > 1 else "none"},
                "roofline": roofline,
                "e2e": {"value": units / e2e_t / 1e9, "unit": "Gbase*motif/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": int(counts_host.numel() * 4 + 64), "ms_per_step": e2e_t * 1e3,
                        "note": "dist.scan_long_sequence(pinned host uint8): H2D of the rank's range + pack + scan + adjudicate + "
                                "sort + all_reduce + D2H of the count table and counters; the sorted site list stays on the device"},
                "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall}
    return line, (pw, seq)


# ------------------------------------------------------------------------------------------------ c2 (rider / --config c2)
def run_c2(args, rank, world, local_rank, device, peaks, peaks_kind, warmup, steps, with_frames):
    import torch
    import torch.distributed as dist

    from smeagol_b200 import _lib
    from smeagol_b200 import device as dev
    from smeagol_b200.models import PWMModel

    pw, genomes = build_c2(rank)
    model = PWMModel(pw)
    dset = model.device_set
    M, n = dset.n_motifs, len(genomes)
    total_bases = int(sum(g.shape[0] for g in genomes))
    units = total_bases * M
    thr = model.thresholds(THRESHOLD)
    out_rows = np.stack([2 * np.arange(n), 2 * np.arange(n) + 1], axis=1).astype(np.int32)  # group_by=None row order
    host_flat, src_off = dev.flatten_rows(genomes, pinned=True)
    lens = np.array([g.shape[0] for g in genomes], dtype=np.int64)
    batch = dev.SeqBatch(None, flat=(host_flat, src_off, lens))
    plan = dev.ScanPlan(dset, batch, thr, 3, out_rows, 2 * n, want_hits=True, want_counts=True, adjudicate=True,
                        engine=args.engine)
    pset = plan.pwmset
    plan.launch()
    res0 = plan.finish(sort=True)
    n_hits = res0.n_hits
    keys_out = torch.empty(n_hits, dtype=torch.int64, device=device)
    scores_out = torch.empty(n_hits, dtype=torch.float32, device=device)
    sort_tmp = torch.empty(64 << 20, dtype=torch.uint8, device=device)
    gathered = (torch.empty((world * plan.counts.shape[0],) + tuple(plan.counts.shape[1:]), dtype=torch.int32, device=device)
                if world > 1 else None)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def step(ev=None):
        if ev:
            ev[0].record()
        plan.counters.zero_()
        plan.counts.zero_()
        if plan.engine == "auto":
            plan.launch_kmer_build()  # per-call index of (PWM set, thresholds): inside the step
            plan.launch_kmer_side()
            plan.launch_regtab_side()
            if ev:
                ev[1].record()
            plan.launch_tc_side()
        else:
            plan.launch()
            if ev:
                ev[1].record()
        if ev:
            ev[2].record()
        dev.sort_hits(plan.keys, plan.scores, n_hits, plan.key_bits, keys_out, scores_out, sort_tmp, motif_bits=plan.motif_bits,
                      total_positions=2 * total_bases)
        if world > 1:
            dist.all_gather_into_tensor(gathered, plan.counts)
        if ev:
            ev[3].record()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    launches0 = _lib.launch_count()
    for i in range(steps):
        flush_buf.zero_()  # L2 flush, outside the event-timed region
        step(evs[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = _lib.launch_count() - launches0
    step_ms = float(np.mean([e[0].elapsed_time(e[3]) for e in evs]))
    regtab_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    tc_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    sort_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
    ms_per_step = allreduce_max(step_ms, world, device)
    value = units * world / (ms_per_step * 1e-3) / 1e9
    stats = res0.stats
    del keys_out, scores_out, sort_tmp, flush_buf

    # ---- end to end from host buffers through the device API: H2D + pack + scan + sort + D2H of the count table
    e2e_times = []
    counts_host = torch.empty(tuple(plan.counts.shape), dtype=torch.int32, pin_memory=True)
    for i in range(2 + min(steps, 5)):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        b2 = dev.SeqBatch(None, flat=(host_flat, src_off, lens))
        p2 = dev.ScanPlan(dset, b2, thr, 3, out_rows, 2 * n, hit_cap=plan.hit_cap, near_cap=plan.near_cap,
                          chunk_len=plan.chunk_len, engine=args.engine)
        p2.launch()
        r2 = p2.finish(sort=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered, p2.counts)
        counts_host.copy_(p2.counts, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= 2:
            e2e_times.append(dt)
        assert r2.n_hits == n_hits
    e2e_t = allreduce_max(float(np.mean(e2e_times)), world, device)
    out = {"workload": C2_WORKLOAD, "value": value, "unit": "Gbase*motif/s", "ms_per_step": ms_per_step, "steps": steps,
           "scaling": "weak", "n_hits_per_gpu": n_hits, "hit_density": n_hits / (2.0 * units), "gpu_launches": int(launches),
           "step_breakdown_ms": {"index_builds_kmer_scan_and_register_table_side": regtab_ms,
                                 ("seed_side" if stats.get("tc_kind") == "seed" else "tensor_core_side"): tc_ms,
                                 "sort_and_collective": sort_ms},
           "engine": stats.get("engine"), "tc_kind": stats.get("tc_kind"), "kmer_max_width": stats.get("kmer_max_width"),
           "e2e_device_api": {"value": units * world / e2e_t / 1e9, "ms_per_step": e2e_t * 1e3,
                              "h2d_bytes_per_step": int(batch.h2d_bytes + batch.n_rows * 40),
                              "d2h_bytes_per_step": int(plan.counts.numel() * 4 + 64),
                              "note": "H2D ASCII genomes (pinned) + pack + scan + sort + D2H count table; site list stays on device"}}
    if rank == 0:
        if stats.get("tc_kind") == "seed":
            # index builds + k-mer scan are timed together here (no event between them); hits of the narrow PWMs ~ all hits
            out["roofline"] = index_scan_roofline(total_bases, n_hits, regtab_ms, step_ms, peaks, peaks_kind,
                                                  load_traffic("c2_kmer_scan_dram_bytes"), int(stats.get("kmer_max_width") or 0),
                                                  int((pset.widths <= int(stats.get("kmer_max_width") or 0)).sum()))
            out["roofline"]["note"] = "kernel_ms here includes the per-call index builds of both engines"
        else:
            out["roofline"] = tensor_side_roofline(pset, total_bases, tc_ms, step_ms, peaks, peaks_kind,
                                                   stats.get("tc_kind") or "hmma", load_traffic("c2_dominant_kernel_dram_bytes"))
    if with_frames and rank == 0:
        # the PUBLIC call, like for like with the reference arm: records in, `counts` (and `sites`) DataFrames out
        import smeagol_b200 as sb

        recs = [sb.SeqRecord(sb.Seq(g.tobytes().decode("latin-1")), id="genome_%04d" % i, name="genome_%04d" % i)
                for i, g in enumerate(genomes)]
        tms = {}
        for outputs in (["counts"], ["sites", "counts"]):
            ts = []
            for i in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                o = sb.scan.scan_sequences(recs, model, THRESHOLD, sense="+", rcomp="both", outputs=outputs, score=True)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            tms["+".join(outputs)] = {"seconds": min(ts), "value": units / min(ts) / 1e9,
                                      "rows": {k: int(len(v)) for k, v in o.items()}}
            del o
        out["e2e_public_api"] = {"call": "smeagol_b200.scan.scan_sequences(1000 SeqRecords, model, 0.7, '+', 'both', outputs=..., score=True)",
                                 "unit": "Gbase*motif/s", "outputs": tms,
                                 "note": "wall clock incl. string encoding, H2D, pack, scan, sort, D2H and pandas DataFrame assembly "
                                         "(19 M-row sites frame); the reference arm builds the same frames"}
    return out, (pw, genomes)


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=["c4", "c2"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the c4 chromosome (testing only)")
    ap.add_argument("--cpu-chunks-per-step", type=int, default=2, help="reference arm, c4: 100 kb chunks per step")
    ap.add_argument("--cpu-chunks", type=int, default=10, help="in-run cpu_baseline, c4: 100 kb chunks (bounded sample)")
    ap.add_argument("--cpu-sample", type=int, default=120, help="c2: genomes in the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rider", action="store_true", help="skip the c2 rider at N=1")
    ap.add_argument("--engine", default=None, choices=["regtab", "hmma", "tc5"], help="scan engine (default: the library default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks, peaks_kind = load_peaks()

    if args.config == "c4":
        line, (pw, seq) = run_c4(args, rank, world, local_rank, device, peaks, peaks_kind, warmup)
        if rank == 0:
            if world == 1 and not args.no_cpu_baseline:
                v, det = cpu_reference_c4(pw, seq, args.cpu_chunks)
                line["cpu_baseline"] = {"value": v, "unit": "Gbase*motif/s", "cores": det["cores"], "kind": "port",
                                        "sample": det["sample"] + ", once (%.1f s)" % det["seconds"],
                                        "stage_split_s": det["stage_split_s"]}
            else:
                line["cpu_baseline"] = None
            if world == 1 and not args.no_rider:
                torch.cuda.empty_cache()
                line["c2"], _ = run_c2(args, rank, world, local_rank, device, peaks, peaks_kind, warmup, min(args.steps, 10), True)
            print(json.dumps(line))
    else:
        c2, (pw, genomes) = run_c2(args, rank, world, local_rank, device, peaks, peaks_kind, warmup, args.steps, world == 1)
        if rank == 0:
            line = {"metric": METRIC, "value": c2["value"], "unit": "Gbase*motif/s", "n_gpus": world, "steps": args.steps,
                    "warmup": warmup, "ms_per_step": c2["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of("c2", 1.0),
                    "details": {k: c2[k] for k in ("n_hits_per_gpu", "hit_density", "step_breakdown_ms", "engine", "tc_kind")},
                    "roofline": c2.get("roofline"),
                    "e2e": {"value": c2["e2e_device_api"]["value"], "unit": "Gbase*motif/s",
                            "h2d_bytes_per_step": c2["e2e_device_api"]["h2d_bytes_per_step"],
                            "d2h_bytes_per_step": c2["e2e_device_api"]["d2h_bytes_per_step"],
                            "ms_per_step": c2["e2e_device_api"]["ms_per_step"], "note": c2["e2e_device_api"]["note"]},
                    "e2e_public_api": c2.get("e2e_public_api"), "gpu_launches": c2["gpu_launches"], "cpu_baseline": None}
            if world == 1 and not args.no_cpu_baseline:
                v, det = cpu_reference_c2(pw, genomes, args.cpu_sample)
                line["cpu_baseline"] = {"value": v, "unit": "Gbase*motif/s", "cores": det["cores"], "kind": "port",
                                        "sample": det["sample"] + ", once (%.1f s)" % det["seconds"],
                                        "stage_split_s": det["stage_split_s"]}
            print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
