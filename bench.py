#!/usr/bin/env python
"""Benchmark of the PWM-scan hot path (BASELINE.json metric: Gbase x motif / s, both strands).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c2]

One "step" = one pass of the hot path over one batch of synthetic input.  Default workload (N=1): BASELINE config
c2 -- 1,000 synthetic virus genomes (10 Mb total) x 1,000 PWMs (width 7-12), both strands, threshold 0.7 of the
maximum score, site list + per-genome count table.  At N>1 every rank scans its own c2-sized shard (weak scaling;
sequences are independent units, no data-path collective) and the count tables are exchanged with ONE NCCL
all-gather per step.

Timed regions
  value : packed sequences already resident in HBM; per step = reset counters/counts, fused scan kernels,
          fp64 adjudicator, radix sort of the hit list into the reference's (row, pos, motif) order
          (+ the count-table all-gather at N>1).  CUDA events on the launching stream, per step; L2 is flushed
          (untimed) between steps; max over ranks.
  e2e   : the same work through the public device API from HOST buffers: H2D of the ASCII genomes from pinned
          memory, 2-bit pack kernel, scan, adjudicate, sort, D2H of the count table and counters.  The sorted site
          list stays on the device (config c2's result is the per-genome count table); wall clock around
          synchronised steps.
  roofline : the dominant side of the scan timed live with CUDA events.  Hybrid engine (default): the tensor-core
          prefilter + exact second pass (smg_scan_tc) -- bound "tensor", algorithmic FLOP = 2 strands x 2 x 4*W per
          base x motif against the measured dense bf16 peak; the register-table side and the whole-scan lookup-add
          rate against the measured FP32 add peak are reported beside it.  Engine regtab: bound "fp32_pipe".
  cpu_baseline / --impl reference : the reference pipeline's CPU port (oracle/ref_port.py) on a bounded sample.
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
LDS_PEAK_T = 9.31        # T lookup/s, 148 SMs x 32 lanes/clk x 1.965 GHz
THRESHOLD = 0.7


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


def build_workload(rank):
    from smeagol_b200 import synth

    pw = synth.synth_pwms(1000, 7, 12)
    genomes = synth.genome_set(1000, 10_000_000, 5000, 15000, seed=2 + 1000 * rank)
    return pw, genomes


def cpu_reference_run(pw, genomes, n_sample, steps=1, warmup=0):
    """The reference pipeline's CPU port on the first n_sample genomes x all PWMs.  Returns (value, details)."""
    import torch

    from oracle import ref_port

    model = ref_port.PortModel(list(pw.Matrix_id), list(pw.weights))
    seqs = [g.tobytes().decode("latin-1") for g in genomes[:n_sample]]
    ids = ["genome_%04d" % i for i in range(n_sample)]
    units = sum(len(s) for s in seqs) * len(pw)
    times, split = [], {}
    for it in range(warmup + steps):
        tm = {}
        t0 = time.perf_counter()
        out = ref_port.scan_sequences(seqs, ids, ids, model, THRESHOLD, sense="+", rcomp="both",
                                      outputs=("sites", "counts"), score=True, timings=tm)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            split = tm
    t = float(np.mean(times))
    return units / t / 1e9, {"seconds_per_step": t, "units_per_step": units, "cores": torch.get_num_threads(),
                             "n_sites": int(len(out.get("sites", []))), "stage_split_s": {k: round(v, 3) for k, v in split.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=120, help="genomes in the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine", default=None, choices=["regtab", "hmma"], help="scan engine (default: the library default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    workload = "c2: 1000 synthetic virus genomes (10 Mb) x 1000 PWMs (W 7-12), both strands, thr 0.7, sites + per-genome counts"

    if args.impl == "reference":
        if rank != 0:
            return 0
        pw, genomes = build_workload(0)
        n_sample = min(args.cpu_sample, 24)
        value, det = cpu_reference_run(pw, genomes, n_sample, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        sample = "first %d of 1000 genomes (%d bases) x 1000 PWMs per step" % (n_sample, det["units_per_step"] // 1000)
        line = {"metric": METRIC, "value": value, "unit": "Gbase*motif/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": det["seconds_per_step"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {"workload": workload, "sample": sample},
                "cpu_baseline": {"value": value, "unit": "Gbase*motif/s", "cores": det["cores"], "kind": "port",
                                 "sample": sample, "stage_split_s": det["stage_split_s"]},
                "e2e": {"value": value, "unit": "Gbase*motif/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    from smeagol_b200 import _lib
    from smeagol_b200 import device as dev
    from smeagol_b200.models import PWMModel

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks, peaks_kind = load_peaks()

    pw, genomes = build_workload(rank)
    model = PWMModel(pw)
    dset = model.device_set
    M = dset.n_motifs
    n = len(genomes)
    total_bases = int(sum(g.shape[0] for g in genomes))
    units = total_bases * M
    lookups = total_bases * 2 * int(dset.widths.astype(np.int64).sum())  # 2*W_m lookup-adds per base x motif
    thr = model.thresholds(THRESHOLD)
    out_rows = np.stack([2 * np.arange(n), 2 * np.arange(n) + 1], axis=1).astype(np.int32)  # group_by=None row order

    host_flat, src_off = dev.flatten_rows(genomes, pinned=True)
    lens = np.array([g.shape[0] for g in genomes], dtype=np.int64)
    batch = dev.SeqBatch(None, flat=(host_flat, src_off, lens))
    plan = dev.ScanPlan(dset, batch, thr, 3, out_rows, 2 * n, want_hits=True, want_counts=True, adjudicate=True,
                        engine=args.engine)
    plan.launch()
    res0 = plan.finish(sort=True)  # calibrates the hit buffer; deterministic workload -> n_hits is fixed
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
        if ev:
            ev[1].record()
        if plan.engine == "auto":  # hybrid: time the two sides of the width split separately
            plan.launch_regtab_side()
            if ev:
                ev[4].record()
            plan.launch_tc_side()
        else:
            plan.launch_scan()
            if ev:
                ev[4].record()
        if ev:
            ev[2].record()
        if plan.engine == "regtab":
            plan.launch_adjudicate()
        dev.sort_hits(plan.keys, plan.scores, n_hits, plan.key_bits, keys_out, scores_out, sort_tmp)
        if world > 1:
            dist.all_gather_into_tensor(gathered, plan.counts)
        if ev:
            ev[3].record()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.t_start = time.perf_counter()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    launches0 = _lib.launch_count()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush_buf.zero_()  # L2 flush, outside the event-timed region
        step(evs[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() - launches0
    step_ms = np.array([e[0].elapsed_time(e[3]) for e in evs])
    scan_ms = np.array([e[1].elapsed_time(e[2]) for e in evs])
    side_a_ms = np.array([e[1].elapsed_time(e[4]) for e in evs])  # register-table side (hybrid) / the whole scan
    side_b_ms = np.array([e[4].elapsed_time(e[2]) for e in evs])  # tensor-core side (hybrid) / 0
    c = plan.counters.cpu().numpy()
    assert int(c[_lib.CTR_HITS_APPENDED]) == n_hits, "non-deterministic hit count"
    ms_per_step = float(step_ms.mean())
    if world > 1:
        tt = torch.tensor([ms_per_step], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_per_step = float(tt.item())
    value = units * world / (ms_per_step * 1e-3) / 1e9

    # ---- end to end from host buffers
    e2e_times = []
    d2h_bytes = plan.counts.numel() * 4 + plan.counters.numel() * 8
    counts_host = torch.empty(tuple(plan.counts.shape), dtype=torch.int32, pin_memory=True)
    for i in range(2 + min(args.steps, 5)):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        b2 = dev.SeqBatch(None, flat=(host_flat, src_off, lens))  # H2D + pack kernel (+ status readback)
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
    e2e_t = float(np.mean(e2e_times))
    if world > 1:
        tt = torch.tensor([e2e_t], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_t = float(tt.item())
    e2e_value = units * world / e2e_t / 1e9
    sampler.t_end = time.perf_counter()
    sampler.stop()
    sampler.join(timeout=2)
    clocks = sampler.summary()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    scan_s = float(scan_ms.mean()) * 1e-3
    achieved_t = lookups / scan_s / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_smg_scan")
        except Exception:
            traffic = None
    algo_bytes = batch.total_words * 6 + n_hits * 12 + plan.counts.numel() * 4
    widths64 = dset.widths.astype(np.int64)
    fp32_view = {"achieved": achieved_t, "peak": FP32_ADD_PEAK_T, "unit": "Tlookup-add/s", "frac": achieved_t / FP32_ADD_PEAK_T,
                 "scan_ms": scan_s * 1e3,
                 "note": "whole scan (both engines): 2*W lookup-adds per base x motif (SURVEY 8d) over the event-timed scan, against "
                         "the measured FP32 add peak that bounds the register-table kernel (profiles/microbench_fp32_r01.txt)",
                 "lds_gather_roofline": {"peak": LDS_PEAK_T, "frac": achieved_t / LDS_PEAK_T,
                                         "note": "BASELINE.md section 4 bound of a shared-memory gather; neither engine does "
                                                 "shared-memory lookups, so this fraction exceeds 1"}}
    hbm_view = {"algorithmic_bytes": int(algo_bytes), "achieved_gbs": algo_bytes / scan_s / 1e9, "peak_gbs": peaks["hbm_gbs"],
                "frac": algo_bytes / scan_s / 1e9 / peaks["hbm_gbs"], "peaks": peaks_kind}
    if plan.engine == "auto" and float(side_b_ms.mean()) > float(side_a_ms.mean()):
        # dominant kernel: the tensor-core prefilter (scan_hmma_kernel, one launch per K depth, concurrent)
        tc = widths64 >= dset.tc_min_width
        tc_s = float(side_b_ms.mean()) * 1e-3
        flops = total_bases * 2 * int((2 * 4 * widths64[tc]).sum())            # 2 strands x 2 flop x 4W per base x motif
        flops_padded = total_bases * 2 * int((2 * 16 * ((widths64[tc] + 3) // 4)).sum())  # K padded to 16-row slices
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        roofline = {"bound": "tensor", "achieved": flops / tc_s / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": flops / tc_s / 1e12 / peak_tf, "traffic": traffic,
                    "kernel": "smg::scan_hmma_kernel + finalize_candidates_kernel (smg_scan_tc: %d PWMs of width >= %d)" % (
                        int(tc.sum()), dset.tc_min_width),
                    "kernel_ms": tc_s * 1e3, "kernel_share_of_step": tc_s * 1e3 / float(step_ms.mean()),
                    "algorithmic": "2 strands x 2 flop x 4*W per base x motif (one-hot contraction, SURVEY 8d); padded to "
                                   "16-row K slices the kernel executes %.1f TFLOP/s" % (flops_padded / tc_s / 1e12),
                    "peak_source": "MEASURED_PEAKS.json dense bf16 (%s, %s); the kernel issues warp-level mma.sync (fp16), whose "
                                   "own measured ceiling on this pool is 553 TFLOP/s (DESIGN 4.4): frac of that = %.3f" % (
                                       "sustained" if "bf16_tflops_sustained" in peaks else "burst", peaks_kind,
                                       flops_padded / tc_s / 1e12 / 553.0),
                    "other_side": {"kernel": "smg::scan_regtab_kernel + adjudicate_kernel (smg_scan: %d PWMs narrower than %d)" % (
                        int((~tc).sum()), dset.tc_min_width), "kernel_ms": float(side_a_ms.mean()),
                        "Tlookup_add_per_s": total_bases * 2 * int(widths64[~tc].sum()) / (float(side_a_ms.mean()) * 1e-3) / 1e12,
                        "frac_of_fp32_add_peak": total_bases * 2 * int(widths64[~tc].sum()) / (float(side_a_ms.mean()) * 1e-3) / 1e12 / FP32_ADD_PEAK_T},
                    "fp32_pipe": fp32_view, "hbm": hbm_view}
    else:
        roofline = {"bound": "fp32_pipe", "achieved": achieved_t, "peak": FP32_ADD_PEAK_T, "unit": "Tlookup-add/s",
                    "frac": achieved_t / FP32_ADD_PEAK_T, "traffic": traffic,
                    "kernel": "smg::scan_regtab_kernel (%d launches per step, one per width bucket)" % dset.n_buckets,
                    "kernel_ms": scan_s * 1e3, "kernel_share_of_step": scan_s * 1e3 / float(step_ms.mean()),
                    "peak_source": "FADD2 microbenchmark on this pool (profiles/microbench_fp32_r01.txt); nominal 148*128*1.965e9=37.2",
                    "lds_gather_roofline": fp32_view["lds_gather_roofline"], "hbm": hbm_view}
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        v, det = cpu_reference_run(pw, genomes, args.cpu_sample)
        cpu_baseline = {"value": v, "unit": "Gbase*motif/s", "cores": det["cores"], "kind": "port",
                        "sample": "first %d of 1000 genomes (%d bases) x 1000 PWMs, once (%.1f s)" % (
                            args.cpu_sample, det["units_per_step"] // 1000, det["seconds_per_step"]),
                        "stage_split_s": det["stage_split_s"]}
    line = {"metric": METRIC, "value": value, "unit": "Gbase*motif/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "per_gpu": "one c2 shard per rank (seed 2+1000*rank)",
                       "l2": "flushed between steps (256 MiB memset, untimed); timing = CUDA events per step",
                       "n_hits_per_gpu": n_hits, "hit_density": n_hits / (2.0 * units),
                       "near_threshold_adjudicated": int(c[_lib.CTR_NEAR]), "chunk_len": plan.chunk_len,
                       "lane_groups": dset.n_groups, "width_buckets": dset.n_buckets,
                       "collective": "all_gather(count tables) per step" if world > 1 else "none"},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "Gbase*motif/s", "h2d_bytes_per_step": int(batch.h2d_bytes + batch.n_rows * 40),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": e2e_t * 1e3,
                    "note": "H2D ASCII genomes (pinned) + pack + scan + adjudicate + sort + D2H count table; sorted site list stays on device"},
            "gpu_launches": int(launches), "clocks": clocks,
            "wall_s_timed_region": t_wall}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
