#!/usr/bin/env python3
"""bench.py -- trio loci/s of the read<->allele alignment hot path on B200 (BASELINE.json metric).

A step = one pass of the hot path (2-bit packing, tiered WFA alignment with on-device backtrace and
clipped scoring, fused per-locus reduction) over the whole synthetic genome-wide trio assigned to
this rank.  The workload is ONE catalog (configs[1]: 937k loci at 30x); with
--gpus N it is cut into N contiguous locus ranges balanced on sum(n + m) over the pairs of each locus
(trgt_denovo_b200.shard.balanced_bounds), one rank / ctx / GPU each, no data-path collective: strong scaling,
as north_star states it ("loci sharded across 1/2/4/8 B200").

  value     loci/s with the inputs already resident in HBM (tdb_assess_batch_device), timed with the
            CUDA events the library records on its launching stream, max over ranks
  e2e       the same pass through the host-buffer C-ABI call (tdb_trio_batch): pinned host inputs as ASCII
            (what the reference hands over), H2D + kernels + D2H of the per-allele results inside the timed region
  e2e_packed  the same through tdb_trio_batch_packed: the host packed to 2 bits at load time (outside the timed
            region, as a loader that reads BAM 4-bit bases would), so the call does no per-base CPU work
  roofline  the dominant alignment kernel against the INT32-issue roofline SURVEY.md 8(d) names: lane-int-ops of
            the pairs that launch aligned (ops = 24*C + 8*E + 4*T, counted exactly on the device and checked
            against the oracle in tests/) over its CUDA-event duration; the HBM figure is the nested note
  cpu_baseline / --impl reference
            the CPU oracle (scalar C restatement of the reference path, "port": the reference's own
            Rust+WFA2-lib binary cannot be built here) on all host cores, on a bounded sample
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

INT_OPS_PER_CELL, INT_OPS_PER_EXT, INT_OPS_PER_COL = 24, 8, 4


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ts, f in self.rows:
            if len(f) < 9 or not (t0 - 0.1 <= ts <= t1 + 0.3):
                continue
            try:
                sm.append(float(f[1])), smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_pass(tdo, batch, n_loci, clip, q, duo, threads):
    """align + reduce of loci [0, n_loci) of `batch` on the CPU oracle; returns seconds."""
    n_pairs = int(batch.meta[n_loci - 1].pair_end)
    t0 = time.perf_counter()
    scores, status, _ = tdo.align_batch(batch.blob, batch.seq_off, batch.seq_len, batch.pair_pattern[:n_pairs],
                                        batch.pair_text[:n_pairs], clip, n_threads=threads)
    tdo.assess_batch(batch.loci_ptr, n_loci, scores, q, duo=duo, n_threads=threads)
    return time.perf_counter() - t0, n_pairs


def workload_name(args, world=1, ref_sample=None):
    kind = "duo" if args.duo else "trio"
    w = (f"genome-wide synthetic {kind}, ONE catalog of {args.loci} loci at 30x (SURVEY.md 8d config "
         f"{'3' if args.duo else '2'}, seed 20250002)")
    if ref_sample is not None:
        return w + f"; the CPU arm times the first {ref_sample} loci of it per step (a rate: loci/s)"
    return w + (f", sharded over {world} GPUs in contiguous locus ranges balanced on sum(n+m)" if world > 1 else "") + \
        ", inputs larger than L2"


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on all host threads, rank 0 only.  Maps no product
    code: the workload generator is its own library (libtdb_synth.so) and build() is told not to load the rest."""
    if rank != 0:
        return 0
    import __graft_entry__ as entry
    entry.build(load=False)
    from importlib import import_module
    from oracle import tdo
    S = import_module("trgt_denovo_b200.synth")
    threads = os.cpu_count() or 1
    sample = min(args.loci, args.cpu_sample_loci)
    params = S.config("duo_genome" if args.duo else "trio_genome")
    batch = S.generate(params, 0, sample, threads)
    for _ in range(args.warmup):
        oracle_pass(tdo, batch, min(sample, 2000), args.clip, args.quantile, args.duo, threads)
    times, n_pairs = [], 0
    for _ in range(args.steps):
        dt, n_pairs = oracle_pass(tdo, batch, sample, args.clip, args.quantile, args.duo, threads)
        times.append(dt)
    tot = sum(times)
    val = sample * args.steps / tot
    line = {
        "impl": "reference", "metric": "trio_loci_per_sec" if not args.duo else "duo_loci_per_sec", "value": val,
        "unit": "loci/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_name(args, ref_sample=sample), "loci_per_step": sample,
                   "pairs_per_step": n_pairs, "clip_len": args.clip},
        "alignments_per_sec": n_pairs * args.steps / tot,
        "cpu_baseline": {"value": val, "unit": "loci/s", "cores": threads, "kind": "port",
                         "sample": f"first {sample} loci of the workload per step, {threads} threads; scalar C "
                                   "restatement of the reference path (not WFA2-lib, which cannot be built here)"},
        "e2e": {"value": val, "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--loci", type=int, default=int(os.environ.get("TDB_BENCH_LOCI", 937000)),
                    help="loci of the ONE catalog (sharded over the ranks)")
    ap.add_argument("--duo", action="store_true")
    ap.add_argument("--clip", type=int, default=50)
    ap.add_argument("--quantile", type=float, default=1.0)
    ap.add_argument("--cpu-sample-loci", type=int, default=int(os.environ.get("TDB_BENCH_CPU_LOCI", 150000)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-census", action="store_true", help="skip the heuristic-sensitivity census")
    ap.add_argument("--weak", action="store_true", help="every rank its own --loci catalog (round-1 behaviour)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    from importlib import import_module
    import trgt_denovo_b200 as tdb
    S = import_module("trgt_denovo_b200.synth")
    B = import_module("trgt_denovo_b200.binding")
    SH = import_module("trgt_denovo_b200.shard")
    if tdb.device_count() == 0:
        raise SystemExit("no CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = tdb.lib()
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()

    # ---- workload: ONE catalog, this rank's locus range --------------------------------------------------
    threads = max(1, (os.cpu_count() or 1) // max(1, world))
    try:  # about 26 KB of pinned host memory per locus (ASCII + packed + results): stay below 80 % of what is free
        import psutil
        fit = int(0.8 * psutil.virtual_memory().available / 26e3)
        if args.loci > fit:
            log(f"[rank {rank}] a catalog of {args.loci} loci does not fit the host memory; using {fit}")
            args.loci = max(1000 * world, fit)
    except ImportError:
        pass
    params = S.config("duo_genome" if args.duo else "trio_genome", pinned=1)
    t0 = time.time()
    if args.weak:
        lo, hi = rank * args.loci, (rank + 1) * args.loci
        bounds = [r * args.loci for r in range(world + 1)]
    elif world == 1:
        lo, hi = 0, args.loci
        bounds = [0, args.loci]
    else:
        # every rank weighs its even slice (sum over the locus' pairs of n + m), the weights are gathered (bookkeeping,
        # 8 bytes per locus) and every rank derives the same balanced boundaries
        eb = SH.even_bounds(args.loci, world)
        probe = S.generate(S.config("duo_genome" if args.duo else "trio_genome"), eb[rank], eb[rank + 1], threads)
        w_mine = SH.locus_weights_from_meta(probe)
        probe.close()
        longest = max(eb[r + 1] - eb[r] for r in range(world))
        mine = torch.zeros(longest, dtype=torch.int64, device=dev)
        mine[:len(w_mine)] = torch.from_numpy(w_mine).to(dev)
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        weights = np.concatenate([parts[r][:eb[r + 1] - eb[r]].cpu().numpy() for r in range(world)])
        bounds = SH.balanced_bounds(weights, world)
        lo, hi = bounds[rank], bounds[rank + 1]
    batch = S.generate(params, lo, hi, threads)
    log(f"[rank {rank}] loci [{lo}, {hi}): {batch.n_loci} loci, {batch.n_pairs} pairs, {batch.blob.size / 1e9:.2f} GB "
        f"of reads in {time.time() - t0:.1f}s")
    n_loci, n_pairs, n_seqs = batch.n_loci, batch.n_pairs, batch.n_seqs
    ctx = tdb.Context(device_id=local_rank, clip_len=args.clip)
    if world > 1:
        ctx.set_host_threads(max(1, threads - 1))  # the ranks share the host's cores
    h = ctx._h

    # ---- device-resident copies (plumbing via torch) ----------------------------------------------
    def dev_copy(arr, pad=0):
        t = torch.empty(arr.nbytes + pad, dtype=torch.uint8, device=dev)
        t[:arr.nbytes].copy_(torch.from_numpy(arr.view(np.uint8)), non_blocking=False)
        if pad:
            t[arr.nbytes:].zero_()
        return t
    loci_np = np.frombuffer((C.c_uint8 * (C.sizeof(B.TrioLocus) * n_loci)).from_address(batch.loci_ptr), dtype=np.uint8)
    d_blob, d_off, d_len = dev_copy(batch.blob, 64), dev_copy(batch.seq_off), dev_copy(batch.seq_len)
    d_pp, d_pt, d_loci = dev_copy(batch.pair_pattern), dev_copy(batch.pair_text), dev_copy(loci_np)
    d_out = torch.zeros(n_loci * 2 * C.sizeof(B.AlleleOut), dtype=torch.uint8, device=dev)
    d_score = torch.zeros(n_pairs, dtype=torch.int32, device=dev)
    d_mask = torch.zeros(n_pairs, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    # The generator appends sequences back to back, as a host filling one byte vector does: that is the ABI's
    # dense layout (seq_off = NULL, offsets derived on the device).  TDB_BENCH_DENSE=0 passes explicit offsets.
    ends = np.cumsum(batch.seq_len, dtype=np.uint64)
    dense = bool(n_seqs == 0 or (batch.seq_off[0] == 0 and np.array_equal(batch.seq_off[1:], ends[:-1])))
    del ends
    dense = dense and os.environ.get("TDB_BENCH_DENSE", "1") != "0"
    h_off_ptr = None if dense else batch.seq_off.ctypes.data
    d_off_ptr = None if dense else d_off.data_ptr()

    def step_device():
        r = L.tdb_assess_batch_device(h, int(args.duo), d_blob.data_ptr(), batch.blob.size, d_off_ptr,
                                      d_len.data_ptr(), n_seqs, d_pp.data_ptr(), d_pt.data_ptr(), n_pairs,
                                      d_loci.data_ptr(), n_loci, args.quantile, d_out.data_ptr(), d_score.data_ptr(),
                                      d_mask.data_ptr())
        if r != 0:
            raise RuntimeError(L.tdb_last_error(h).decode())
        return ctx.counters()

    # results land in pinned host memory too (tdb_host_alloc), like the inputs
    out_bytes = C.sizeof(B.AlleleOut) * 2 * n_loci
    h_out_ptr = L.tdb_host_alloc(out_bytes)
    h_mask_ptr = L.tdb_host_alloc(max(1, n_pairs))
    if not h_out_ptr or not h_mask_ptr:
        raise SystemExit("pinned host allocation failed")
    h_out = (B.AlleleOut * (2 * n_loci)).from_address(h_out_ptr)
    h_mask = np.frombuffer((C.c_uint8 * max(1, n_pairs)).from_address(h_mask_ptr), dtype=np.uint8)

    def step_host():
        r = L.tdb_trio_batch if not args.duo else L.tdb_duo_batch
        rc = r(h, batch.blob.ctypes.data, batch.blob.size, h_off_ptr, batch.seq_len.ctypes.data,
               n_seqs, batch.pair_pattern.ctypes.data, batch.pair_text.ctypes.data, n_pairs, batch.loci_ptr, n_loci,
               args.quantile, C.cast(h_out, C.c_void_p), None, h_mask.ctypes.data)
        if rc != 0:
            raise RuntimeError(L.tdb_last_error(h).decode())

    # packed at load time (outside every timed region): the 2-bit stream in pinned memory, exception list
    n_words = L.tdb_packed_words(batch.blob.size)
    h_packed_ptr = L.tdb_host_alloc(n_words * 4)
    if not h_packed_ptr:
        raise SystemExit("pinned host allocation failed")
    flag = np.zeros(max(1, n_seqs), dtype=np.uint8)
    t0 = time.time()
    rc = L.tdb_host_pack(batch.blob.ctypes.data, batch.blob.size, batch.seq_off.ctypes.data, batch.seq_len.ctypes.data,
                         n_seqs, h_packed_ptr, flag.ctypes.data, threads)
    if rc != 0:
        raise SystemExit("tdb_host_pack failed")
    load_pack_s = time.time() - t0
    exc = np.nonzero(flag[:n_seqs])[0].astype(np.uint32)
    exc_bytes = np.frombuffer(b"".join(batch.seq(int(i)) for i in exc), dtype=np.uint8)
    # sequence lengths as uint16 when every sequence is shorter than 65 536 bases (pinned, like everything handed over)
    len16_ptr = None
    if n_seqs and int(batch.seq_len.max()) < 65536:
        len16_ptr = L.tdb_host_alloc(2 * n_seqs)
        np.frombuffer((C.c_uint16 * n_seqs).from_address(len16_ptr), dtype=np.uint16)[:] = batch.seq_len
    pseqs = B.PackedSeqs(h_packed_ptr, batch.blob.size, h_off_ptr, None if len16_ptr else batch.seq_len.ctypes.data, n_seqs,
                         exc.ctypes.data if len(exc) else None, exc_bytes.ctypes.data if len(exc) else None, len(exc),
                         len16_ptr)

    def step_packed():
        r = L.tdb_trio_batch_packed if not args.duo else L.tdb_duo_batch_packed
        rc = r(h, C.byref(pseqs), batch.pair_pattern.ctypes.data, batch.pair_text.ctypes.data, n_pairs, batch.loci_ptr,
               n_loci, args.quantile, C.cast(h_out, C.c_void_p), None, h_mask.ctypes.data)
        if rc != 0:
            raise RuntimeError(L.tdb_last_error(h).decode())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------------
    # one untimed counting step: exact algorithmic C/E/T of the workload (every pair as the CPU oracle counts it)
    ctx.set_count_work(True)
    algo = step_device()
    algo_work = {"cells": int(algo.cells), "ext16": int(algo.ext16), "columns": int(algo.columns)}
    ctx.set_count_work(False)
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    w0 = time.time()
    dev_ms, launches = 0.0, 0
    stage = {k: 0.0 for k in ("pack", "dedup", "align_thread", "align_quad", "align_warp", "align_global", "scatter",
                              "reduce")}
    cnt = None
    for _ in range(args.steps):
        cnt = step_device()
        dev_ms += cnt.ms_total_device
        stage["pack"] += cnt.ms_pack
        stage["dedup"] += cnt.ms_dedup
        stage["align_thread"] += cnt.ms_thread_tier
        stage["align_quad"] += cnt.ms_tier[0]
        stage["align_warp"] += cnt.ms_tier[1]
        stage["align_global"] += cnt.ms_tier[2]
        stage["scatter"] += cnt.ms_scatter
        stage["reduce"] += cnt.ms_reduce
        launches += cnt.kernel_launches
    barrier()
    w1 = time.time()
    clocks = sampler.stop(w0, w1)
    wall_ms = 1e3 * (w1 - w0)
    score_ref = d_score.clone()  # production result of this rank's shard (self-check and census below)

    # ---- end-to-end through the host-buffer C ABI: ASCII hand-over, then packed hand-over --------------
    def time_host(fn):
        for _ in range(min(args.warmup, 3)):
            fn()
        barrier()
        e0 = time.time()
        for _ in range(args.steps):
            fn()
        barrier()
        return time.time() - e0, ctx.counters()

    e2e_s, ce = time_host(step_host)
    out_ascii = bytes(h_out)
    e2e_host = {"host_packed": int(ce.host_packed), "align_bytes_h2d": int(ce.bytes_h2d),
                "ascii_bytes_taken_over": int(ce.bytes_raw), "ms_plan": ce.ms_host_plan, "ms_pack_wait": ce.ms_host_pack_wait,
                "ms_align_call": ce.ms_host_call, "ms_device_span": ce.ms_total_device,
                "ms_last_block_packed": ce.ms_host_packed, "ms_last_chunk_queued": ce.ms_host_enqueued,
                "chunks": int(ce.chunks)}
    # bytes that actually cross PCIe per step: what the library copied for the alignment stage (2-bit packed
    # sequences, tables, run-length encoded pair lists) plus the locus descriptors
    host_input_bytes = batch.blob.size + (0 if dense else batch.seq_off.nbytes) + batch.seq_len.nbytes + \
        batch.pair_pattern.nbytes + batch.pair_text.nbytes + loci_np.nbytes
    h2d = int(ce.bytes_h2d) + loci_np.nbytes
    d2h = int(ce.bytes_d2h)  # rows + the read mask (large calls fetch it as a list of its set positions)

    e2ep_s, cp = time_host(step_packed)
    if bytes(h_out) != out_ascii:
        raise SystemExit("bench self-check failed: packed and ASCII entry points disagree")
    del out_ascii
    packed_input_bytes = n_words * 4 + (0 if dense else batch.seq_off.nbytes) + (2 * n_seqs if len16_ptr else batch.seq_len.nbytes) + \
        batch.pair_pattern.nbytes + batch.pair_text.nbytes + loci_np.nbytes
    e2ep_host = {"host_packed": int(cp.host_packed), "align_bytes_h2d": int(cp.bytes_h2d), "ms_plan": cp.ms_host_plan,
                 "ms_scan_wait": cp.ms_host_pack_wait, "ms_align_call": cp.ms_host_call,
                 "ms_device_span": cp.ms_total_device, "ms_last_chunk_queued": cp.ms_host_enqueued,
                 "ms_fused_call": cp.ms_fused_call, "ms_fused_tail": cp.ms_fused_tail,
                 "chunks": int(cp.chunks), "load_time_pack_s": load_pack_s, "exceptions": int(len(exc))}
    h2d_p = int(cp.bytes_h2d) + loci_np.nbytes
    d2h_p = int(cp.bytes_d2h)

    # ---- self-check of the step's result against the oracle (outside the timed regions): a strided sample
    # of 20 000 pairs over this rank's whole shard ---------------------------------------------------------
    check = None
    if rank == 0:
        from oracle import tdo
        idx = np.arange(0, n_pairs, max(1, n_pairs // 20000), dtype=np.int64)
        want, wst, _ = tdo.align_batch(batch.blob, batch.seq_off, batch.seq_len, batch.pair_pattern[idx],
                                       batch.pair_text[idx], args.clip, n_threads=os.cpu_count() or 1)
        got = score_ref[torch.from_numpy(idx).to(dev)].cpu().numpy()
        check = bool((got == want).all() and (wst == 0).all())
        if not check:
            raise SystemExit("bench self-check failed: GPU scores differ from the oracle")

    # ---- heuristic census (rank 0's shard): how many pairs does wf-adaptive touch at all?  For all the others an
    # independent Gotoh DP pins the penalty and no recollection of WFA2-lib's cut-off is involved --------------
    census = None
    if rank == 0 and not args.no_census:
        def scores_with(heur, clip):
            ctx.set_heuristic(heur)
            ctx.set_clip_len(clip)
            step_device()
            return d_score.clone()
        try:
            on50 = score_ref
            off50 = scores_with((0, 0, 0, 0), args.clip)
            off0 = scores_with((0, 0, 0, 0), 0)
            on0 = scores_with((1, 10, 50, 1), 0)
            sens_clip = int((on50 != off50).sum().item())
            sens_pen = int((on0 != off0).sum().item())
            sens_any = int(((on50 != off50) | (on0 != off0)).sum().item())
            census = {"pairs": int(n_pairs), "heuristic_sensitive_pairs": sens_any,
                      "clipped_score_differs": sens_clip, "penalty_differs": sens_pen,
                      "note": "pairs of rank 0's shard whose clipped score (clip_len) or penalty (= clipped score at "
                              "clip 0) differs between wf-adaptive(10,50,1) and Heuristic::None"}
            del off50, off0, on0
        finally:
            ctx.set_heuristic((1, 10, 50, 1))
            ctx.set_clip_len(args.clip)

    # ---- reduce over ranks -------------------------------------------------------------------------------
    mine = torch.tensor([dev_ms, wall_ms, 1e3 * e2e_s, 1e3 * e2ep_s, float(n_loci), float(n_pairs),
                         float(ce.ms_host_pack_wait), float(ce.ms_host_call), float(cp.ms_fused_call), float(h2d),
                         float(d2h), float(h2d_p), float(d2h_p), float(launches), float(host_input_bytes),
                         float(packed_input_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        allv = torch.stack(allv).cpu().numpy()
    else:
        allv = mine.cpu().numpy()[None, :]
    dev_ms_max, wall_ms_max, e2e_ms_max, e2ep_ms_max = [float(allv[:, i].max()) for i in range(4)]
    tot_loci, tot_pairs = float(allv[:, 4].sum()), float(allv[:, 5].sum())
    per_rank = [{"rank": r, "loci": int(allv[r, 4]), "pairs": int(allv[r, 5]),
                 "device_ms_per_step": allv[r, 0] / args.steps, "e2e_ms_per_step": allv[r, 2] / args.steps,
                 "e2e_packed_ms_per_step": allv[r, 3] / args.steps, "last_call_ms_pack_wait": allv[r, 6],
                 "last_call_ms_e2e_align": allv[r, 7], "last_call_ms_e2e_packed_whole": allv[r, 8]} for r in range(world)]

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import tdo
        allthreads = os.cpu_count() or 1
        sample = min(n_loci, args.cpu_sample_loci)
        oracle_pass(tdo, batch, min(sample, 2000), args.clip, args.quantile, args.duo, allthreads)
        dt, np_s = oracle_pass(tdo, batch, sample, args.clip, args.quantile, args.duo, allthreads)
        dt1, np_1 = oracle_pass(tdo, batch, min(sample, 20000), args.clip, args.quantile, args.duo, 1)
        cpu_baseline = {"value": sample / dt, "unit": "loci/s", "cores": allthreads, "kind": "port",
                        "sample": f"first {sample} loci of the workload ({np_s} pairs), {allthreads} threads, "
                                  f"{dt:.1f}s; scalar C restatement of the reference path, not WFA2-lib",
                        "alignments_per_sec": np_s / dt,
                        "single_thread_loci_per_sec": min(sample, 20000) / dt1}

    if rank == 0:
        # roofline of the dominant kernel, per launch, rank 0's shard.  Work = what THAT launch executed
        # (representative pairs only; duplicates are copied, not aligned), SURVEY.md 8(d) accounting.
        tiers = ("align_quad", "align_warp", "align_global", "align_thread")
        names = {"align_quad": "k_align_lock<QuadCfg>", "align_global": "k_align_lock<LongCfg> + k_align_global", "align_thread": "k_align_thread",
                 "align_warp": "k_align_warp" if os.environ.get("TDB_DISABLE_DUO") else
                               ("k_align_lock<DuoCfg>" if os.environ.get("TDB_DUO_G16") else "k_align_lock<Duo4Cfg>")}
        ti = max(range(4), key=lambda i: stage[tiers[i]])
        k_ms = stage[tiers[ti]] / args.steps
        k_s = max(k_ms, 1e-6) * 1e-3
        # ("align_global" = everything behind tier 1: the long tier and the global-scratch tiers)
        t_pairs = [int(x) for x in cnt.tier_pairs[:2]] + [int(cnt.tier_pairs[2] + cnt.tier_pairs[3] + cnt.long_pairs), int(cnt.thread_pairs)]
        t_cells = [int(x) for x in cnt.tier_cells[:2]] + [int(cnt.tier_cells[2] + cnt.tier_cells[3] + cnt.long_cells), int(cnt.thread_cells)]
        t_ext = [int(x) for x in cnt.tier_ext16[:2]] + [int(cnt.tier_ext16[2] + cnt.tier_ext16[3] + cnt.long_ext16), int(cnt.thread_ext16)]
        t_cols = [int(x) for x in cnt.tier_columns[:2]] + [int(cnt.tier_columns[2] + cnt.tier_columns[3] + cnt.long_columns), int(cnt.thread_columns)]
        t_ops = [INT_OPS_PER_CELL * t_cells[i] + INT_OPS_PER_EXT * t_ext[i] + INT_OPS_PER_COL * t_cols[i] for i in range(4)]
        ops = t_ops[ti]
        clk = (clocks["sm_mhz"] or sm_max_mhz) * 1e6
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        int_peak = n_sm * 128 * clk
        # algorithmic HBM bytes of the launch: 2-bit sequences + 16 B descriptor/result per pair it aligned
        mean_pair_bytes = float(np.mean((batch.seq_len[batch.pair_pattern[:200000]].astype(np.int64) + 3) // 4 +
                                        (batch.seq_len[batch.pair_text[:200000]].astype(np.int64) + 3) // 4)) + 16.0
        algo_bytes = mean_pair_bytes * float(t_pairs[ti])
        step_s = dev_ms / args.steps * 1e-3
        align_ms = sum(stage[t] for t in tiers) / args.steps
        # DRAM traffic of that kernel per launch: bytes per aligned pair from the committed ncu --set full capture
        # (profiles/r02/traffic.json, else r01) x the pairs this launch aligned
        traffic, traffic_src = None, None
        for rd in ("r02", "r01"):
            tj = os.path.join(ROOT, "profiles", rd, "traffic.json")
            if os.path.exists(tj):
                tinfo = json.load(open(tj)).get(names[tiers[ti]])
                if tinfo:
                    traffic = tinfo["dram_bytes_per_pair"] * float(t_pairs[ti])
                    traffic_src = tinfo["source"] + ", per-pair bytes scaled to this launch"
                    break
        roofline = {
            "bound": "int32_issue", "kernel": names[tiers[ti]], "achieved": ops / k_s, "peak": int_peak,
            "unit": "lane-int-ops/s", "frac": ops / k_s / int_peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": f"{n_sm} SM x 128 lanes x {clk / 1e6:.0f} MHz (clock observed under load)",
            "accounting": "SURVEY.md 8(d): 24*C + 8*E + 4*T of the pairs this launch aligned (counted exactly on the "
                          "device, equal to the oracle's count in tests/), over its CUDA-event duration",
            "pairs_aligned_by_kernel": t_pairs[ti], "kernel_ms": k_ms,
            "step_share": stage[tiers[ti]] / max(dev_ms, 1e-9),
            "all_alignment_tiers": {"ops_per_s": sum(t_ops) / max(align_ms * 1e-3, 1e-9),
                                    "frac": sum(t_ops) / max(align_ms * 1e-3, 1e-9) / int_peak, "ms": align_ms,
                                    "pairs": sum(t_pairs)},
            "hbm": {"bound": "hbm", "achieved": algo_bytes / k_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": algo_bytes / k_s / 1e9 / hbm_peak, "peak_source": f"{peak_src} MEASURED_PEAKS.json hbm_gbs",
                    "note": "not the binding roofline for this integer DP (SURVEY.md 8d)"},
        }
        def e2e_obj(ms_max, h2d_col, d2h_col, in_col):
            return {"value": tot_loci * args.steps / (ms_max * 1e-3), "unit": "loci/s",
                    "h2d_bytes_per_step": int(allv[:, h2d_col].sum()), "d2h_bytes_per_step": int(allv[:, d2h_col].sum()),
                    "host_input_bytes_per_step": int(allv[:, in_col].sum()), "ms_per_step": ms_max / args.steps,
                    "alignments_per_sec": tot_pairs * args.steps / (ms_max * 1e-3)}
        e2e = e2e_obj(e2e_ms_max, 9, 10, 14)
        e2e["input"] = "ASCII bytes in pinned host memory (what the reference hands over); the library 2-bit packs them on the caller's CPU threads"
        e2e_packed = e2e_obj(e2ep_ms_max, 11, 12, 15)
        e2e_packed["input"] = "2-bit stream packed at load time (tdb_*_batch_packed); no per-base CPU work in the call"
        line = {
            "metric": "trio_loci_per_sec" if not args.duo else "duo_loci_per_sec",
            "value": tot_loci * args.steps / (dev_ms_max * 1e-3), "unit": "loci/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "catalog_loci": int(tot_loci), "catalog_pairs": int(tot_pairs),
                       "clip_len": args.clip, "parent_quantile": args.quantile,
                       "sharding": f"one catalog, {world} contiguous locus ranges balanced on sum(n+m), no collective",
                       "shard_bounds": [int(x) for x in bounds],
                       "layout": "dense (seq_off = NULL)" if dense else "explicit seq_off",
                       "l2": "inputs larger than L2 (no flush needed)"},
            "alignments_per_sec": tot_pairs * args.steps / (dev_ms_max * 1e-3),
            "gcups": batch.sum_nm * args.steps / (dev_ms * 1e-3) / 1e9,
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": e2e, "e2e_packed": e2e_packed,
            "gpu_launches": int(allv[:, 13].sum()),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "per_rank": per_rank,
            "breakdown_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            "tier_pairs": [int(x) for x in cnt.tier_pairs], "thread_tier_pairs": int(cnt.thread_pairs),
            "long_tier_pairs": int(cnt.long_pairs),
            "pairs_identical": int(cnt.pairs_identical), "pairs_duplicate": int(cnt.pairs_duplicate),
            "pairs_closed_form": int(cnt.pairs_closed_form), "pairs_failed": int(cnt.pairs_failed),
            "work": {"cells": algo_work["cells"], "ext16": algo_work["ext16"], "columns": algo_work["columns"],
                     "exec_cells": int(cnt.exec_cells), "exec_ext16": int(cnt.exec_ext16),
                     "exec_columns": int(cnt.exec_columns), "note": "cells/ext16/columns: every pair of rank 0's shard "
                     "as the CPU oracle counts it; exec_*: what the kernels executed"},
            "heuristic_census": census,
            "e2e_host": e2e_host, "e2e_packed_host": e2ep_host,
            "self_check_vs_oracle": check,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
