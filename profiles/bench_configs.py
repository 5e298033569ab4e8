#!/usr/bin/env python3
"""BASELINE.json configs 1, 3, 4, 5 (parity-test / stress cases, not the bench line): throughput on one B200
next to the CPU oracle on a bounded sample, with a bit-exact score check on that sample.

  config 1  1 000-locus trio (seed 20250001)            loci/s, alignments/s
  config 3  duo on the genome-wide catalog              loci/s (bench.py --duo does the full contract line)
  config 4  long-expansion stress, 60x, alleles to 10 kb loci/s, alignments/s, tier split
  config 5  microbenchmark: length x divergence sweep   GCUPS = sum(n*m)/t/1e9, heuristic on and off
"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
entry.build()
from importlib import import_module
import trgt_denovo_b200 as tdb
from oracle import tdo
S = import_module("trgt_denovo_b200.synth")
B = import_module("trgt_denovo_b200.binding")
ap = argparse.ArgumentParser()
ap.add_argument("--long-loci", type=int, default=1000)
ap.add_argument("--duo-loci", type=int, default=300000)
ap.add_argument("--micro-cells", type=float, default=2e10, help="sum(n*m) per microbenchmark cell (SURVEY: 1e8.. per cell)")
ap.add_argument("--skip", default="")
ap.add_argument("--census-long-loci", type=int, default=100, help="loci of config 4 aligned with and without the heuristic")
args = ap.parse_args()
threads = os.cpu_count() or 1
ctx = tdb.Context(device_id=0)


def run_loci(name, batch, duo, cpu_loci, reps=3, census_loci=0):
    L = tdb.lib()
    out = (B.AlleleOut * (2 * batch.n_loci))()
    mask = np.zeros(batch.n_pairs, np.uint8)
    scores = np.zeros(batch.n_pairs, np.int32)
    fn = L.tdb_duo_batch if duo else L.tdb_trio_batch
    def call():
        r = fn(ctx._h, batch.blob.ctypes.data, batch.blob.size, None, batch.seq_len.ctypes.data,  # dense layout
               batch.n_seqs, batch.pair_pattern.ctypes.data, batch.pair_text.ctypes.data, batch.n_pairs, batch.loci_ptr,
               batch.n_loci, 1.0, C.cast(out, C.c_void_p), scores.ctypes.data, mask.ctypes.data)
        assert r == 0, L.tdb_last_error(ctx._h)
    call()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    dt = (time.perf_counter() - t0) / reps
    cnt = ctx.counters()
    ncpu = min(cpu_loci, batch.n_loci)
    npc = int(batch.meta[ncpu - 1].pair_end)
    t1 = time.perf_counter()
    want, wst, _ = tdo.align_batch(batch.blob, batch.seq_off, batch.seq_len, batch.pair_pattern[:npc], batch.pair_text[:npc], 50, n_threads=threads)
    cpu_dt = time.perf_counter() - t1
    ok = bool((want == scores[:npc]).all())
    census = None
    if census_loci:
        # heuristic census on the GPU: pairs of the first `census_loci` loci whose clipped score (clip 50) or penalty (clip 0)
        # differs between wf-adaptive(10,50,1) and Heuristic::None -- only for those does the result rest on the restated
        # cut-off; for the others the independent Gotoh DP of the checker pins the penalty
        nl = min(census_loci, batch.n_loci)
        nq = int(batch.meta[nl - 1].pair_end)
        pp, pt = batch.pair_pattern[:nq], batch.pair_text[:nq]
        res = {}
        for hname, heur in (("on", (1, 10, 50, 1)), ("off", (0, 0, 0, 0))):
            c2 = tdb.Context(device_id=0, heuristic=heur)
            for clip in (50, 0):
                c2.set_clip_len(clip)
                sc, st = c2.align_batch(batch.blob, None, batch.seq_len, pp, pt)
                assert (st == 0).all()
                res[hname, clip] = sc.copy()
            c2.close()
        d50, d0 = res["on", 50] != res["off", 50], res["on", 0] != res["off", 0]
        census = {"loci": nl, "pairs": nq, "heuristic_sensitive_pairs": int((d50 | d0).sum()),
                  "clipped_score_differs": int(d50.sum()), "penalty_differs": int(d0.sum())}
    print(json.dumps({"config": name, "loci": batch.n_loci, "pairs": batch.n_pairs, "heuristic_census": census, "e2e_loci_per_s": batch.n_loci / dt,
                      "e2e_alignments_per_s": batch.n_pairs / dt, "e2e_ms": dt * 1e3, "gcups": batch.sum_nm / dt / 1e9,
                      "thread_tier_pairs": int(cnt.thread_pairs), "long_tier_pairs": int(cnt.long_pairs), "tier_pairs": [int(x) for x in cnt.tier_pairs], "identical": int(cnt.pairs_identical),
                      "duplicate": int(cnt.pairs_duplicate), "closed_form": int(cnt.pairs_closed_form),
                      "cpu_port": {"threads": threads, "loci": ncpu, "pairs": npc, "loci_per_s": ncpu / cpu_dt,
                                   "alignments_per_s": npc / cpu_dt, "align_only": True},
                      "scores_bit_exact_on_cpu_sample": ok}), flush=True)
    assert ok


if "1" not in args.skip:
    run_loci("1: 1000-locus trio", S.generate(S.config("trio_1k", pinned=1), 0, 1000, threads), False, 1000, reps=10, census_loci=1000)
if "3" not in args.skip:
    run_loci("3: duo genome-wide (%d loci)" % args.duo_loci, S.generate(S.config("duo_genome", pinned=1), 0, args.duo_loci, threads), True, 50000, census_loci=50000)
if "4" not in args.skip:
    run_loci("4: long-expansion stress 60x (%d loci)" % args.long_loci,
             S.generate(S.config("long_expansion", pinned=1), 0, args.long_loci, threads), False, 40, reps=1, census_loci=args.census_long_loci)
if "5" not in args.skip:
    for heur, hname in (((1, 10, 50, 1), "wf-adaptive(10,50,1) = reference config"), ((0, 0, 0, 0), "none = optimal")):
        c2 = tdb.Context(device_id=0, heuristic=heur)
        for length in (100, 200, 500, 1000, 2000, 5000, 10000):
            for div in (0.0, 0.01, 0.02, 0.05, 0.10, 0.20):
                n_pairs = int(max(1000, min(2_000_000, args.micro_cells / (length * length))))
                b = S.micro(20250005 + length, length, div, n_pairs, threads, pinned=True)
                c2.align_batch(b.blob, None, b.seq_len, b.pair_pattern, b.pair_text)
                t0 = time.perf_counter()
                got, st = c2.align_batch(b.blob, None, b.seq_len, b.pair_pattern, b.pair_text)
                dt = time.perf_counter() - t0
                cnt = c2.counters()
                ncpu = int(max(16, min(n_pairs, 2e9 / (length * length) / (1 + 30 * div))))
                t1 = time.perf_counter()
                want, _, _ = tdo.align_batch(b.blob, b.seq_off, b.seq_len, b.pair_pattern[:ncpu], b.pair_text[:ncpu], 50,
                                             heuristic=heur, n_threads=threads)
                cdt = time.perf_counter() - t1
                ok = bool((want == got[:ncpu]).all())
                nm = float(b.sum_nm)
                print(json.dumps({"config": "5: microbenchmark", "heuristic": hname, "length": length, "divergence": div,
                                  "pairs": n_pairs, "e2e_gcups": nm / dt / 1e9, "e2e_alignments_per_s": n_pairs / dt,
                                  "thread_tier_pairs": int(cnt.thread_pairs), "long_tier_pairs": int(cnt.long_pairs), "tier_pairs": [int(x) for x in cnt.tier_pairs],
                                  "cpu_port_gcups": (nm * ncpu / n_pairs) / cdt / 1e9, "cpu_threads": threads,
                                  "scores_bit_exact_on_cpu_sample": ok}), flush=True)
                assert ok
                b.close()
        c2.close()
