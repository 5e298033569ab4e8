#!/usr/bin/env python3
"""Randomised parity stress aimed at the thread-per-pair tier: millions of light pairs (a target -- random or a
short tandem repeat -- against a copy with 0..5 small edits), every length from 16 to 340, clip 50 and 0, heuristic on
and off: GPU clipped scores and exact C/E/T work totals against the CPU oracle."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
entry.build()
import trgt_denovo_b200 as tdb
from oracle import tdo

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(20250019)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make(n_pairs):
    chunks, lens, total = [], [], 0
    for i in range(n_pairs):
        L = int(rng.integers(16, 341))
        if rng.random() < 0.6:  # tandem repeat with flanks
            k = int(rng.integers(1, 7)); motif = ACGT[rng.integers(0, 4, k)]
            fl = int(rng.integers(0, min(60, L // 2) + 1))
            body = np.tile(motif, (L - 2 * fl) // k + 1)[: L - 2 * fl]
            t = np.concatenate([ACGT[rng.integers(0, 4, fl)], body, ACGT[rng.integers(0, 4, fl)]])
        else:
            t = ACGT[rng.integers(0, 4, L)]
        p = t.copy()
        for _ in range(int(rng.integers(0, 6))):
            kind = rng.random(); pos = int(rng.integers(0, len(p) + 1))
            if kind < 0.4 and pos < len(p):
                p[pos] = ACGT[rng.integers(0, 4)]
            elif kind < 0.7:
                g = int(rng.integers(1, 5)); p = np.concatenate([p[:pos], ACGT[rng.integers(0, 4, g)], p[pos:]])
            else:
                g = int(rng.integers(1, 5)); p = np.concatenate([p[:pos], p[pos + g:]])
        if len(p) == 0:
            p = ACGT[:1].copy()
        chunks += [t, p]; lens += [len(t), len(p)]
    blob = np.concatenate(chunks)
    ln = np.array(lens, np.uint32)
    off = np.concatenate([[0], np.cumsum(ln, dtype=np.uint64)[:-1]]).astype(np.uint64)
    idx = np.arange(n_pairs, dtype=np.uint32)
    return blob, off, ln, 2 * idx + 1, 2 * idx


t0 = time.time()
blob, off, ln, pp, pt = make(n_pairs)
print(f"{n_pairs} pairs, {blob.size/1e6:.0f} MB, generated in {time.time()-t0:.0f}s", flush=True)
for heur, hname in ((tdo.DEFAULT_HEURISTIC, "wf-adaptive"), (tdo.NO_HEURISTIC, "none")):
    ctx = tdb.Context(device_id=0, heuristic=heur)
    for clip in (50, 0):
        ctx.set_clip_len(clip)
        want, wst, wtot = tdo.align_batch(blob, off, ln, pp, pt, clip, heuristic=heur, n_threads=os.cpu_count())
        for count in (False, True):
            ctx.set_count_work(count)
            got, st = ctx.align_batch(blob, None, ln, pp, pt)
            c = ctx.counters()
            bad = int((got != want).sum()) + int((st != wst).sum())
            work_ok = (not count) or (c.cells, c.ext16, c.columns) == (wtot.cells, wtot.ext16, wtot.columns)
            print(f"heuristic {hname} clip {clip} count_work {count}: mismatches {bad}, work totals equal {work_ok}, "
                  f"thread tier {c.thread_pairs} quad {c.tier_pairs[0]} tier1 {c.tier_pairs[1]} global {c.tier_pairs[2] + c.tier_pairs[3]} "
                  f"closed form {c.pairs_closed_form} identical {c.pairs_identical} duplicate {c.pairs_duplicate}", flush=True)
            assert bad == 0 and work_ok
    ctx.close()
print("stress ok")
