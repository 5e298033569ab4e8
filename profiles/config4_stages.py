"""Stage times of one tdb_trio_batch call on the long-expansion stress set (config 4, 1000 loci): where the call goes."""
import os, sys, ctypes as C, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
entry.build()
from importlib import import_module
import trgt_denovo_b200 as tdb
S = import_module("trgt_denovo_b200.synth"); B = import_module("trgt_denovo_b200.binding")
b = S.generate(S.config("long_expansion", pinned=1), 0, 1000, os.cpu_count())
ctx = tdb.Context(device_id=0)
L = tdb.lib()
out = (B.AlleleOut * (2 * b.n_loci))(); mask = np.zeros(b.n_pairs, np.uint8)
for rep in range(2):
    t0 = time.perf_counter()
    r = L.tdb_trio_batch(ctx._h, b.blob.ctypes.data, b.blob.size, None, b.seq_len.ctypes.data, b.n_seqs, b.pair_pattern.ctypes.data,
                         b.pair_text.ctypes.data, b.n_pairs, b.loci_ptr, b.n_loci, 1.0, C.cast(out, C.c_void_p), None, mask.ctypes.data)
    dt = time.perf_counter() - t0
    c = ctx.counters()
    print("call ms", round(dt*1e3,1), "device", round(c.ms_total_device,1), "pack", round(c.ms_pack,2), "dedup", round(c.ms_dedup,2), "thread", round(c.ms_thread_tier,2),
          "quad", round(c.ms_tier[0],2), "tier1", round(c.ms_tier[1],2), "global", round(c.ms_tier[2],2), "long", round(c.ms_long_tier,2), c.long_pairs, "pairs", c.thread_pairs, list(c.tier_pairs),
          "cells", list(c.tier_cells), "ext16", list(c.tier_ext16), "cols", list(c.tier_columns), "chunks", c.chunks)
ln = b.seq_len
print("len pct", np.percentile(ln, [50, 90, 99, 100]))
