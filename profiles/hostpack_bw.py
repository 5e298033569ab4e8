#!/usr/bin/env python3
"""Host packer throughput vs thread count, and plain memory bandwidth of the box (numpy copy), to tell
whether the e2e path is bound by the packer's code or by host DRAM."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e
e.build()
from importlib import import_module
import trgt_denovo_b200 as tdb
S = import_module("trgt_denovo_b200.synth")
L = tdb.lib()
for pinned in (1, 0):
    b = S.generate(S.config("trio_genome", pinned=pinned), 0, 100000, 16)
    n = L.tdb_packed_words(b.blob.size)
    if pinned:
        pp = L.tdb_host_alloc(n * 4); packed = np.frombuffer((C.c_uint8 * (n * 4)).from_address(pp), dtype=np.uint32)
    else:
        packed = np.zeros(n, np.uint32)
    flag = np.zeros(b.n_seqs, np.uint8)
    for nt in (1, 2, 4, 8, 16, 32):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            r = L.tdb_host_pack(b.blob.ctypes.data, b.blob.size, b.seq_off.ctypes.data, b.seq_len.ctypes.data, b.n_seqs,
                                packed.ctypes.data, flag.ctypes.data, nt)
            best = min(best, time.perf_counter() - t0)
        print(f"pinned={pinned} isa={L.tdb_host_pack_isa().decode()} threads={nt:2d} {best*1e3:7.2f} ms  {b.blob.size/best/1e9:6.1f} GB/s in", flush=True)
    a = np.frombuffer(b.blob, dtype=np.uint8); d = np.empty_like(a)
    t0 = time.perf_counter(); d[:] = a; dt = time.perf_counter() - t0
    print(f"pinned={pinned} single-thread numpy copy {a.size/dt/1e9:.1f} GB/s")
    b.close()
