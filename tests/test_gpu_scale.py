"""GPU, at scale: properties that do not need the CPU oracle to finish (it checks a sample) on a 100k-locus
slice of the genome-wide synthetic trio (config 2), 15 M pairs:
  * production mode (closed forms, duplicate elimination, chunked + host-packed upload) and counting mode
    (every distinct content through the wavefront kernels) give the same scores      -> checksum equality
  * device-resident call == host-buffer call == two half-batches concatenated        -> sharding / chunking
  * duplicates and identical pairs: every pair with pattern bytes == text bytes scores 0; scores <= 0
  * the per-allele outputs of the fused call agree between the paths
  * a strided sample of 200k pairs is bit-exact against the oracle
"""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_LOCI = int(os.environ.get("TDB_SCALE_TEST_LOCI", 100000))


@pytest.fixture(scope="module")
def env():
    import __graft_entry__ as entry
    entry.build()
    from importlib import import_module
    import trgt_denovo_b200 as tdb
    assert tdb.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    S = import_module("trgt_denovo_b200.synth")
    batch = S.generate(S.config("trio_genome", pinned=1), 0, N_LOCI, os.cpu_count() or 1)
    ctx = tdb.Context(device_id=0)
    yield tdb, S, batch, ctx
    ctx.close()
    batch.close()


def _assess(tdb, ctx, b, lo=0, hi=None):
    """tdb_trio_batch on loci [lo, hi) of batch b (pair / sequence indices rebased) -> scores, outs, mask"""
    hi = b.n_loci if hi is None else hi
    B = tdb.binding if hasattr(tdb, "binding") else None
    from importlib import import_module
    B = import_module("trgt_denovo_b200.binding")
    s0, s1 = int(b.meta[lo].seq_begin), int(b.meta[hi - 1].seq_end)
    p0, p1 = int(b.meta[lo].pair_begin), int(b.meta[hi - 1].pair_end)
    off = (b.seq_off[s0:s1] - b.seq_off[s0]).astype(np.uint64)
    ln = b.seq_len[s0:s1]
    blob = b.blob[int(b.seq_off[s0]):int(b.seq_off[s1 - 1]) + int(b.seq_len[s1 - 1])]
    pp = (b.pair_pattern[p0:p1] - s0).astype(np.uint32)
    pt = (b.pair_text[p0:p1] - s0).astype(np.uint32)
    loci = (B.TrioLocus * (hi - lo))()
    C.memmove(loci, C.byref(b.loci[lo]), C.sizeof(B.TrioLocus) * (hi - lo))
    arr = np.frombuffer(loci, dtype=np.uint32).reshape(hi - lo, C.sizeof(B.TrioLocus) // 4)
    arr[:, 3:15] -= p0  # list_off[2][6]
    out, scores, mask = ctx.assess_batch(blob, off, ln, pp, pt, loci)
    return scores.copy(), np.frombuffer(out, dtype=np.uint8).reshape(-1, C.sizeof(B.AlleleOut)).copy(), mask.copy()


def test_paths_agree_at_scale(env):
    tdb, S, b, ctx = env
    from oracle import tdo
    ctx.set_count_work(False)
    s_fast, o_fast, m_fast = _assess(tdb, ctx, b)
    cnt = ctx.counters()
    assert cnt.chunks > 1 and cnt.host_packed == 1 and cnt.pairs_closed_form > 0 and cnt.pairs_duplicate > 0
    ctx.set_count_work(True)  # every distinct content through the wavefront kernels
    s_full, o_full, m_full = _assess(tdb, ctx, b)
    assert ctx.counters().pairs_closed_form == 0
    ctx.set_count_work(False)
    assert zlib.crc32(s_fast.tobytes()) == zlib.crc32(s_full.tobytes())
    assert (o_fast == o_full).all() and (m_fast == m_full).all()
    assert (s_fast <= 0).all()
    # identical bytes => score 0 (checked on the pairs whose lengths agree, by comparing the bytes)
    same_len = np.nonzero(b.seq_len[b.pair_pattern] == b.seq_len[b.pair_text])[0][:200000]
    for j in same_len[::97]:
        if b.seq(int(b.pair_pattern[j])) == b.seq(int(b.pair_text[j])):
            assert s_fast[j] == 0
    # two half-batches (what two GPUs would each see) concatenate to the whole
    half = b.n_loci // 2
    sa, oa, ma = _assess(tdb, ctx, b, 0, half)
    sb, ob, mb = _assess(tdb, ctx, b, half, b.n_loci)
    assert (np.concatenate([sa, sb]) == s_fast).all()
    keep = np.array([c < b.loci[i].n_child for i in range(b.n_loci) for c in range(2)])
    assert (np.concatenate([oa, ob])[keep] == o_fast[keep]).all()
    # the whole batch fetched its mask as a list of set positions (>= 8 Mi pairs), the halves byte per pair
    assert (np.concatenate([ma, mb]) == m_fast).all()
    # strided oracle sample
    idx = np.arange(0, b.n_pairs, max(1, b.n_pairs // 200000), dtype=np.int64)
    want, st, _ = tdo.align_batch(b.blob, b.seq_off, b.seq_len, b.pair_pattern[idx], b.pair_text[idx], 50,
                                  n_threads=os.cpu_count() or 1)
    assert (st == 0).all() and (want == s_fast[idx]).all()
