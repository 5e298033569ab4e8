"""GPU: parity of the CUDA path (through the C ABI) with the CPU oracle and the reference's vectors.

Bar: bit-exact for every integer output (clipped score, penalty, status, CIGAR bytes, counts);
floating-point reduction outputs within 1e-9 relative (tolerance stated by north_star).
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import tdo
from tests.util import make_pairs, mutate, rand_seq, to_arrays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tdb():
    import __graft_entry__ as entry
    entry.build()
    import trgt_denovo_b200 as m
    assert m.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return m


@pytest.fixture(scope="module")
def ctx(tdb):
    c = tdb.Context(device_id=0)
    yield c
    c.close()


def test_reference_known_answers(tdb, golden_dir):
    """src/aligner.rs tests B5,B6,B7,B9 through the WFAligner shim (B4 is single-piece affine: oracle only)."""
    for x in json.load(open(os.path.join(golden_dir, "aligner_vectors.json"))):
        if x["penalties"][3] < 0:
            continue
        heur = (1, 10, 50, 1) if x["heuristic"] == "default" else (0, 0, 0, 0)
        c = tdb.Context(penalties=tuple(x["penalties"]), heuristic=heur)
        a = tdb.WFAligner(c)
        st = a.align_end_to_end(x["pattern"].encode(), x["text"].encode())
        assert st == tdb.AlignmentStatus.StatusAlgCompleted, x["id"]
        assert a.score() == x["score"], x["id"]
        if "cigar" in x:
            assert a.cigar() == x["cigar"], x["id"]
        for clip, exp in (x.get("clipped") or {}).items():
            assert a.cigar_score_clipped(int(clip)) == exp
        for clip, exp in (x.get("clipped_cigar") or {}).items():
            assert a.cigar(int(clip)) == exp
        assert a.cigar_score() == x["score"]
        c.close()


def test_b9_default_heuristic_matches_oracle(tdb, golden_dir):
    x = [v for v in json.load(open(os.path.join(golden_dir, "aligner_vectors.json"))) if v["id"] == "B9"][0]
    a = tdb.create_aligner_with_scoring()
    assert a.align_end_to_end(x["pattern"].encode(), x["text"].encode()) == 0
    o = tdo.Aligner()
    o.align_end_to_end(x["pattern"].encode(), x["text"].encode())
    assert a.score() == o.score() == -1007
    assert a.cigar_wfa() == o.cigar_ops()
    assert a.cigar_score_clipped(50) == o.cigar_score_clipped(50) == -969


def _check_batch(ctx, seqs, pp, pt, clip, heuristic=tdo.DEFAULT_HEURISTIC, cigars=True):
    blob, off, ln = to_arrays(seqs)
    ctx.set_clip_len(clip)
    want, wst, wtot = tdo.align_batch(blob, off, ln, pp, pt, clip, heuristic=heuristic, n_threads=8)
    assert (wst == 0).all()
    # production mode: closed form for single-substitution / single-short-gap contents, no work records
    ctx.set_count_work(False)
    got, st = ctx.align_batch(blob, off, ln, pp, pt)
    fast = ctx.counters()
    assert (st == wst).all()
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, f"{len(bad)} score mismatches (closed form on), first pair {bad[:5]}: gpu {got[bad[:5]]} oracle {want[bad[:5]]}"
    assert sum(fast.tier_pairs) + fast.long_pairs + fast.thread_pairs + fast.pairs_identical + fast.pairs_duplicate + fast.pairs_closed_form == len(pp)
    # dense layout (seq_off = NULL): to_arrays lays the sequences back to back, the device derives the offsets
    got, st = ctx.align_batch(blob, None, ln, pp, pt)
    assert (st == wst).all() and (got == want).all()
    assert ctx.counters().bytes_h2d <= fast.bytes_h2d - 8 * (int(max(pp.max(), pt.max())) if len(pp) else 0)
    # packed entry point (the caller packed at load time): explicit offsets and dense layout
    import trgt_denovo_b200 as tdbm
    for dense in (False, True):
        pk = tdbm.pack_sequences(blob, off, ln, dense=dense, len16=dense and (len(ln) == 0 or int(ln.max()) < 65536))
        got, st = ctx.align_batch_packed(pk, pp, pt)
        assert (st == wst).all() and (got == want).all(), f"packed entry, dense={dense}"
        assert ctx.counters().host_packed == 2
    # counting mode: every content through the wavefront kernels
    ctx.set_count_work(True)
    got, st = ctx.align_batch(blob, off, ln, pp, pt)
    cnt = ctx.counters()
    assert (st == wst).all()
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, f"{len(bad)} score mismatches, first pair {bad[:5]}: gpu {got[bad[:5]]} oracle {want[bad[:5]]}"
    assert cnt.pairs_closed_form == 0
    # exact algorithmic work accounting agrees with the oracle (C, E, T)
    assert (cnt.cells, cnt.ext16, cnt.columns) == (wtot.cells, wtot.ext16, wtot.columns)
    if cigars:
        g2, st2, pen2, cigs = ctx.align_batch_cigar(blob, off, ln, pp, pt)
        assert (g2 == want).all() and (st2 == 0).all()
        o = tdo.Aligner(heuristic=heuristic)
        for i in range(0, len(pp), max(1, len(pp) // 300)):
            o.align_end_to_end(seqs[pp[i]], seqs[pt[i]])
            assert pen2[i] == -o.score()
            assert cigs[i] == o.cigar_ops(), f"pair {i}"
    return cnt


@pytest.mark.parametrize("clip", [50, 0])
def test_str_pairs_bit_exact(ctx, clip):
    rng = random.Random(100 + clip)
    seqs, pp, pt = make_pairs(rng, n_targets=300, reads_per_target=12)
    cnt = _check_batch(ctx, seqs, pp, pt, clip)
    aligned = sum(cnt.tier_pairs) + cnt.thread_pairs + cnt.long_pairs
    assert aligned + cnt.pairs_identical + cnt.pairs_duplicate == len(pp)
    assert cnt.thread_pairs + cnt.tier_pairs[0] + cnt.tier_pairs[1] > 0.8 * aligned  # the shared-memory tiers carry the bulk
    assert cnt.thread_pairs > 0 and cnt.tier_pairs[0] > 0 and cnt.tier_pairs[1] > 0


def test_random_divergent_pairs(ctx):
    """Unrelated / highly divergent pairs: wide wavefronts, heuristic active, global-scratch tier."""
    rng = random.Random(5)
    seqs, pp, pt = [], [], []
    for _ in range(150):
        L = rng.randint(1, 700)
        t = rand_seq(rng, L)
        p = mutate(rng, t, rng.choice([0.0, 0.02, 0.1, 0.3, 0.6]))
        pt.append(len(seqs)); seqs.append(t.encode())
        pp.append(len(seqs)); seqs.append(p.encode())
    cnt = _check_batch(ctx, seqs, np.array(pp, np.uint32), np.array(pt, np.uint32), 50)
    assert cnt.tier_pairs[2] + cnt.long_pairs > 0


def test_quad_tier_disabled_gives_same_scores(tdb, monkeypatch):
    """The one-pair-per-warp tier must agree with the four-pairs-per-warp tier on the same batch."""
    rng = random.Random(31)
    seqs, pp, pt = make_pairs(rng, n_targets=120, reads_per_target=10, max_units=25)
    monkeypatch.setenv("TDB_DISABLE_QUAD", "1")
    c = tdb.Context()
    monkeypatch.delenv("TDB_DISABLE_QUAD")
    cnt = _check_batch(c, seqs, pp, pt, 50)
    assert cnt.tier_pairs[0] == 0
    c.close()


def test_estimate_order_disabled_gives_same_scores(tdb, monkeypatch):
    """The thread-per-pair tier is ordered by an upper bound of the penalty (k_estimate) and pairs bounded beyond its last
    score skip it: only order and tier depend on the estimate, so scores and work counts are those of the |n-m| order."""
    rng = random.Random(91)
    seqs, pp, pt = make_pairs(rng, n_targets=200, reads_per_target=12, max_units=25, err=0.006)
    c = tdb.Context()
    cnt_on = _check_batch(c, seqs, pp, pt, 50, cigars=False)
    c.close()
    monkeypatch.setenv("TDB_DISABLE_EST_SORT", "1")
    c = tdb.Context()
    monkeypatch.delenv("TDB_DISABLE_EST_SORT")
    cnt_off = _check_batch(c, seqs, pp, pt, 50, cigars=False)
    c.close()
    assert cnt_on.thread_pairs > 0 and cnt_off.thread_pairs >= cnt_on.thread_pairs  # some pairs go straight to the lockstep tier
    assert (cnt_on.cells, cnt_on.ext16, cnt_on.columns) == (cnt_off.cells, cnt_off.ext16, cnt_off.columns)


def test_frame_kernels_give_same_scores(tdb, monkeypatch):
    """k_align_frame (the lockstep tiers on NULL-filled fixed rows, TDB_FRAME_LOCK=1) against the oracle: scores, status,
    exact C / E / T, with and without the thread-per-pair tier in front of it."""
    rng = random.Random(77)
    seqs, pp, pt = make_pairs(rng, n_targets=200, reads_per_target=12, max_units=40, err=0.004)
    monkeypatch.setenv("TDB_FRAME_LOCK", "1")
    c = tdb.Context()
    cnt = _check_batch(c, seqs, pp, pt, 50)
    assert cnt.tier_pairs[0] > 0 and cnt.tier_pairs[1] > 0
    c.close()
    monkeypatch.setenv("TDB_DISABLE_THREAD_TIER", "1")
    c = tdb.Context()
    monkeypatch.delenv("TDB_DISABLE_THREAD_TIER")
    monkeypatch.delenv("TDB_FRAME_LOCK")
    cnt = _check_batch(c, seqs, pp, pt, 0, cigars=False)
    assert cnt.thread_pairs == 0 and cnt.tier_pairs[0] > 0
    c.close()


def test_quad8_gives_same_scores(tdb, monkeypatch):
    """Tier 0 with eight pairs per warp (Quad8Cfg, TDB_QUAD8=1: four lanes per pair, provenance in global scratch) against
    the oracle: scores, status, exact C / E / T, with and without the thread-per-pair tier in front of it."""
    rng = random.Random(78)
    seqs, pp, pt = make_pairs(rng, n_targets=200, reads_per_target=12, max_units=40, err=0.004)
    monkeypatch.setenv("TDB_QUAD8", "1")
    c = tdb.Context()
    cnt = _check_batch(c, seqs, pp, pt, 50)
    assert cnt.tier_pairs[0] > 0 and cnt.tier_pairs[1] > 0
    c.close()
    monkeypatch.setenv("TDB_DISABLE_THREAD_TIER", "1")
    c = tdb.Context()
    monkeypatch.delenv("TDB_DISABLE_THREAD_TIER")
    monkeypatch.delenv("TDB_QUAD8")
    cnt = _check_batch(c, seqs, pp, pt, 0, cigars=False)
    assert cnt.thread_pairs == 0 and cnt.tier_pairs[0] > 0
    c.close()


def test_thread_tier_disabled_gives_same_scores(tdb, monkeypatch):
    """The thread-per-pair tier takes the lightest pairs (final penalty <= 24) in front of the lockstep tier: the
    lockstep tier alone must give the same scores and work counts, and with it on most light pairs end there."""
    rng = random.Random(38)
    seqs, pp, pt = make_pairs(rng, n_targets=150, reads_per_target=12, max_units=25, err=0.006)
    monkeypatch.setenv("TDB_DISABLE_THREAD_TIER", "1")
    c = tdb.Context()
    monkeypatch.delenv("TDB_DISABLE_THREAD_TIER")
    cnt = _check_batch(c, seqs, pp, pt, 50)
    assert cnt.thread_pairs == 0 and cnt.tier_pairs[0] > 0
    c.close()
    c = tdb.Context()
    cnt2 = _check_batch(c, seqs, pp, pt, 50)
    assert cnt2.thread_pairs > 0.3 * cnt.tier_pairs[0]
    assert cnt2.thread_pairs + sum(cnt2.tier_pairs) == sum(cnt.tier_pairs)
    # every length from empty upwards, tiny clip windows, heuristic off: the tier's range guards
    seqs, pp, pt = [], [], []
    for L in list(range(0, 40)) + [63, 64, 65, 200]:
        for _ in range(6):
            t = rand_seq(rng, L)
            p = mutate(rng, t, rng.choice([0.02, 0.05, 0.1])) if L else rng.choice(["", "A", "ACG"])
            pt.append(len(seqs)); seqs.append(t.encode())
            pp.append(len(seqs)); seqs.append(p.encode())
    for clip in (0, 3, 50):
        _check_batch(c, seqs, np.array(pp, np.uint32), np.array(pt, np.uint32), clip)
    c.close()
    c = tdb.Context(heuristic=(0, 0, 0, 0))
    seqs, pp, pt = make_pairs(rng, n_targets=60, reads_per_target=8, max_units=30, err=0.01)
    cnt3 = _check_batch(c, seqs, pp, pt, 50, heuristic=tdo.NO_HEURISTIC)
    assert cnt3.thread_pairs > 0
    c.close()


def test_tier1_variants_give_same_scores(tdb, monkeypatch):
    """Tier 1 has two implementations (two pairs per warp in lockstep -- the default with the production penalties --
    and one pair per warp): both must agree with the oracle on wide-wavefront pairs."""
    rng = random.Random(36)
    seqs, pp, pt = [], [], []
    for _ in range(200):
        lf, rf, motif = rand_seq(rng, 50), rand_seq(rng, 50), rand_seq(rng, rng.randint(2, 6))
        t = lf + motif * rng.randint(10, 60) + rf
        for _ in range(4):
            u = rng.randint(16, 45) // len(motif) + 1
            src = lf + motif * max(1, t.count(motif) + rng.choice([-u, u])) + rf
            pp.append(len(seqs) + 1); pt.append(len(seqs))
            seqs.append(t.encode()); seqs.append(mutate(rng, src, 0.004).encode())
    pp, pt = np.array(pp, np.uint32), np.array(pt, np.uint32)
    c = tdb.Context()
    cnt = _check_batch(c, seqs, pp, pt, 50, cigars=False)
    assert cnt.tier_pairs[1] > 100
    c.close()
    monkeypatch.setenv("TDB_DISABLE_DUO", "1")
    c = tdb.Context()
    monkeypatch.delenv("TDB_DISABLE_DUO")
    cnt2 = _check_batch(c, seqs, pp, pt, 50, cigars=False)
    assert cnt2.tier_pairs[1] > 100
    c.close()


def test_dedup_disabled_gives_same_scores(tdb, monkeypatch):
    """Duplicate elimination must be invisible: every pair aligned on its own gives the same scores."""
    rng = random.Random(32)
    seqs, pp, pt = make_pairs(rng, n_targets=60, reads_per_target=10, max_units=25, err=0.001)
    monkeypatch.setenv("TDB_DISABLE_DEDUP", "1")
    c = tdb.Context()
    monkeypatch.delenv("TDB_DISABLE_DEDUP")
    cnt = _check_batch(c, seqs, pp, pt, 50)
    assert cnt.pairs_duplicate == 0
    c.close()


def test_duplicate_heavy_batch_counts(ctx):
    """Error-free reads repeat the allele sequence: most pairs are duplicates or identical, and the
    algorithmic work counters still equal the oracle's (every pair counted)."""
    rng = random.Random(33)
    seqs, pp, pt = make_pairs(rng, n_targets=80, reads_per_target=15, max_units=20, err=0.0005)
    cnt = _check_batch(ctx, seqs, pp, pt, 50)
    assert cnt.pairs_duplicate + cnt.pairs_identical > 0.3 * len(pp) and cnt.pairs_duplicate > 0
    assert cnt.exec_cells < cnt.cells and cnt.exec_columns < cnt.columns


@pytest.mark.parametrize("clip", [50, 0, 7])
def test_closed_form_pairs(tdb, clip):
    """Single substitution / single gap of 1..5 bases anywhere (inside and outside the clip window, at the very
    ends, inside repeats where the gap can slide): the closed form must equal the wavefront result."""
    rng = random.Random(40 + clip)
    seqs, pp, pt = [], [], []
    for _ in range(3000):
        k = rng.randint(1, 6)
        motif = rand_seq(rng, k)
        lf, rf = rand_seq(rng, rng.randint(0, 60)), rand_seq(rng, rng.randint(0, 60))
        t = lf + motif * rng.randint(0, 90) + rf or "A"  # repeats long enough for wf-adaptive to act
        mode = rng.random()
        if mode < 0.2:
            i = rng.randrange(len(t))
            p = t[:i] + rng.choice("ACGT") + t[i + 1:]
        elif mode < 0.5:  # whole motif units in or out (the gap can slide inside the repeat)
            u = rng.randint(1, max(1, 22 // k))
            j = len(lf) + k * rng.randint(0, 5)
            p = t[:j] + motif * u + t[j:] if rng.random() < 0.5 else t[:j] + t[j + k * u:]
        elif mode < 0.9:
            i, d = rng.randrange(len(t) + 1), rng.randint(1, 24)  # beyond CF_MAXGAP = 20 too
            p = t[:i] + rand_seq(rng, d) + t[i:] if rng.random() < 0.5 else t[:i] + t[i + d:]
        else:
            p = mutate(rng, t, 0.02)
        pt.append(len(seqs)); seqs.append(t.encode())
        pp.append(len(seqs)); seqs.append((p or "C").encode())
    c = tdb.Context()
    cnt = _check_batch(c, seqs, np.array(pp, np.uint32), np.array(pt, np.uint32), clip, cigars=False)
    c.set_count_work(False)
    blob, off, ln = to_arrays(seqs)
    c.align_batch(blob, off, ln, np.array(pp, np.uint32), np.array(pt, np.uint32))
    assert c.counters().pairs_closed_form > 1000
    c.close()


def test_chunked_host_pipeline(tdb, monkeypatch):
    """Many small pipeline chunks (upload of chunk k+1 overlapping the alignment of chunk k)."""
    rng = random.Random(34)
    seqs, pp, pt = make_pairs(rng, n_targets=400, reads_per_target=12, max_units=20, err=0.002)
    monkeypatch.setenv("TDB_CHUNK_PAIRS", "1024")
    c = tdb.Context()
    monkeypatch.delenv("TDB_CHUNK_PAIRS")
    cnt = _check_batch(c, seqs, pp, pt, 50, cigars=False)
    assert cnt.chunks >= 4
    # a pair list that jumps around the sequence table falls back to one chunk, same scores
    perm = np.array(random.Random(1).sample(range(len(pp)), len(pp)))
    cnt = _check_batch(c, seqs, pp[perm], pt[perm], 50, cigars=False)
    assert cnt.chunks == 1
    c.close()


def test_host_packed_batches(tdb, monkeypatch):
    """Host batches 2-bit packed by the caller's CPU threads (default for batches >= 8 MB): same scores,
    including sequences with non-ACGT bytes whose raw bytes are shipped separately."""
    monkeypatch.setenv("TDB_HOSTPACK_MIN_BYTES", "0")
    monkeypatch.setenv("TDB_CHUNK_PAIRS", "1024")
    c = tdb.Context()
    monkeypatch.delenv("TDB_HOSTPACK_MIN_BYTES")
    monkeypatch.delenv("TDB_CHUNK_PAIRS")
    rng = random.Random(35)
    seqs, pp, pt = make_pairs(rng, n_targets=300, reads_per_target=12, max_units=20, err=0.002)
    cnt = _check_batch(c, seqs, pp, pt, 50, cigars=False)
    assert cnt.host_packed == 1 and cnt.chunks >= 3
    seqs, pp, pt = make_pairs(rng, n_targets=40, reads_per_target=6, alphabet="ACGTN")  # few flagged: per-sequence copies
    seqs[0] = seqs[0].lower()
    cnt = _check_batch(c, seqs, pp, pt, 50, cigars=False)
    assert cnt.host_packed == 1
    seqs, pp, pt = make_pairs(rng, n_targets=100, reads_per_target=6, err=0.2, alphabet="ACGTN")  # > 256 flagged: whole blob
    _check_batch(c, seqs, pp, pt, 50, cigars=False)
    # every 2nd chunk left to the GPU as ASCII (what thread-starved ranks do automatically)
    monkeypatch.setenv("TDB_HOSTPACK_MIN_BYTES", "0")
    monkeypatch.setenv("TDB_CHUNK_PAIRS", "1024")
    monkeypatch.setenv("TDB_RAW_EVERY", "2")
    monkeypatch.setenv("TDB_TEST_RAW_BLOCKS", "2")  # deterministic: every 2nd 256 KB block is never packed on the host
    c2 = tdb.Context()
    seqs, pp, pt = make_pairs(rng, n_targets=600, reads_per_target=12, max_units=30, err=0.002, alphabet="ACGT")
    seqs[7] = seqs[7][:30] + b"N" + seqs[7][31:]
    seqs[len(seqs) // 2] = seqs[len(seqs) // 2].lower()
    # sequences straddling a 256 KB block boundary with their only odd byte on the GPU-packed side: the device flags
    # them, and their bytes in the neighbouring host-packed block must reach the device too
    _, off, ln = to_arrays(seqs)
    n_straddle = 0
    for bi, B in enumerate(range(262144, int(off[-1]), 262144)):
        i = int(np.searchsorted(off, B, side="right")) - 1
        if off[i] < B < off[i] + ln[i] - 1:
            raw_is_next = bi % 2 == 0  # blocks 1, 3, .. are left to the GPU
            x = int(B - off[i]) if raw_is_next else int(B - off[i]) - 1
            seqs[i] = seqs[i][:x] + b"N" + seqs[i][x + 1:]
            n_straddle += 1
    assert n_straddle >= 2
    cnt = _check_batch(c2, seqs, pp, pt, 50, cigars=False)
    for v in ("TDB_HOSTPACK_MIN_BYTES", "TDB_CHUNK_PAIRS", "TDB_RAW_EVERY", "TDB_TEST_RAW_BLOCKS"):
        monkeypatch.delenv(v)
    assert cnt.host_packed == 1 and cnt.chunks >= 4 and cnt.bytes_raw > 0
    c2.close()
    blob = np.frombuffer(b"ACGTACGTAC" * 4, dtype=np.uint8)
    with pytest.raises(tdb.TdbError):  # layout violations are caught by the host packer too
        c.align_batch(blob, np.array([0, 5], np.uint64), np.array([10, 10], np.uint32), np.array([0], np.uint32),
                      np.array([1], np.uint32))
    c.close()


def test_packed_fused_calls_match_ascii(tdb, monkeypatch):
    """tdb_trio_batch_packed / tdb_duo_batch_packed == tdb_trio_batch / tdb_duo_batch on the same family data
    (chunked pipeline, sequences with bytes outside ACGT listed as exceptions), and less H2D traffic."""
    from importlib import import_module
    S = import_module("trgt_denovo_b200.synth")
    monkeypatch.setenv("TDB_CHUNK_PAIRS", "4096")
    c = tdb.Context()
    monkeypatch.delenv("TDB_CHUNK_PAIRS")
    for duo in (False, True):
        b = S.generate(S.config("duo_genome" if duo else "trio_genome"), 1000, 1400, 4)
        blob = b.blob.copy()
        for i in (3, 500, b.n_seqs - 1):  # N and lower case: raw byte equality must survive the packed path
            o = int(b.seq_off[i])
            blob[o + 5] = ord("N")
            blob[o + 7] = blob[o + 7] | 0x20
        out_a, sc_a, mk_a = c.assess_batch(blob, None, b.seq_len, b.pair_pattern, b.pair_text, b.loci, duo=duo)
        h2d_ascii = c.counters().bytes_h2d
        pk = tdb.pack_sequences(blob, None, b.seq_len, n_threads=4)
        assert len(pk.exc_seq) == 3
        out_p, sc_p, mk_p = c.assess_batch_packed(pk, b.pair_pattern, b.pair_text, b.loci, duo=duo)
        cnt = c.counters()
        assert cnt.host_packed == 2 and cnt.chunks > 2 and cnt.pairs_failed == 0
        assert cnt.bytes_h2d < 0.5 * h2d_ascii
        assert (sc_a == sc_p).all() and (mk_a == mk_p).all()
        assert bytes(out_a) == bytes(out_p)
        want, _, _ = tdo.align_batch(blob, b.seq_off, b.seq_len, b.pair_pattern, b.pair_text, 50, n_threads=8)
        assert (want == sc_p).all()
        b.close()
    c.close()


def test_failed_pairs_stay_out_of_the_reductions(tdb, monkeypatch):
    """A pair no tier can finish has no score (status OOM, TDB_SCORE_FAILED); the fused call drops it from every
    list as the reference drops non-completed alignments (src/denovo.rs:31-35) instead of counting a perfect 0."""
    import ctypes as C
    rng = random.Random(91)
    lf, rf = rand_seq(rng, 50), rand_seq(rng, 50)
    target = (lf + "CAG" * 20 + rf).encode()
    far = rand_seq(rng, 160).encode()           # unrelated read: penalty far above the test cap
    near = (lf + "CAG" * 19 + rf).encode()      # one unit short
    seqs = [target, far, near, target, near, far]
    # lists: mother a0 = {far, near}, father a0 = {near}, child own = {target copy, near, far}
    pp = np.array([1, 2, 4, 3, 2, 5], np.uint32)
    pt = np.zeros(6, np.uint32)
    monkeypatch.setenv("TDB_TEST_MAX_SCORE", "200")
    c = tdb.Context()
    monkeypatch.delenv("TDB_TEST_MAX_SCORE")
    blob, off, ln = to_arrays(seqs)
    got, st = c.align_batch(blob, off, ln, pp, pt)
    assert c.counters().pairs_failed == 2
    assert list(st) == [-200, 0, 0, 0, 0, -200]
    from importlib import import_module
    assert got[0] == got[5] == import_module("trgt_denovo_b200.binding").SCORE_FAILED
    loc = tdb.TrioLocus()
    loc.n_child, loc.n_mother, loc.n_father = 1, 1, 1
    for j, v in enumerate([0, 2, 2, 3, 3, 6]):
        loc.list_off[0][j] = v
    arr = (tdb.TrioLocus * 1)(loc)
    out, scores, mask = c.assess_batch(blob, off, ln, pp, pt, arr)
    # the same locus with the failed pairs removed, on the oracle
    loc2, flat = tdo.make_locus([[[int(got[1])]]], [[[int(got[2])]]], [[int(got[3]), int(got[4])]])
    want, wmask = tdo.trio_assess(loc2, flat, 1.0)
    g, w = out[0], want[0]
    assert (g.denovo_coverage, g.error) == (w.denovo_coverage, w.error)
    assert g.top_mother == w.top_mother and g.top_father == w.top_father and g.child_median == w.child_median
    assert list(g.mother_overlap) == list(w.mother_overlap) and list(g.father_overlap) == list(w.father_overlap)
    assert list(g.mean_col) == list(w.mean_col)
    assert list(mask[3:6]) == [int(wmask[loc2.list_off[0][4]]), int(wmask[loc2.list_off[0][4] + 1]), 0]
    # a list made only of failed pairs is an empty list: the reference would panic -> error flag
    loc.list_off[0][0], loc.list_off[0][1], loc.list_off[0][2] = 0, 1, 1
    pp2 = np.array([1, 4, 3, 2, 5], np.uint32)
    loc.list_off[0][3], loc.list_off[0][4], loc.list_off[0][5] = 2, 2, 5
    arr = (tdb.TrioLocus * 1)(loc)
    out, _, _ = c.assess_batch(blob, off, ln, pp2, np.zeros(5, np.uint32), arr)
    assert out[0].error == 1
    c.close()


def test_dense_layout(tdb, ctx):
    """seq_off = NULL on the host and device entry points; gaps between sequences need explicit offsets; a dense
    call whose lengths overrun the blob fails."""
    rng = random.Random(37)
    seqs, pp, pt = make_pairs(rng, n_targets=50, reads_per_target=8, err=0.01)
    blob, off, ln = to_arrays(seqs)
    want, _, _ = tdo.align_batch(blob, off, ln, pp, pt, 50)
    ctx.set_clip_len(50)
    ctx.set_count_work(False)
    got, st = ctx.align_batch(blob, None, ln, pp, pt)
    assert (got == want).all() and (st == 0).all()
    g2, st2, pen2, cigs = ctx.align_batch_cigar(blob, None, ln, pp, pt)
    assert (g2 == want).all()
    # device-resident entry point
    import torch
    dev = lambda a: torch.from_numpy(np.array(a)).cuda()
    d_blob = torch.zeros(blob.size + 64, dtype=torch.uint8, device="cuda")
    d_blob[:blob.size] = dev(blob)
    d_ln, d_pp, d_pt = dev(ln.view(np.int32)), dev(pp.view(np.int32)), dev(pt.view(np.int32))
    d_sc = torch.zeros(len(pp), dtype=torch.int32, device="cuda")
    d_st = torch.zeros(len(pp), dtype=torch.int32, device="cuda")
    rc = tdb.lib().tdb_align_batch_device(ctx._h, d_blob.data_ptr(), blob.size, None, d_ln.data_ptr(), len(ln),
                                                  d_pp.data_ptr(), d_pt.data_ptr(), len(pp), d_sc.data_ptr(),
                                                  d_st.data_ptr())
    assert rc == 0
    assert (d_sc.cpu().numpy() == want).all()
    # explicit offsets with gaps between the sequences still work and differ from the dense reading
    gap_seqs = [s + b"TTTT" for s in seqs]
    gblob, goff, _ = to_arrays(gap_seqs)
    got, st = ctx.align_batch(gblob, goff, ln, pp, pt)
    assert (got == want).all()
    with pytest.raises(tdb.TdbError):  # dense lengths overrun the blob
        ctx.align_batch(blob[:blob.size // 2], None, ln, pp, pt)
    got, st = ctx.align_batch(blob, None, ln, pp, pt)
    assert (got == want).all()


def test_two_contexts_from_two_host_threads(tdb):
    """One ctx per host thread is the threading model (like the reference's thread-local aligner): two contexts on
    the same GPU, driven concurrently, each give the oracle's scores."""
    import threading
    rng = random.Random(39)
    jobs = []
    for t in range(2):
        seqs, pp, pt = make_pairs(rng, n_targets=200, reads_per_target=10, max_units=30, err=0.01)
        blob, off, ln = to_arrays(seqs)
        want, _, _ = tdo.align_batch(blob, off, ln, pp, pt, 50, n_threads=4)
        jobs.append((blob, ln, pp, pt, want))
    results, errors = [None, None], []

    def work(i):
        try:
            c = tdb.Context(device_id=0)
            for _ in range(5):
                got, st = c.align_batch(jobs[i][0], None, jobs[i][1], jobs[i][2], jobs[i][3])
                assert (st == 0).all() and (got == jobs[i][4]).all()
            results[i] = True
            c.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(repr(e))

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors and results == [True, True]


def test_bad_layout_and_bad_pair_index_fail(tdb, ctx):
    blob = np.frombuffer(b"ACGTACGTAC" * 4, dtype=np.uint8)
    ln = np.array([10, 10], np.uint32)
    pp, pt = np.array([0], np.uint32), np.array([1], np.uint32)
    with pytest.raises(tdb.TdbError):  # overlapping sequences
        ctx.align_batch(blob, np.array([0, 5], np.uint64), ln, pp, pt)
    with pytest.raises(tdb.TdbError):  # outside the blob
        ctx.align_batch(blob, np.array([0, 35], np.uint64), ln, pp, pt)
    with pytest.raises(tdb.TdbError):  # pair index out of range
        ctx.align_batch(blob, np.array([0, 10], np.uint64), ln, pp, np.array([2], np.uint32))
    got, st = ctx.align_batch(blob, np.array([0, 10], np.uint64), ln, pp, pt)  # the ctx is still usable
    assert st[0] == 0


def test_heuristic_none(tdb):
    rng = random.Random(9)
    seqs, pp, pt = make_pairs(rng, n_targets=60, reads_per_target=6, err=0.05)
    c = tdb.Context(heuristic=(0, 0, 0, 0))
    _check_batch(c, seqs, pp, pt, 50, heuristic=tdo.NO_HEURISTIC)
    c.close()


def test_non_acgt_bytes_take_byte_path(ctx):
    """Raw byte equality: N==N matches, lower case differs from upper case (SURVEY.md hard parts)."""
    rng = random.Random(21)
    seqs, pp, pt = make_pairs(rng, n_targets=40, reads_per_target=6, alphabet="ACGTN")
    seqs[0] = seqs[0].lower()
    seqs[3] = seqs[3][:20] + b"nnRYK" + seqs[3][25:]
    _check_batch(ctx, seqs, pp, pt, 50)


def test_edge_cases(ctx):
    seqs = [b"A", b"C", b"ACGT", b"ACGT", b"ACGTACGTACGTACGTA", b"ACGTACGTACGTACGT", b"G" * 300, b"G" * 299 + b"T",
            b"T", b"ACGT" * 60, b"ACGT" * 40]
    pairs = [(0, 1), (2, 3), (4, 5), (5, 4), (6, 7), (7, 6), (8, 6), (6, 8), (9, 10), (10, 9), (0, 0)]
    pp = np.array([p for p, _ in pairs], np.uint32)
    pt = np.array([t for _, t in pairs], np.uint32)
    for clip in (0, 3, 50):
        _check_batch(ctx, seqs, pp, pt, clip)


def test_empty_and_very_long_sequences(ctx):
    """Zero-length sequences (pure gap alignments) and sequences beyond the int16 tiers' 8000-base limit."""
    rng = random.Random(55)
    long_t = rand_seq(rng, 9000)
    seqs = [b"", b"ACGT", b"", b"A" * 30, b"A", long_t.encode(), mutate(rng, long_t, 0.01).encode(),
            (long_t[:4000] + long_t[4300:]).encode(), b"ACGTTGCA" * 1200, b"ACGTTGCA" * 1190]
    pairs = [(0, 1), (1, 0), (0, 2), (0, 3), (3, 0), (4, 0), (6, 5), (5, 6), (7, 5), (5, 7), (9, 8), (8, 9), (5, 5)]
    pp = np.array([p for p, _ in pairs], np.uint32)
    pt = np.array([t for _, t in pairs], np.uint32)
    for clip in (50, 0):
        cnt = _check_batch(ctx, seqs, pp, pt, clip)
    assert cnt.long_pairs + cnt.tier_pairs[2] + cnt.tier_pairs[3] >= 4
    assert cnt.long_pairs > 0  # up to 16 000 bases and 192 diagonals: the long tier


def test_empty_batch(ctx):
    z8, z4 = np.zeros(0, np.uint64), np.zeros(0, np.uint32)
    got, st = ctx.align_batch(np.zeros(0, np.uint8), z8, z4, z4, z4)
    assert len(got) == 0 and len(st) == 0


def test_long_expansion_pairs(ctx):
    """Config-4 style: 300 bp allele reads vs kb-scale expansions (long wavefronts, second gap piece)."""
    rng = random.Random(77)
    seqs, pp, pt = [], [], []
    for _ in range(12):
        lf, rf, motif = rand_seq(rng, 50), rand_seq(rng, 50), rand_seq(rng, rng.randint(2, 6))
        big = lf + motif * rng.randint(300, 900) + rf
        small = lf + motif * rng.randint(20, 60) + rf
        t = len(seqs); seqs.append(big.encode())
        for src in (big, big, small):
            pp.append(len(seqs)); pt.append(t); seqs.append(mutate(rng, src, 0.003).encode())
    _check_batch(ctx, seqs, np.array(pp, np.uint32), np.array(pt, np.uint32), 50)


def _long_pairs(seed):
    rng = random.Random(seed)
    seqs, pp, pt = [], [], []
    for i in range(10):
        lf, rf, motif = rand_seq(rng, 50), rand_seq(rng, 50), rand_seq(rng, rng.randint(2, 6))
        big = lf + motif * rng.randint(400, 1500) + rf
        small = lf + motif * rng.randint(20, 200) + rf
        t = len(seqs); seqs.append(big.encode())
        u = len(seqs); seqs.append(small.encode())
        for src, tgt in ((big, t), (small, t), (big, u), (small, u)):  # long deletions AND long insertions (both gap pieces)
            pp.append(len(seqs)); pt.append(tgt); seqs.append(mutate(rng, src, 0.004).encode())
    # two interrupted repeats: gaps of several hundred columns split by matches (runs of the backtrace start and stop)
    a = rand_seq(rng, 60) + "CAG" * 300 + "CAA" + "CAG" * 200 + rand_seq(rng, 60)
    b = rand_seq(rng, 5) + a[5:60] + "CAG" * 40 + "CAA" + "CAG" * 350 + a[-60:]
    pp.append(len(seqs)); seqs.append(a.encode()); pt.append(len(seqs)); seqs.append(b.encode())
    pp.append(len(seqs) - 1); pt.append(len(seqs) - 2)
    return seqs, np.array(pp, np.uint32), np.array(pt, np.uint32)


def test_long_tier_and_global_tiers_give_same_scores(tdb, monkeypatch):
    """Long gaps (hundreds to thousands of columns, insertions and deletions): the long tier (k_align_lock<LongCfg>:
    shared-memory wavefronts, gap runs of the backtrace walked 32 columns per round) and, with the long tier off, the
    global-scratch tiers must both reproduce the oracle."""
    seqs, pp, pt = _long_pairs(2024)
    c = tdb.Context()
    cnt = _check_batch(c, seqs, pp, pt, 50)
    assert cnt.long_pairs > 20 and cnt.tier_pairs[2] + cnt.tier_pairs[3] == 0
    c.close()
    monkeypatch.setenv("TDB_LONG_SORT_PER_WARP", "0")  # the list sorted heaviest first (k_sort_long) whatever its length
    c = tdb.Context()
    monkeypatch.delenv("TDB_LONG_SORT_PER_WARP")
    cnt1 = _check_batch(c, seqs, pp, pt, 50, cigars=False)
    assert cnt1.long_pairs == cnt.long_pairs
    c.close()
    monkeypatch.setenv("TDB_DISABLE_LONG", "1")
    c = tdb.Context()
    monkeypatch.delenv("TDB_DISABLE_LONG")
    cnt2 = _check_batch(c, seqs, pp, pt, 50)
    assert cnt2.long_pairs == 0 and cnt2.tier_pairs[2] == cnt.long_pairs
    c.close()


def _random_loci(rng, n_loci, duo=False):
    loci, flat = [], []
    for _ in range(n_loci):
        nc, nm = rng.randint(1, 2), rng.randint(1, 2)
        nf = 0 if duo else rng.randint(1, 2)
        loc = tdo.TrioLocus()
        loc.n_child, loc.n_mother, loc.n_father = nc, nm, nf
        for c in range(nc):
            sizes = [rng.randint(0, 25), rng.randint(0, 25) if nm > 1 else 0,
                     rng.randint(0, 25) if nf > 0 else 0, rng.randint(0, 25) if nf > 1 else 0, rng.randint(0, 30)]
            if rng.random() < 0.7:
                sizes[0] = max(sizes[0], 1)
                if nf:
                    sizes[2] = max(sizes[2], 1)
            base = rng.choice([0, -6, -20, -60])
            for j, sz in enumerate(sizes):
                loc.list_off[c][j] = len(flat)
                flat.extend(int(base + rng.choice([0, 0, 0, -2, -4, -6, -8, -12, -30, -100]) * rng.random()) for _ in range(sz))
            loc.list_off[c][5] = len(flat)
        for i in range(2):
            loc.child_len[i], loc.mother_len[i], loc.father_len[i] = (rng.randint(100, 400) for _ in range(3))
        loci.append(loc)
    return loci, np.array(flat, dtype=np.int32)


FIELDS_INT = ["denovo_coverage", "allele_origin", "denovo_status", "error"]
FIELDS_F = ["mean_diff_father", "mean_diff_mother", "top_mother", "top_father", "child_median"]


def _close(a, b, tol=1e-9):
    if a == b:
        return True
    if np.isnan(a) and np.isnan(b):
        return True
    return abs(a - b) <= tol * max(abs(a), abs(b))


@pytest.mark.parametrize("duo", [False, True])
@pytest.mark.parametrize("q", [1.0, 0.9, 0.5, 0.0])
def test_reduction_matches_oracle(tdb, ctx, duo, q):
    rng = random.Random(1000 + int(q * 10) + duo)
    loci, scores = _random_loci(rng, 400, duo=duo)
    arr = (tdb.TrioLocus * len(loci))()
    for i, l in enumerate(loci):
        C = __import__("ctypes")
        C.memmove(C.byref(arr[i]), C.byref(l), C.sizeof(l))
    out, mask = ctx.reduce(arr, scores, q=q, duo=duo)
    for i, l in enumerate(loci):
        want, wmask = tdo.trio_assess(l, scores, q, duo=duo)
        for c in range(l.n_child):
            g, w = out[2 * i + c], want[c]
            for f in FIELDS_INT:
                assert getattr(g, f) == getattr(w, f), (i, c, f)
            if w.error:
                continue
            for f in FIELDS_F:
                assert _close(getattr(g, f), getattr(w, f)), (i, c, f, getattr(g, f), getattr(w, f))
            assert list(g.mother_overlap) == list(w.mother_overlap)
            assert list(g.father_overlap) == list(w.father_overlap)
            for r in range(4):
                assert _close(g.mean_col[r], w.mean_col[r]) or duo
            o4, o5 = l.list_off[c][4], l.list_off[c][5]
            assert (mask[o4:o5] == wmask[o4:o5]).all()


@pytest.mark.parametrize("q", [1.0, 0.75, 0.5])
def test_reduction_deep_coverage_lists(tdb, ctx, q):
    """Lists of hundreds to thousands of scores per allele (amplicon-depth coverage): the k-th smallest behind
    quantile / median switches from count-and-compare to a radix select at 64 entries; both must give the oracle's rows
    (src/math.rs:82-155).  Sizes straddle the switch, odd and even, with ties and with failed-pair sentinels."""
    import ctypes as C
    from importlib import import_module
    failed = import_module("trgt_denovo_b200.binding").SCORE_FAILED
    rng = random.Random(4242 + int(q * 100))
    loci, flat = [], []
    for sizes in ([63, 64, 65, 66, 64], [1500, 1, 777, 0, 2001], [64, 0, 0, 128, 4000], [200, 201, 202, 203, 204]):
        loc = tdo.TrioLocus()
        loc.n_child, loc.n_mother, loc.n_father = 1, 2, 2
        base = rng.choice([0, -8, -40])
        for j, sz in enumerate(sizes):
            loc.list_off[0][j] = len(flat)
            flat.extend(int(base - rng.choice([0, 0, 2, 4, 6, 8, 12, 30, 100, 4000]) * rng.random()) for _ in range(sz))
        loc.list_off[0][5] = len(flat)
        for i in range(2):
            loc.child_len[i], loc.mother_len[i], loc.father_len[i] = (rng.randint(100, 400) for _ in range(3))
        loci.append(loc)
    scores = np.array(flat, dtype=np.int32)
    for variant in range(2):
        if variant == 1:  # a sprinkle of pairs no tier could finish: dropped from every list (src/denovo.rs:31-35)
            scores = scores.copy()
            scores[rng.sample(range(len(scores)), 200)] = failed
        arr = (tdb.TrioLocus * len(loci))()
        for i, l in enumerate(loci):
            C.memmove(C.byref(arr[i]), C.byref(l), C.sizeof(l))
        out, mask = ctx.reduce(arr, scores, q=q)
        for i, l in enumerate(loci):
            oscores = scores if variant == 0 else None
            if variant == 1:  # the oracle has no sentinel: hand it the lists with the failed entries removed
                keep, lo2 = [], tdo.TrioLocus()
                C.memmove(C.byref(lo2), C.byref(l), C.sizeof(l))
                for j in range(5):
                    a, b = l.list_off[0][j], l.list_off[0][j + 1]
                    lo2.list_off[0][j] = len(keep)
                    keep.extend(int(x) for x in scores[a:b] if x != failed)
                lo2.list_off[0][5] = len(keep)
                want, _ = tdo.trio_assess(lo2, np.array(keep, dtype=np.int32), q)
            else:
                want, wmask = tdo.trio_assess(l, oscores, q)
            g, w = out[2 * i], want[0]
            for f in FIELDS_INT:
                assert getattr(g, f) == getattr(w, f), (variant, i, f)
            for f in FIELDS_F:
                assert _close(getattr(g, f), getattr(w, f)), (variant, i, f, getattr(g, f), getattr(w, f))
            assert list(g.mother_overlap) == list(w.mother_overlap) and list(g.father_overlap) == list(w.father_overlap)
            for r in range(4):
                assert _close(g.mean_col[r], w.mean_col[r])
            if variant == 0:
                o4, o5 = l.list_off[0][4], l.list_off[0][5]
                assert (mask[o4:o5] == wmask[o4:o5]).all()


def test_reduction_reference_vectors(tdb, ctx, golden_dir):
    r = json.load(open(os.path.join(golden_dir, "reduction_vectors.json")))
    import ctypes as C
    for case in r["trio_get_denovo_coverage"]:
        loc, scores = tdo.make_locus([case["mother"]], [case["father"]], [case["child"]])
        arr = (tdb.TrioLocus * 1)()
        C.memmove(C.byref(arr[0]), C.byref(loc), C.sizeof(loc))
        out, mask = ctx.reduce(arr, scores, q=case["q"])
        e = case["expect"]
        assert out[0].denovo_coverage == e["denovo_coverage"]
        assert out[0].mean_diff_father == e["mean_diff_father"] and out[0].mean_diff_mother == e["mean_diff_mother"]
        o4 = loc.list_off[0][4]
        assert list(mask[o4:o4 + 3]) == e["mask"]
    m = np.array(r["parent_origin_matrix"]["matrix_rowmajor_4x2"]).reshape(4, 2)
    mother = [[[int(m[0, c])], [int(m[1, c])]] for c in range(2)]
    father = [[[int(m[2, c])], [int(m[3, c])]] for c in range(2)]
    loc, scores = tdo.make_locus(mother, father, [[0], [0]])
    arr = (tdb.TrioLocus * 1)()
    C.memmove(C.byref(arr[0]), C.byref(loc), C.sizeof(loc))
    out, _ = ctx.reduce(arr, scores, q=1.0)
    assert [out[0].allele_origin, out[1].allele_origin] == r["parent_origin_matrix"]["expect_origin"]
