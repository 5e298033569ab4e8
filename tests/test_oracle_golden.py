"""CPU: pin the oracle (oracle/) on the reference's own known-answer vectors.

Vectors: tests/golden/*.json, extracted from /root/reference/src/aligner.rs tests (SURVEY.md
Appendix B), src/trio/denovo.rs tests and src/math.rs tests by tests/golden/make_golden.py.
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import tdo
from tests.util import rand_seq


def _vectors(golden_dir):
    return json.load(open(os.path.join(golden_dir, "aligner_vectors.json")))


def test_aligner_known_answers(golden_dir):
    for x in _vectors(golden_dir):
        heur = tdo.DEFAULT_HEURISTIC if x["heuristic"] == "default" else tdo.NO_HEURISTIC
        a = tdo.Aligner(tuple(x["penalties"]), heur)
        st = a.align_end_to_end(x["pattern"].encode(), x["text"].encode())
        assert st == tdo.STATUS_COMPLETED, x["id"]
        assert a.score() == x["score"], x["id"]
        if "cigar" in x:
            assert a.cigar() == x["cigar"], x["id"]
        for clip, exp in (x.get("clipped") or {}).items():
            assert a.cigar_score_clipped(int(clip)) == exp, (x["id"], clip)
        for clip, exp in (x.get("clipped_cigar") or {}).items():
            assert a.cigar(int(clip)) == exp, (x["id"], clip)
        # the optimal score is independent of WFA: Gotoh DP must agree where the heuristic is inert
        assert tdo.gotoh2p_penalty(x["pattern"].encode(), x["text"].encode(), tuple(x["penalties"])) == -x["score"]


def test_b9_default_heuristic_regression(golden_dir):
    """Restatement-derived (NOT reference-pinned): wf-adaptive changes B9's path."""
    x = [v for v in _vectors(golden_dir) if v["id"] == "B9"][0]
    a = tdo.Aligner()
    assert a.align_end_to_end(x["pattern"].encode(), x["text"].encode()) == 0
    assert a.score() == -1007
    assert a.cigar_score_clipped(50) == -969


def _mutate(rng, s, div):
    out = []
    for ch in s:
        r = rng.random()
        if r < div / 3:
            out.append(rng.choice("ACGT"))
        elif r < 2 * div / 3:
            pass
        elif r < div:
            out.append(ch)
            out.append(rng.choice("ACGT"))
        else:
            out.append(ch)
    return "".join(out) or "A"


def test_random_pairs_optimal_and_self_consistent():
    rng = random.Random(7)
    pens = [(8, 4, 2, 24, 1), (6, 2, 2, 4, 1), (4, 6, 2, 20, 1)]
    for it in range(300):
        L = rng.randint(1, 300)
        if rng.random() < 0.5:
            motif = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 6)))
            t = ("".join(rng.choice("ACGT") for _ in range(20)) + motif * (L // len(motif) + 1)
                 + "".join(rng.choice("ACGT") for _ in range(20)))
        else:
            t = "".join(rng.choice("ACGT") for _ in range(L))
        p = _mutate(rng, t, rng.choice([0, 0.01, 0.05, 0.2, 0.5]))
        pen = rng.choice(pens)
        for heur in (tdo.NO_HEURISTIC, tdo.DEFAULT_HEURISTIC):
            a = tdo.Aligner(pen, heur)
            assert a.align_end_to_end(p.encode(), t.encode()) == 0
            ops = a.cigar_ops()
            v = h = 0
            for c in ops.decode():
                if c == "M":
                    assert p[v] == t[h]
                    v, h = v + 1, h + 1
                elif c == "X":
                    assert p[v] != t[h]
                    v, h = v + 1, h + 1
                elif c == "I":
                    h += 1
                else:
                    v += 1
            assert (v, h) == (len(p), len(t))
            assert a.cigar_score_clipped(0) == a.score()
            g = tdo.gotoh2p_penalty(p.encode(), t.encode(), pen)
            if heur is tdo.NO_HEURISTIC:
                assert -a.score() == g
            else:
                assert -a.score() >= g


def test_closed_form_territory_is_heuristic_inert():
    """The product scores one substitution / one gap of <= 20 bases in closed form (tdb_kernels.cuh: closed_form).  That
    shortcut must not rest on the restated wf-adaptive cut-off: on such pairs -- gaps inside repeats (where they can slide),
    in the flanks, at the very ends; insertions and deletions -- the checker gives the SAME CIGAR with the heuristic on and
    off, its penalty is the optimum of the independent Gotoh DP and equals 8 resp. min(4 + 2L, 24 + L).  (For L = 21 the
    heuristic changes the penalty -- 46 through the first gap piece instead of 45 -- but not the CIGAR, and only the
    CIGAR enters the clipped score: that case is left to the wavefront kernels.)"""
    rng = random.Random(7)
    on, off = tdo.Aligner(), tdo.Aligner(heuristic=tdo.NO_HEURISTIC)
    seen = 0
    while seen < 1200:
        k = rng.randint(1, 6)
        motif = rand_seq(rng, k)
        lf, rf = rand_seq(rng, 50), rand_seq(rng, 50)
        t = lf + motif * rng.randint(3, 90) + rf
        L = rng.choice([0, 1, 2, 5, 10, 19, 20])
        if L == 0:  # one substitution
            i = rng.randrange(len(t))
            p = t[:i] + rng.choice([c for c in "ACGT" if c != t[i]]) + t[i + 1:]
        elif rng.random() < 0.5:  # whole motif units in or out
            u = max(1, L // k)
            L = u * k
            j = len(lf) + k * rng.randint(0, 3)
            p = t[:j] + motif * u + t[j:] if rng.random() < 0.5 else t[:j] + t[j + L:]
        else:
            i = rng.randrange(len(t) + 1)
            p = t[:i] + rand_seq(rng, L) + t[i:] if rng.random() < 0.5 else t[:i] + t[i + L:]
        if L > 20 or abs(len(p) - len(t)) != L or p == t or not p:
            continue
        pe, te = p.encode(), t.encode()
        assert on.align_end_to_end(pe, te) == 0 and off.align_end_to_end(pe, te) == 0
        g = tdo.gotoh2p_penalty(pe, te)
        if g != (8 if L == 0 else min(4 + 2 * L, 24 + L)):
            continue  # a random insertion that happens to align cheaper: not a single-event pair
        seen += 1
        assert on.cigar_ops() == off.cigar_ops(), (p, t)
        assert -on.score() == -off.score() == g, (p, t)
        for clip in (0, 50):
            assert on.cigar_score_clipped(clip) == off.cigar_score_clipped(clip)


def test_edge_cases():
    a = tdo.Aligner()
    assert a.align_end_to_end(b"ACGT", b"ACGT") == 0 and a.score() == 0 and a.cigar() == "4M"
    assert a.cigar_score_clipped(50) == 0  # empty window, SURVEY.md C2
    assert a.align_end_to_end(b"A", b"C") == 0 and a.score() == -8
    # N == N matches, case matters (raw byte equality, SURVEY.md "Alphabet semantics")
    assert a.align_end_to_end(b"ANNT", b"ANNT") == 0 and a.score() == 0
    assert a.align_end_to_end(b"GaT", b"GAT") == 0 and a.score() == -8
    # long gap uses the second affine piece: min(4+2L, 24+L)
    t = b"ACGTTGCATTAGCCAGTCAAGGCT" * 2
    p = t[:10] + t[40:]
    assert a.align_end_to_end(p, t) == 0 and a.score() == -(24 + 30)
    assert a.cigar() == "10M30I8M"


def test_reduction_known_answers(golden_dir):
    r = json.load(open(os.path.join(golden_dir, "reduction_vectors.json")))
    g = r["get_score_count_diff"]
    assert list(tdo.get_score_count_diff(g["top"], g["aligns"])) == g["expect"]
    for case in r["get_overlap_coverage"]:
        assert tdo.get_overlap_coverage(case["threshold"], case["lists"]) == case["expect"]
    for case in r["median"]:
        assert tdo.median(case["data"]) == case["expect"]
    for case in r["trio_get_denovo_coverage"]:
        loc, scores = tdo.make_locus([case["mother"]], [case["father"]], [case["child"]])
        out, mask = tdo.trio_assess(loc, scores, case["q"])
        e = case["expect"]
        assert out[0].denovo_coverage == e["denovo_coverage"]
        assert out[0].mean_diff_father == e["mean_diff_father"]
        assert out[0].mean_diff_mother == e["mean_diff_mother"]
        o4 = loc.list_off[0][4]
        assert list(mask[o4:o4 + len(case["child"])]) == e["mask"]


def test_parent_origin_matrix(golden_dir):
    """src/trio/denovo.rs:369-403: craft score lists whose means equal the test's matrix."""
    r = json.load(open(os.path.join(golden_dir, "reduction_vectors.json")))["parent_origin_matrix"]
    m = np.array(r["matrix_rowmajor_4x2"]).reshape(4, 2)
    mother = [[[int(m[0, c])], [int(m[1, c])]] for c in range(2)]
    father = [[[int(m[2, c])], [int(m[3, c])]] for c in range(2)]
    loc, scores = tdo.make_locus(mother, father, [[0], [0]])
    out, _ = tdo.trio_assess(loc, scores, 1.0)
    assert [out[0].allele_origin, out[1].allele_origin] == r["expect_origin"]
    assert out[0].mean_col[2] + out[1].mean_col[1] == r["expect_top_sum"]


def test_quantile_restatement():
    assert tdo.quantile([], 0.5) is None
    assert tdo.quantile([3.0, 1.0, 2.0], 1.0) == 3.0
    assert tdo.quantile([3.0, 1.0, 2.0], 0.5) == 2.0
    assert tdo.quantile([1.0, 2.0], 0.25) == 1.25
    assert tdo.quantile([-10.0, -20.0, -30.0, -40.0], 0.9) == pytest.approx(-13.0, rel=1e-15)


def test_batch_matches_single():
    rng = random.Random(3)
    seqs = []
    for _ in range(40):
        base = "".join(rng.choice("ACGT") for _ in range(50)) + "CAG" * rng.randint(3, 30) + "".join(
            rng.choice("ACGT") for _ in range(50))
        seqs.append(base.encode())
        seqs.append(_mutate(rng, base, 0.02).encode())
    blob = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    pp = np.array([rng.randrange(len(seqs)) for _ in range(200)], dtype=np.uint32)
    pt = np.array([rng.randrange(len(seqs)) for _ in range(200)], dtype=np.uint32)
    s1, st1, tot = tdo.align_batch(blob, offs, lens, pp, pt, 50, n_threads=1)
    s4, st4, _ = tdo.align_batch(blob, offs, lens, pp, pt, 50, n_threads=4)
    assert (s1 == s4).all() and (st1 == 0).all() and (st4 == 0).all()
    a = tdo.Aligner()
    for i in range(0, 200, 17):
        a.align_end_to_end(seqs[pp[i]], seqs[pt[i]])
        assert a.cigar_score_clipped(50) == s1[i]
    assert tot.cells > 0 and tot.columns > 0
