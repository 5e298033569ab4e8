"""CPU ORACLE for the per-locus pipeline (TEST INFRASTRUCTURE ONLY, see tdo_oracle.h).

Loop-for-loop restatement, one locus and one alignment at a time like the reference's worker thread, of
  load_alleles / assign_reads_by_alignment / assign_reads_by_classification   src/allele.rs:146-259
  trio process_alleles + quick mode          src/trio/allele.rs:71-349
  duo process_alleles + quick mode           src/duo/allele.rs:33-246
  align_alleleset / align_allele             src/denovo.rs:22-54
  serialize_with_precision                   src/allele.rs:32-50 (restated with exact decimal arithmetic)
  writer thread                              src/commands/shared.rs:161-195
Alignment = oracle Aligner (tdo_wfa.c); reductions = tdo_trio_assess / tdo_duo_assess (tdo_reduce.c).
Inputs are duck-typed: any objects with the attributes of trgt_denovo_b200.rows.SampleInput / Locus / Read.
PARITY: no reference test covers row assembly end to end ("parity unpinned" beyond restating the Rust).
"""
import ctypes as C
import math
from decimal import ROUND_HALF_EVEN, Decimal

import numpy as np

from . import tdo

ORIGIN = {0: "M:1", 1: "M:2", 2: "F:1", 3: "F:2", 4: "?", 5: "."}
STATUS = {0: "X", 1: "Y:+", 2: "Y:-", 3: "Y:=", 4: "Y:?", 5: "."}


def fmt_prec(value) -> str:
    v = float(value)
    if v != v:
        return "NaN"
    if v in (math.inf, -math.inf):
        return "inf" if v > 0 else "-inf"
    d = Decimal(v)  # exact binary value

    def fixed(places):
        q = d.quantize(Decimal(1).scaleb(-places), rounding=ROUND_HALF_EVEN)
        s = format(q, "f")
        return "-" + s.lstrip("-") if q.is_signed() and not s.startswith("-") else s

    s = fixed(1)
    if not s.endswith(".0"):
        s = fixed(3).rstrip("0")
        if s.endswith("."):
            s += "0"
    return s


class _Allele:
    def __init__(self, seq, mc, al, gt, index, read_aligns):
        self.seq, self.motif_count, self.allele_length, self.genotype, self.index = seq, mc, al, gt, index
        self.read_aligns = read_aligns


def load_alleles(sample, params, aligner):
    if sample is None:
        return None
    n_all = len(sample.allele_seqs)
    hp = [0, 0, 0]
    for r in sample.reads:
        if r.haplotype == 1:
            hp[0] += 1
        elif r.haplotype == 2:
            hp[1] += 1
        elif r.haplotype is None:
            hp[2] += 1
    by = [[] for _ in range(n_all)]
    if params.partition_by_alignment:
        index_flip = 0
        for read in sample.reads:
            max_score, max_aligns = None, []
            for i, a in enumerate(sample.allele_seqs):
                if aligner.align_end_to_end(read.bases, a) == 0:
                    score = aligner.cigar_score_clipped(params.clip_len)
                    if max_score is None or score > max_score:
                        max_score, max_aligns = score, [(i, score)]
                    elif score == max_score:
                        max_aligns.append((i, score))
            if max_aligns:
                if len(max_aligns) > 1:
                    index_flip = (index_flip + 1) % len(max_aligns)
                ai, al = max_aligns[index_flip % len(max_aligns)]
                by[ai].append((read, al))
    else:
        for read in sample.reads:
            by[read.classification].append((read, 0))
    alleles = [_Allele(sample.allele_seqs[i], sample.motif_counts[i], sample.allele_lengths[i],
                       sample.genotype_indices[i], i, by[i]) for i in range(n_all)]
    return {"alleles": alleles, "hp": hp, "karyotype": sample.karyotype}


def _ploidy(kary, chrom):
    if kary == "XX":
        return 0 if chrom in ("Y", "chrY") else 2
    return 1 if chrom in ("X", "chrX", "Y", "chrY") else 2


def _dropout(aset, locus):
    hp1, hp2, un = aset["hp"]
    if hp1 + hp2 + un < 2:
        return "FD"
    if un < 2 and _ploidy(aset["karyotype"], locus.chrom) != 1 and (hp1 < 2 or hp2 < 2):
        return "HD"
    return "N"


def _fld(a, qm):
    return a.allele_length if qm[0] == "AL" else a.motif_count


def _fdiv(a, b):
    if b == 0:
        return math.nan if a == 0 else math.copysign(math.inf, a)
    return a / b


def _trio_skip(f, m, c, qm):
    if qm[1] is None:
        fv, mv = [_fld(a, qm) for a in f["alleles"]], [_fld(a, qm) for a in m["alleles"]]
        cv = sorted(_fld(a, qm) for a in c["alleles"])
        if len(cv) == 1:
            return cv[0] in fv or cv[0] in mv
        return any(sorted([x, y]) == cv for x in fv for y in mv)
    tol = qm[1]
    fv, mv = [int(_fld(a, qm)) for a in f["alleles"]], [int(_fld(a, qm)) for a in m["alleles"]]
    cv = sorted(int(_fld(a, qm)) for a in c["alleles"])
    if len(cv) == 1:
        bf = min(float(abs(x - cv[0])) for x in fv)
        bm = min(float(abs(x - cv[0])) for x in mv)
        rf, rm = _fdiv(bf, float(max(fv))), _fdiv(bm, float(max(mv)))
        rel = rm if rf != rf else (rf if rm != rm else min(rf, rm))
    else:
        best, bsum = None, math.inf
        for x in fv:
            for y in mv:
                pair = (x, y) if x < y else (y, x)
                dsum = float(abs(pair[0] - cv[0])) + float(abs(pair[1] - cv[1]))
                if dsum < bsum:
                    bsum, best = dsum, pair
        if best is None:
            rel = math.inf
        else:
            d0 = 0.0 if best[0] == 0 and cv[0] == 0 else _fdiv(float(abs(best[0] - cv[0])), float(max(best[0], cv[0])))
            d1 = 0.0 if best[1] == 0 and cv[1] == 0 else _fdiv(float(abs(best[1] - cv[1])), float(max(best[1], cv[1])))
            rel = d1 if d0 != d0 else (d0 if d1 != d1 else max(d0, d1))
    return rel <= tol


def _duo_skip(a, b, qm):
    if qm[1] is None:
        return {_fld(x, qm) for x in a["alleles"]} == {_fld(x, qm) for x in b["alleles"]}
    av = sorted(int(_fld(x, qm)) for x in a["alleles"])
    bv = sorted(int(_fld(x, qm)) for x in b["alleles"])
    if len(av) != len(bv):
        return False
    return all(_fdiv(float(abs(x - y)), float(max(x, y))) <= qm[1] for x, y in zip(av, bv))


def _align_alleleset(aset, target, clip, aligner):
    out = []
    for allele in aset["alleles"]:
        scores = []
        for read, _ in allele.read_aligns:
            if aligner.align_end_to_end(read.bases, target) == 0:
                scores.append(aligner.cigar_score_clipped(clip))
        out.append(scores)
    return out


def _assess(mother, father, child, params, aligner, duo):
    """assess_denovo: the alignment loops here, the reductions in tdo_reduce.c on the collected scores."""
    loc = tdo.TrioLocus()
    loc.n_child, loc.n_mother = len(child["alleles"]), len(mother["alleles"])
    loc.n_father = 0 if duo else len(father["alleles"])
    flat, own = [], {}
    for c, ca in enumerate(child["alleles"]):
        ms = _align_alleleset(mother, ca.seq, params.clip_len, aligner)
        fs = [] if duo else _align_alleleset(father, ca.seq, params.clip_len, aligner)
        cs = []
        for read, _ in ca.read_aligns:
            if aligner.align_end_to_end(read.bases, ca.seq) == 0:
                cs.append(aligner.cigar_score_clipped(params.clip_len))
        lists = [ms[0], ms[1] if len(ms) > 1 else [], fs[0] if fs else [], fs[1] if len(fs) > 1 else [], cs]
        for j, lst in enumerate(lists):
            loc.list_off[c][j] = len(flat)
            flat.extend(lst)
        loc.list_off[c][5] = len(flat)
        own[c] = (loc.list_off[c][4], len(cs))
    for a in range(2):
        loc.child_len[a] = len(child["alleles"][a].seq) if a < loc.n_child else 0
        loc.mother_len[a] = len(mother["alleles"][a].seq) if a < loc.n_mother else 0
        loc.father_len[a] = len(father["alleles"][a].seq) if (not duo and a < loc.n_father) else 0
    scores = np.array(flat, dtype=np.int32)
    out, mask = tdo.trio_assess(loc, scores, params.parent_quantile, duo=duo)
    return out, mask, own


def _cols(row, names, aset, locus, prob=None):
    if aset is None:
        return
    row[names[0]] = _dropout(aset, locus)
    row[names[1]] = ",".join(str(len(a.read_aligns)) for a in aset["alleles"])
    row[names[2]] = ",".join(a.motif_count for a in aset["alleles"])
    row[names[3]] = ",".join(a.allele_length for a in aset["alleles"])
    if prob:
        row[prob] = math.pow(0.5, float(sum(len(a.read_aligns) for a in aset["alleles"])))


def process_alleles_trio(locus, father_in, mother_in, child_in, params, aligner):
    row = {"chrom": locus.chrom, "start": locus.start, "end": locus.end, "motifs": ",".join(locus.motifs),
           "trid": locus.trid, "genotype": 0, "denovo_coverage": 0, "allele_coverage": 0, "allele_ratio": 0.0,
           "child_coverage": 0, "child_ratio": 0.0, "mean_diff_father": 0.0, "mean_diff_mother": 0.0,
           "father_dropout_prob": 0.0, "mother_dropout_prob": 0.0, "allele_origin": ".", "denovo_status": ".",
           "per_allele_reads_father": ".", "per_allele_reads_mother": ".", "per_allele_reads_child": ".",
           "father_dropout": ".", "mother_dropout": ".", "child_dropout": ".", "index": 0, "father_MC": ".",
           "mother_MC": ".", "child_MC": ".", "father_AL": ".", "mother_AL": ".", "child_AL": ".",
           "father_overlap_coverage": ".", "mother_overlap_coverage": ".", "read_ids": None}
    f = load_alleles(father_in, params, aligner)
    m = load_alleles(mother_in, params, aligner)
    c = load_alleles(child_in, params, aligner)
    _cols(row, ("father_dropout", "per_allele_reads_father", "father_MC", "father_AL"), f, locus, "father_dropout_prob")
    _cols(row, ("mother_dropout", "per_allele_reads_mother", "mother_MC", "mother_AL"), m, locus, "mother_dropout_prob")
    _cols(row, ("child_dropout", "per_allele_reads_child", "child_MC", "child_AL"), c, locus)
    if f is None or m is None or c is None:
        return [row]
    if params.quick_mode is not None and _trio_skip(f, m, c, params.quick_mode):
        return [row]
    out, mask, own = _assess(m, f, c, params, aligner, duo=False)
    child_cov = sum(len(a.read_aligns) for a in c["alleles"])
    rows = []
    for ci, ca in enumerate(c["alleles"]):
        o = out[ci]
        assert not o.error, "reference panics: all parental score lists empty"
        r = dict(row)
        o4, n_own = own[ci]
        ids = [ca.read_aligns[k][0].name for k in range(n_own) if mask[o4 + k]]
        r.update(genotype=ca.genotype, denovo_coverage=o.denovo_coverage, allele_coverage=len(ca.read_aligns),
                 allele_ratio=0.0 if not ca.read_aligns else o.denovo_coverage / len(ca.read_aligns),
                 child_coverage=child_cov, child_ratio=_fdiv(float(o.denovo_coverage), float(child_cov)),
                 mean_diff_father=o.mean_diff_father, mean_diff_mother=o.mean_diff_mother,
                 allele_origin=ORIGIN[o.allele_origin], denovo_status=STATUS[o.denovo_status], index=ca.index,
                 father_overlap_coverage=",".join(str(o.father_overlap[a]) for a in range(len(f["alleles"]))),
                 mother_overlap_coverage=",".join(str(o.mother_overlap[a]) for a in range(len(m["alleles"]))),
                 read_ids=ids or None)
        rows.append(r)
    return rows or [row]


def process_alleles_duo(locus, a_in, b_in, params, aligner):
    row = {"chrom": locus.chrom, "start": locus.start, "end": locus.end, "motifs": ",".join(locus.motifs),
           "trid": locus.trid, "genotype": 0, "denovo_coverage": 0, "allele_coverage": 0, "allele_ratio": 0.0,
           "a_coverage": 0, "a_ratio": 0.0, "mean_diff_b": 0.0, "denovo_status": ".", "per_allele_reads_a": ".",
           "per_allele_reads_b": ".", "a_dropout": ".", "b_dropout": ".", "index": 0, "a_MC": ".", "b_MC": ".",
           "a_AL": ".", "b_AL": ".", "b_overlap_coverage": ".", "read_ids": None}
    a = load_alleles(a_in, params, aligner)
    b = load_alleles(b_in, params, aligner)
    _cols(row, ("a_dropout", "per_allele_reads_a", "a_MC", "a_AL"), a, locus)
    _cols(row, ("b_dropout", "per_allele_reads_b", "b_MC", "b_AL"), b, locus)
    if a is None or b is None:
        return [row]
    if params.quick_mode is not None and _duo_skip(a, b, params.quick_mode):
        return [row]
    out, mask, own = _assess(b, None, a, params, aligner, duo=True)
    a_cov = sum(len(x.read_aligns) for x in a["alleles"])
    rows = []
    for ci, ca in enumerate(a["alleles"]):
        o = out[ci]
        assert not o.error
        r = dict(row)
        o4, n_own = own[ci]
        ids = [ca.read_aligns[k][0].name for k in range(n_own) if mask[o4 + k]]
        r.update(genotype=ca.genotype, denovo_coverage=o.denovo_coverage, allele_coverage=len(ca.read_aligns),
                 allele_ratio=0.0 if not ca.read_aligns else o.denovo_coverage / len(ca.read_aligns),
                 a_coverage=a_cov, a_ratio=_fdiv(float(o.denovo_coverage), float(a_cov)), mean_diff_b=o.mean_diff_mother,
                 denovo_status=STATUS[o.denovo_status], index=ca.index,
                 b_overlap_coverage=",".join(str(o.mother_overlap[x]) for x in range(len(b["alleles"]))),
                 read_ids=ids or None)
        rows.append(r)
    return rows or [row]


FLOAT_COLS = {"allele_ratio", "child_ratio", "mean_diff_father", "mean_diff_mother", "father_dropout_prob",
              "mother_dropout_prob", "a_ratio", "mean_diff_b"}


def write_outputs(rows_per_locus, columns):
    """-> (tsv text, read-id text) as the writer thread produces them with -@1 (BED order)."""
    tsv, ids = ["\t".join(columns)], []
    for rows in rows_per_locus:
        for i, row in enumerate(rows):
            if row.get("read_ids"):
                ids.append(">TRID=%s\tALLELE=%d\tN=%d\n%s\n" % (row["trid"], i, len(row["read_ids"]),
                                                               "\n".join(row["read_ids"])))
            tsv.append("\t".join(fmt_prec(row[c]) if c in FLOAT_COLS else str(row[c]) for c in columns))
    return "\n".join(tsv) + "\n", "".join(ids)
