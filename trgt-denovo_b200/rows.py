"""Per-locus pipeline above the hot path: loaded alleles -> batched GPU assessment -> TSV rows.

Host-side mirror (the reference's host language, Rust, is not available here) of
  trio::allele::process_alleles      src/trio/allele.rs:185-349  (row schema :26-69, quick mode :71-183)
  duo::allele::process_alleles       src/duo/allele.rs:124-246   (row schema :89-122, quick mode :33-87)
  allele::assign_reads_by_alignment  src/allele.rs:208-247       (--partition-by-aln; tie flip stays on host)
  allele::assign_reads_by_classification src/allele.rs:249-259
  AlleleSet::get_naive_dropout       src/allele.rs:447-469, Karyotype::get_ploidy src/handles.rs:49-62
  math::get_per_allele_reads / get_dropout_prob   src/math.rs:18-49
  serialize_with_precision           src/allele.rs:32-50
  process_writer_thread              src/commands/shared.rs:161-195 (TSV + read-id side file)
Instead of one locus per worker thread with a thread-local aligner (src/commands/trio.rs:113-151), many
loci are gathered into one batch and every alignment + reduction runs on the GPU through the C ABI
(Context.align_batch / Context.assess_batch).  Input here is what `load_alleles` (src/allele.rs:146-191)
produces from VCF + BAM; reading those files (htslib) is out of scope for this path.
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import binding as B

ORIGIN_STR = ["M:1", "M:2", "F:1", "F:2", "?", "."]  # src/model.rs:128-138
STATUS_STR = ["X", "Y:+", "Y:-", "Y:=", "Y:?", "."]  # src/model.rs:57-83

TRIO_COLUMNS = ["chrom", "start", "end", "motifs", "trid", "genotype", "denovo_coverage", "allele_coverage",
                "allele_ratio", "child_coverage", "child_ratio", "mean_diff_father", "mean_diff_mother",
                "father_dropout_prob", "mother_dropout_prob", "allele_origin", "denovo_status",
                "per_allele_reads_father", "per_allele_reads_mother", "per_allele_reads_child", "father_dropout",
                "mother_dropout", "child_dropout", "index", "father_MC", "mother_MC", "child_MC", "father_AL",
                "mother_AL", "child_AL", "father_overlap_coverage", "mother_overlap_coverage"]
DUO_COLUMNS = ["chrom", "start", "end", "motifs", "trid", "genotype", "denovo_coverage", "allele_coverage",
               "allele_ratio", "a_coverage", "a_ratio", "mean_diff_b", "denovo_status", "per_allele_reads_a",
               "per_allele_reads_b", "a_dropout", "b_dropout", "index", "a_MC", "b_MC", "a_AL", "b_AL",
               "b_overlap_coverage"]


@dataclass
class Read:  # TrgtRead, src/read.rs:12-29 (fields this path uses)
    name: str
    bases: bytes
    classification: Optional[int] = None  # AL tag
    haplotype: Optional[int] = None  # HP tag


@dataclass
class Locus:  # src/locus.rs:18-29
    chrom: str
    start: int
    end: int
    motifs: List[str]
    trid: str


@dataclass
class SampleInput:
    """What get_vcf_data + get_reads return for one sample at one locus (src/allele.rs:274-426)."""
    allele_seqs: List[bytes]  # left_flank + allele + right_flank
    motif_counts: List[str]
    allele_lengths: List[str]
    genotype_indices: List[int]
    reads: List[Read]
    karyotype: str = "XX"


@dataclass
class Allele:  # src/allele.rs:54-67
    seq: bytes
    motif_count: str
    allele_length: str
    genotype: int
    index: int
    read_aligns: List[Tuple[Read, int]] = field(default_factory=list)


@dataclass
class AlleleSet:  # src/allele.rs:429-432
    alleles: List[Allele]
    hp_counts: Tuple[int, int, int]
    karyotype: str = "XX"


@dataclass
class Params:  # src/model.rs:37-42
    clip_len: int = 50
    parent_quantile: float = 1.0
    partition_by_alignment: bool = False
    quick_mode: Optional[Tuple[str, Optional[float]]] = None  # ("AL"|"MC", tolerance or None)


def serialize_with_precision(value) -> str:
    """src/allele.rs:32-50 (Rust `{:.1}` / `{:.3}` round the exact binary value, like C printf)."""
    v = float(value)
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    s = "%.1f" % v
    if not s.endswith(".0"):
        s = ("%.3f" % v).rstrip("0")
        if s.endswith("."):
            s += "0"
    return s


def get_ploidy(karyotype: str, chrom: str) -> int:  # src/handles.rs:40-62
    if karyotype == "XX":
        return 0 if chrom in ("Y", "chrY") else 2
    return 1 if chrom in ("X", "chrX", "Y", "chrY") else 2


def get_naive_dropout(aset: AlleleSet, locus: Locus) -> str:  # src/allele.rs:447-469
    hp1, hp2, unphased = aset.hp_counts
    total = hp1 + hp2 + unphased
    ploidy = get_ploidy(aset.karyotype, locus.chrom)
    if total < 2:
        return "FD"
    if unphased < 2 and ploidy != 1 and (hp1 < 2 or hp2 < 2):
        return "HD"
    return "N"


def calculate_hp_counts(reads):  # src/allele.rs:117-128
    c = [0, 0, 0]
    for r in reads:
        if r.haplotype == 1:
            c[0] += 1
        elif r.haplotype == 2:
            c[1] += 1
        elif r.haplotype is None:
            c[2] += 1
    return tuple(c)


def _field(allele: Allele, quick_mode):
    return allele.allele_length if quick_mode[0] == "AL" else allele.motif_count


def trio_check_field_equivalence(father, mother, child, quick_mode) -> bool:  # src/trio/allele.rs:71-99
    fv = [_field(a, quick_mode) for a in father.alleles]
    mv = [_field(a, quick_mode) for a in mother.alleles]
    cv = sorted(_field(a, quick_mode) for a in child.alleles)
    if len(cv) == 1:
        return cv[0] in fv or cv[0] in mv
    pairs = {tuple(sorted((f, m))) for f in fv for m in mv}
    return tuple(cv) in pairs


def trio_check_field_similarity(father, mother, child, quick_mode) -> bool:  # src/trio/allele.rs:101-183
    tol = quick_mode[1]
    fv = [int(_field(a, quick_mode)) for a in father.alleles]
    mv = [int(_field(a, quick_mode)) for a in mother.alleles]
    cv = sorted(int(_field(a, quick_mode)) for a in child.alleles)

    def div(a, b):  # Rust f64 division
        if b == 0:
            return math.nan if a == 0 else math.copysign(math.inf, a)
        return a / b

    if len(cv) == 1:
        c = cv[0]
        best_f = min([float(abs(f - c)) for f in fv], default=math.inf)
        best_m = min([float(abs(m - c)) for m in mv], default=math.inf)
        rel_f, rel_m = div(best_f, float(max(fv))), div(best_m, float(max(mv)))
        rel = _rust_fmin(rel_f, rel_m)
    else:
        best_pair, best_sum = None, math.inf
        for f in fv:
            for m in mv:
                pair = (f, m) if f < m else (m, f)
                dsum = float(abs(pair[0] - cv[0])) + float(abs(pair[1] - cv[1]))
                if dsum < best_sum:
                    best_sum, best_pair = dsum, pair
        if best_pair is None:
            rel = math.inf
        else:
            bf, bm = best_pair
            d0 = 0.0 if (bf == 0 and cv[0] == 0) else div(float(abs(bf - cv[0])), float(max(bf, cv[0])))
            d1 = 0.0 if (bm == 0 and cv[1] == 0) else div(float(abs(bm - cv[1])), float(max(bm, cv[1])))
            rel = _rust_fmax(d0, d1)
    return rel <= tol


def _rust_fmin(a, b):  # f64::min ignores a NaN operand
    if math.isnan(a):
        return b
    if math.isnan(b):
        return a
    return min(a, b)


def _rust_fmax(a, b):
    if math.isnan(a):
        return b
    if math.isnan(b):
        return a
    return max(a, b)


def duo_check_field_equivalence(a, b, quick_mode) -> bool:  # src/duo/allele.rs:33-49
    return {_field(x, quick_mode) for x in a.alleles} == {_field(x, quick_mode) for x in b.alleles}


def duo_check_field_similarity(a, b, quick_mode) -> bool:  # src/duo/allele.rs:51-87
    tol = quick_mode[1]
    av = sorted(int(_field(x, quick_mode)) for x in a.alleles)
    bv = sorted(int(_field(x, quick_mode)) for x in b.alleles)
    if len(av) != len(bv):
        return False
    for x, y in zip(av, bv):
        diff, mx = float(abs(x - y)), float(max(x, y))
        q = (math.nan if diff == 0 else math.inf) if mx == 0 else diff / mx
        if not q <= tol:
            return False
    return True


def build_allele_seqs(vcf_allele_seqs, genotype_indices, left_flank: bytes, right_flank: bytes):
    """src/allele.rs:406-426: sorted genotype indices, allele = left flank + VCF allele + right flank."""
    idx = sorted(genotype_indices)
    return [left_flank + vcf_allele_seqs[i] + right_flank for i in idx], idx


def decode_info_field(encoding: str):
    """src/locus.rs:249-260: 'name=value' of a catalog INFO entry; anything else is an error."""
    name, sep, value = encoding.partition("=")
    if not sep:
        raise ValueError(f"Invalid entry: {encoding}")
    return name, value


class _BatchBuilder:
    """Sequence blob + pair list in the reference's visiting order (src/denovo.rs:29-37,48-52)."""

    def __init__(self):
        self.chunks, self.lens, self.pp, self.pt = [], [], [], []

    def seq(self, s: bytes) -> int:
        self.chunks.append(s)
        self.lens.append(len(s))
        return len(self.lens) - 1

    def pair(self, p: int, t: int) -> int:
        self.pp.append(p)
        self.pt.append(t)
        return len(self.pp) - 1

    def arrays(self):
        blob = np.frombuffer(b"".join(self.chunks), dtype=np.uint8) if self.chunks else np.zeros(0, np.uint8)
        lens = np.array(self.lens, dtype=np.uint32)
        offs = None  # dense layout: sequences back to back, the library derives the offsets
        return blob, offs, lens, np.array(self.pp, dtype=np.uint32), np.array(self.pt, dtype=np.uint32)


def load_allele_sets(ctx, samples: List[Optional[SampleInput]], params: Params) -> List[Optional[AlleleSet]]:
    """load_alleles (src/allele.rs:146-191) for many (locus, sample) inputs at once.  With
    partition_by_alignment every read is aligned against every allele of its own sample in ONE batched
    GPU call; only the order-dependent tie flip (src/allele.rs:215,238-243) runs on the host."""
    out: List[Optional[AlleleSet]] = []
    scores = None
    pair_base = []
    if params.partition_by_alignment:
        bb = _BatchBuilder()
        for s in samples:
            pair_base.append(len(bb.pp))
            if s is None:
                continue
            ids = [bb.seq(a) for a in s.allele_seqs]
            for r in s.reads:
                ri = bb.seq(r.bases)
                for ai in ids:
                    bb.pair(ri, ai)
        if bb.pp:
            ctx.set_clip_len(params.clip_len)
            scores, status = ctx.align_batch(*bb.arrays())
            scores = np.where(status == 0, scores, np.iinfo(np.int32).min)  # non-completed alignments are skipped
    for si, s in enumerate(samples):
        if s is None:
            out.append(None)
            continue
        n_all = len(s.allele_seqs)
        by_allele = [[] for _ in range(n_all)]
        if params.partition_by_alignment:
            flip, j = 0, pair_base[si]
            for r in s.reads:
                sc = scores[j:j + n_all]
                j += n_all
                done = [(i, int(x)) for i, x in enumerate(sc) if x != np.iinfo(np.int32).min]
                if not done:
                    continue
                mx = max(x for _, x in done)
                best = [(i, x) for i, x in done if x == mx]
                if len(best) > 1:
                    flip = (flip + 1) % len(best)
                ai, al = best[flip % len(best)]
                by_allele[ai].append((r, al))
        else:
            for r in s.reads:
                if r.classification is None:
                    raise ValueError("read without AL tag under classification partitioning (src/allele.rs:255 panics)")
                by_allele[r.classification].append((r, 0))
        alleles = [Allele(seq=s.allele_seqs[i], motif_count=s.motif_counts[i], allele_length=s.allele_lengths[i],
                          genotype=s.genotype_indices[i], index=i, read_aligns=by_allele[i]) for i in range(n_all)]
        out.append(AlleleSet(alleles=alleles, hp_counts=calculate_hp_counts(s.reads), karyotype=s.karyotype))
    return out


def _join(xs):
    return ",".join(str(x) for x in xs)


def _sample_columns(row, prefix_cols, aset: Optional[AlleleSet], locus, with_prob=None):
    """Template columns filled from one loaded sample (src/trio/allele.rs:231-268)."""
    dropout, reads, mc, al = prefix_cols
    if aset is None:
        return
    row[dropout] = get_naive_dropout(aset, locus)
    row[reads] = _join(len(a.read_aligns) for a in aset.alleles)
    row[mc] = ",".join(a.motif_count for a in aset.alleles)
    row[al] = ",".join(a.allele_length for a in aset.alleles)
    if with_prob:
        row[with_prob] = 0.5 ** float(sum(len(a.read_aligns) for a in aset.alleles))  # math::get_dropout_prob


def _ratio(num, den):
    if den == 0:
        return math.nan if num == 0 else math.inf
    return num / den


def process_trio_batch(ctx, loci: List[Locus], father: List[Optional[SampleInput]],
                       mother: List[Optional[SampleInput]], child: List[Optional[SampleInput]], params: Params):
    """process_alleles for many loci.  -> list (per locus) of list of rows; a row is a dict over TRIO_COLUMNS
    plus "read_ids" (None or list of names)."""
    n = len(loci)
    sets = load_allele_sets(ctx, father + mother + child, params)
    fs, ms, cs = sets[:n], sets[n:2 * n], sets[2 * n:]
    templates, todo = [], []
    for i, locus in enumerate(loci):
        row = {"chrom": locus.chrom, "start": locus.start, "end": locus.end, "motifs": ",".join(locus.motifs),
               "trid": locus.trid, "genotype": 0, "denovo_coverage": 0, "allele_coverage": 0, "allele_ratio": 0.0,
               "child_coverage": 0, "child_ratio": 0.0, "mean_diff_father": 0.0, "mean_diff_mother": 0.0,
               "father_dropout_prob": 0.0, "mother_dropout_prob": 0.0, "allele_origin": ".", "denovo_status": ".",
               "per_allele_reads_father": ".", "per_allele_reads_mother": ".", "per_allele_reads_child": ".",
               "father_dropout": ".", "mother_dropout": ".", "child_dropout": ".", "index": 0, "father_MC": ".",
               "mother_MC": ".", "child_MC": ".", "father_AL": ".", "mother_AL": ".", "child_AL": ".",
               "father_overlap_coverage": ".", "mother_overlap_coverage": ".", "read_ids": None}
        _sample_columns(row, ("father_dropout", "per_allele_reads_father", "father_MC", "father_AL"), fs[i], locus,
                        "father_dropout_prob")
        _sample_columns(row, ("mother_dropout", "per_allele_reads_mother", "mother_MC", "mother_AL"), ms[i], locus,
                        "mother_dropout_prob")
        _sample_columns(row, ("child_dropout", "per_allele_reads_child", "child_MC", "child_AL"), cs[i], locus)
        templates.append(row)
        if fs[i] is None or ms[i] is None or cs[i] is None:
            continue  # missing genotyping: one template row (src/trio/allele.rs:270-286)
        qm = params.quick_mode
        if qm is not None:
            skip = (trio_check_field_equivalence(fs[i], ms[i], cs[i], qm) if qm[1] is None
                    else trio_check_field_similarity(fs[i], ms[i], cs[i], qm))
            if skip:
                continue
        todo.append(i)
    # ---- one batch for every locus that reaches assess_denovo (src/trio/denovo.rs:83-186) ----
    bb = _BatchBuilder()
    tl = (B.TrioLocus * max(1, len(todo)))()
    own_range = {}
    for li, i in enumerate(todo):
        loc = tl[li]
        loc.n_child, loc.n_mother, loc.n_father = len(cs[i].alleles), len(ms[i].alleles), len(fs[i].alleles)
        m_reads = [[bb.seq(r.bases) for r, _ in a.read_aligns] for a in ms[i].alleles]
        f_reads = [[bb.seq(r.bases) for r, _ in a.read_aligns] for a in fs[i].alleles]
        for a in range(2):
            loc.child_len[a] = len(cs[i].alleles[a].seq) if a < loc.n_child else 0
            loc.mother_len[a] = len(ms[i].alleles[a].seq) if a < loc.n_mother else 0
            loc.father_len[a] = len(fs[i].alleles[a].seq) if a < loc.n_father else 0
        for c, ca in enumerate(cs[i].alleles[:2]):
            tgt = bb.seq(ca.seq)
            lists = [m_reads[0], m_reads[1] if len(m_reads) > 1 else [], f_reads[0], f_reads[1] if len(f_reads) > 1 else []]
            for j, lst in enumerate(lists):
                loc.list_off[c][j] = len(bb.pp)
                for ri in lst:
                    bb.pair(ri, tgt)
            loc.list_off[c][4] = len(bb.pp)
            for r, _ in ca.read_aligns:
                bb.pair(bb.seq(r.bases), tgt)
            loc.list_off[c][5] = len(bb.pp)
            own_range[(i, c)] = (loc.list_off[c][4], loc.list_off[c][5])
    out = mask = None
    if todo:
        ctx.set_clip_len(params.clip_len)
        blob, off, ln, pp, pt = bb.arrays()
        out, _, mask = ctx.assess_batch(blob, off, ln, pp, pt, tl, q=params.parent_quantile, duo=False)
    result = [[t] for t in templates]
    for li, i in enumerate(todo):
        rows = []
        child_cov = sum(len(a.read_aligns) for a in cs[i].alleles)
        for c, ca in enumerate(cs[i].alleles[:2]):
            o = out[2 * li + c]
            if o.error:
                raise RuntimeError(f"TRID={loci[i].trid}: no parental alignment scores "
                                   "(the reference panics here, src/trio/denovo.rs:333-334)")
            row = dict(templates[i])
            a0, a1 = own_range[(i, c)]
            ids = [ca.read_aligns[k][0].name for k in range(a1 - a0) if mask[a0 + k]]
            row.update(genotype=ca.genotype, denovo_coverage=int(o.denovo_coverage), allele_coverage=len(ca.read_aligns),
                       allele_ratio=0.0 if len(ca.read_aligns) == 0 else o.denovo_coverage / len(ca.read_aligns),
                       child_coverage=child_cov, child_ratio=_ratio(o.denovo_coverage, child_cov),
                       mean_diff_father=np.float32(o.mean_diff_father), mean_diff_mother=np.float32(o.mean_diff_mother),
                       allele_origin=ORIGIN_STR[o.allele_origin], denovo_status=STATUS_STR[o.denovo_status],
                       index=ca.index,
                       father_overlap_coverage=_join(o.father_overlap[a] for a in range(len(fs[i].alleles))),
                       mother_overlap_coverage=_join(o.mother_overlap[a] for a in range(len(ms[i].alleles))),
                       read_ids=ids or None)
            rows.append(row)
        if rows:
            result[i] = rows
    return result


def process_duo_batch(ctx, loci: List[Locus], sample_a: List[Optional[SampleInput]],
                      sample_b: List[Optional[SampleInput]], params: Params):
    """duo process_alleles for many loci (src/duo/allele.rs:124-246)."""
    n = len(loci)
    sets = load_allele_sets(ctx, sample_a + sample_b, params)
    As, Bs = sets[:n], sets[n:]
    templates, todo = [], []
    for i, locus in enumerate(loci):
        row = {"chrom": locus.chrom, "start": locus.start, "end": locus.end, "motifs": ",".join(locus.motifs),
               "trid": locus.trid, "genotype": 0, "denovo_coverage": 0, "allele_coverage": 0, "allele_ratio": 0.0,
               "a_coverage": 0, "a_ratio": 0.0, "mean_diff_b": 0.0, "denovo_status": ".", "per_allele_reads_a": ".",
               "per_allele_reads_b": ".", "a_dropout": ".", "b_dropout": ".", "index": 0, "a_MC": ".", "b_MC": ".",
               "a_AL": ".", "b_AL": ".", "b_overlap_coverage": ".", "read_ids": None}
        _sample_columns(row, ("a_dropout", "per_allele_reads_a", "a_MC", "a_AL"), As[i], locus)
        _sample_columns(row, ("b_dropout", "per_allele_reads_b", "b_MC", "b_AL"), Bs[i], locus)
        templates.append(row)
        if As[i] is None or Bs[i] is None:
            continue
        qm = params.quick_mode
        if qm is not None:
            skip = (duo_check_field_equivalence(As[i], Bs[i], qm) if qm[1] is None
                    else duo_check_field_similarity(As[i], Bs[i], qm))
            if skip:
                continue
        todo.append(i)
    bb = _BatchBuilder()
    tl = (B.TrioLocus * max(1, len(todo)))()
    own_range = {}
    for li, i in enumerate(todo):
        loc = tl[li]
        loc.n_child, loc.n_mother, loc.n_father = len(As[i].alleles), len(Bs[i].alleles), 0
        b_reads = [[bb.seq(r.bases) for r, _ in a.read_aligns] for a in Bs[i].alleles]
        for a in range(2):
            loc.child_len[a] = len(As[i].alleles[a].seq) if a < loc.n_child else 0
            loc.mother_len[a] = len(Bs[i].alleles[a].seq) if a < loc.n_mother else 0
        for c, ca in enumerate(As[i].alleles[:2]):
            tgt = bb.seq(ca.seq)
            lists = [b_reads[0], b_reads[1] if len(b_reads) > 1 else [], [], []]
            for j, lst in enumerate(lists):
                loc.list_off[c][j] = len(bb.pp)
                for ri in lst:
                    bb.pair(ri, tgt)
            loc.list_off[c][4] = len(bb.pp)
            for r, _ in ca.read_aligns:
                bb.pair(bb.seq(r.bases), tgt)
            loc.list_off[c][5] = len(bb.pp)
            own_range[(i, c)] = (loc.list_off[c][4], loc.list_off[c][5])
    out = mask = None
    if todo:
        ctx.set_clip_len(params.clip_len)
        blob, off, ln, pp, pt = bb.arrays()
        out, _, mask = ctx.assess_batch(blob, off, ln, pp, pt, tl, q=params.parent_quantile, duo=True)
    result = [[t] for t in templates]
    for li, i in enumerate(todo):
        rows = []
        a_cov = sum(len(a.read_aligns) for a in As[i].alleles)
        for c, ca in enumerate(As[i].alleles[:2]):
            o = out[2 * li + c]
            if o.error:
                raise RuntimeError(f"TRID={loci[i].trid}: no sample-B alignment scores "
                                   "(the reference panics here, src/duo/denovo.rs:107)")
            row = dict(templates[i])
            a0, a1 = own_range[(i, c)]
            ids = [ca.read_aligns[k][0].name for k in range(a1 - a0) if mask[a0 + k]]
            row.update(genotype=ca.genotype, denovo_coverage=int(o.denovo_coverage), allele_coverage=len(ca.read_aligns),
                       allele_ratio=0.0 if len(ca.read_aligns) == 0 else o.denovo_coverage / len(ca.read_aligns),
                       a_coverage=a_cov, a_ratio=_ratio(o.denovo_coverage, a_cov),
                       mean_diff_b=np.float32(o.mean_diff_mother), denovo_status=STATUS_STR[o.denovo_status],
                       index=ca.index, b_overlap_coverage=_join(o.mother_overlap[a] for a in range(len(Bs[i].alleles))),
                       read_ids=ids or None)
            rows.append(row)
        if rows:
            result[i] = rows
    return result


_FLOAT_COLS = {"allele_ratio", "child_ratio", "mean_diff_father", "mean_diff_mother", "father_dropout_prob",
               "mother_dropout_prob", "a_ratio", "mean_diff_b"}


def format_tsv(rows_per_locus, columns, header=True) -> str:
    """csv::Writer with tab delimiter and serde field order (src/commands/shared.rs:117-119,185)."""
    lines = ["\t".join(columns)] if header else []
    for rows in rows_per_locus:
        for row in rows:
            lines.append("\t".join(serialize_with_precision(row[c]) if c in _FLOAT_COLS else str(row[c])
                                   for c in columns))
    return "\n".join(lines) + "\n"


def format_read_ids(rows_per_locus) -> str:
    """The `_denovo_reads.txt` side file (src/commands/shared.rs:169-182): ALLELE = row index within the locus."""
    out = []
    for rows in rows_per_locus:
        for i, row in enumerate(rows):
            ids = row.get("read_ids")
            if ids:
                out.append(">TRID=%s\tALLELE=%d\tN=%d\n%s\n" % (row["trid"], i, len(ids), "\n".join(ids)))
    return "".join(out)
