#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors for the alignment hot path into JSON fixtures.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Reads the inline unit tests of
    /root/reference/src/aligner.rs      (tests at :764-1252)
    /root/reference/src/trio/denovo.rs  (tests at :364-612)
    /root/reference/src/math.rs         (tests at :157-201)
    /root/reference/src/trio/allele.rs  (quick-mode tests at :351-416)
    /root/reference/src/duo/allele.rs   (quick-mode tests at :248-364)
    /root/reference/src/allele.rs       (build_allele_seqs tests at :523-580)
    /root/reference/src/locus.rs        (decode_info_field tests at :262-277)
and writes tests/golden/aligner_vectors.json, tests/golden/reduction_vectors.json and
tests/golden/host_vectors.json.
Only test *data* (input byte strings, asserted outputs) is extracted; no code is copied.
"""
import json
import os
import re
import sys

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def fn_body(src: str, name: str) -> str:
    i = src.index(f"fn {name}(")
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        c = src[k]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return src[j : k + 1]
        k += 1


def bytestrings(body: str):
    return re.findall(r'b"([A-Za-z]*)"', body)


def ints_after(body: str, what: str):
    return [int(x) for x in re.findall(re.escape(what) + r"\s*,\s*(-?\d+)\)", body)]


def main():
    a = open(os.path.join(REF, "aligner.rs")).read()
    pattern = re.search(r'const PATTERN: &\[u8\] = b"([A-Z]+)"', a).group(1)
    text = re.search(r'const TEXT: &\[u8\] = b"([A-Z]+)"', a).group(1)
    vec = []

    # B4 gap-affine (x6,o4,e2) MemoryLow: score -40
    b = fn_body(a, "test_aligner_gap_affine")
    vec.append(dict(id="B4", cite="src/aligner.rs:820-838", pattern=pattern, text=text,
                    penalties=[6, 4, 2, -1, 0], heuristic="default",
                    score=ints_after(b, "aligner.score()")[0],
                    cigar=re.search(r'aligner\.cigar\(None\), "([0-9MXID]+)"', b).group(1)))
    # B5 2-piece (6,2,2,4,1): -34
    b = fn_body(a, "test_aligner_gap_affine_2pieces")
    vec.append(dict(id="B5", cite="src/aligner.rs:857-877", pattern=pattern, text=text,
                    penalties=[6, 2, 2, 4, 1], heuristic="default",
                    score=ints_after(b, "aligner.score()")[0],
                    cigar=re.search(r'aligner\.cigar\(None\), "([0-9MXID]+)"', b).group(1)))
    # B7 memory modes, production penalties 8,4,2,24,1: -48, clipped(0) -48
    b = fn_body(a, "test_memory_modes")
    vec.append(dict(id="B7", cite="src/aligner.rs:1056-1100", pattern=pattern, text=text,
                    penalties=[8, 4, 2, 24, 1], heuristic="default",
                    score=int(re.search(r"expected_score = (-?\d+)", b).group(1)),
                    cigar=re.search(r'expected_cigar = "([0-9MXID]+)"', b).group(1),
                    clipped={"0": int(re.search(r"expected_score = (-?\d+)", b).group(1))}))
    # B6 clipping score: flanks + ATTT x8 vs x10
    b = fn_body(a, "test_clipping_score")
    bs = bytestrings(b)
    text_lf, text_rf, pat_lf, pat_rf, motif = bs[0], bs[1], bs[2], bs[3], bs[4]
    vec.append(dict(id="B6", cite="src/aligner.rs:1014-1054",
                    pattern=pat_lf + motif * 8 + pat_rf, text=text_lf + motif * 10 + text_rf,
                    penalties=[8, 4, 2, 24, 1], heuristic="default",
                    score=ints_after(b, "aligner.score()")[0],
                    cigar=re.search(r'aligner\.cigar\(None\), "([0-9MXID]+)"', b).group(1),
                    clipped={"50": ints_after(b, "aligner.cigar_score_clipped(50)")[0]},
                    clipped_cigar={"50": re.search(r'aligner\.cigar\(Some\(50\)\), "([0-9MXID]+)"', b).group(1)}))
    # B9 invalid sequence: heuristic none => -881 (default heuristic result in MemoryLow is unpinned)
    b = fn_body(a, "test_invalid_sequence")
    bs = bytestrings(b)
    vec.append(dict(id="B9", cite="src/aligner.rs:1120-1142", pattern=bs[0], text=bs[1],
                    penalties=[8, 4, 2, 24, 1], heuristic="none",
                    score=[x for x in ints_after(b, "aligner.score()") if x > -2147483648][0]))
    json.dump(vec, open(os.path.join(HERE, "aligner_vectors.json"), "w"), indent=1)

    # ---- reductions ----
    red = {}
    d = open(os.path.join(REF, "trio/denovo.rs")).read()
    red["get_score_count_diff"] = dict(cite="src/trio/denovo.rs:405-413", top=-8.0,
                                       aligns=[-6, 3, 0], expect=[3, 7.0])
    red["trio_get_denovo_coverage"] = [
        dict(cite="src/trio/denovo.rs:415-444", mother=[[-24, -24, -24], [-24, -12, -24]],
             father=[[-8, -8, -12], [-8, -8, -20]], child=[-6, 0, 0], q=1.0,
             expect=dict(denovo_coverage=3, mean_diff_father=6.0, mean_diff_mother=10.0, mask=[1, 1, 1])),
    ]
    # sanity: the literals above are what the reference asserts
    assert "(3, 7.0)" in fn_body(d, "test_get_parental_count_diff")
    b = fn_body(d, "test_get_denovo_coverage")
    assert "vec![-24, -12, -24]" in b and "vec![-8, -8, -20]" in b and "6.0," in b and "10.0," in b
    b = fn_body(d, "test_get_child_count")
    cases = re.findall(r"\((-?\d+\.\d+), vec!\[(.*?)\]\], vec!\[([\d, ]+)\]\)", b)
    red["get_overlap_coverage"] = []
    for thr, lists, exp in cases:
        ls = [[int(x) for x in l.split(",")] for l in re.findall(r"vec!\[([-\d, ]+)", lists + "]")]
        red["get_overlap_coverage"].append(dict(cite="src/trio/denovo.rs:592-612", threshold=float(thr),
                                                lists=ls, expect=[int(x) for x in exp.split(",")]))
    b = fn_body(d, "test_parent_origin_matrix_1")
    vals = [float(x) for x in re.findall(r"matrix\[\[\d, \d\]\] = (-?\d+\.\d+);", b)]
    red["parent_origin_matrix"] = dict(cite="src/trio/denovo.rs:369-403", matrix_rowmajor_4x2=vals,
                                       expect_top_sum=-30.0, expect_origin=[2, 1])
    m = open(os.path.join(REF, "math.rs")).read()
    assert "Some(3.5)" in fn_body(m, "test_median_even") and "Some(3.0)" in fn_body(m, "test_median_odd")
    red["median"] = [dict(cite="src/math.rs:170-178", data=[3, 1, 4, 1, 5, 9, 2, 6, 5, 3], expect=3.5),
                     dict(cite="src/math.rs:175-178", data=[1, 3, 2, 5, 4], expect=3.0),
                     dict(cite="src/math.rs:165-168", data=[5], expect=5.0),
                     dict(cite="src/math.rs:160-163", data=[], expect=None)]
    json.dump(red, open(os.path.join(HERE, "reduction_vectors.json"), "w"), indent=1)
    print("wrote", len(vec), "aligner vectors and", len(red), "reduction groups")
    host = host_vectors()
    json.dump(host, open(os.path.join(HERE, "host_vectors.json"), "w"), indent=1)
    print("wrote", sum(len(v) for v in host.values()), "host-side vectors")


def _vec_strs(s):
    return re.findall(r'"([^"]*)"', s)


def _quick_mode(s):
    m = re.search(r"QuickMode::(AL|MC)\((None|Some\(([0-9.]+)\))\)", s)
    return [m.group(1), None if m.group(2) == "None" else float(m.group(3))]


def host_vectors():
    """Quick-mode pre-filter, build_allele_seqs and decode_info_field: the reference's own assertions as data."""
    out = {}
    # ---- src/trio/allele.rs:351-416 test_check_minimal_cost_inheritance ----
    b = fn_body(open(os.path.join(REF, "trio/allele.rs")).read(), "test_check_minimal_cost_inheritance")
    sets = {}
    for name, mc, al in re.findall(r"let (\w+) = create_allele_set\(vec!\[(.*?)\], vec!\[(.*?)\]\);", b, re.S):
        sets[name] = [list(x) for x in zip(_vec_strs(mc), _vec_strs(al))]
    cases = []
    for neg, fn, args in re.findall(r"assert!\(\s*(!?)(check_field_equivalence|check_field_similarity)\((.*?)\)\s*\);", b, re.S):
        parts = []
        for a in re.split(r",\s*\n", args.strip()):
            a = a.strip().rstrip(",")
            if a.startswith("&QuickMode"):
                continue
            m = re.match(r"&create_allele_set\(vec!\[(.*?)\], vec!\[(.*?)\]\)", a, re.S)
            parts.append([list(x) for x in zip(_vec_strs(m.group(1)), _vec_strs(m.group(2)))] if m else sets[a.lstrip("&")])
        assert len(parts) == 3, args
        cases.append(dict(cite="src/trio/allele.rs:351-416", fn=fn, quick_mode=_quick_mode(args), father=parts[0],
                          mother=parts[1], child=parts[2], expect=(neg == "")))
    assert len(cases) == 5 and [c["expect"] for c in cases] == [True, False, False, True, True]
    out["trio_quick_mode"] = cases
    # ---- src/duo/allele.rs:248-364 test_check_allele_field_equivalence ----
    b = fn_body(open(os.path.join(REF, "duo/allele.rs")).read(), "test_check_allele_field_equivalence")
    alleles = {}
    for name, mc, al in re.findall(r'let (allele\d+) = Allele \{.*?motif_count: "(\d+)"\.to_string\(\),\s*allele_length: "(\d+)"\.to_string\(\)', b, re.S):
        alleles[name] = [mc, al]
    sets = {}
    for name, members in re.findall(r"let (allele_set_\w) = AlleleSet \{\s*alleles: vec!\[(.*?)\],", b, re.S):
        sets[name] = [alleles[re.match(r"(allele\d+)", m.strip()).group(1)] for m in members.split(",") if m.strip()]
    cases = []
    for neg, fn, a, bb, qm in re.findall(r"assert!\(\s*(!?)(check_allele_field_equivalence|check_field_similarity)\(\s*&(\w+),\s*&(\w+),\s*&(QuickMode::.*?\))\s*\)\s*\);", b, re.S):
        cases.append(dict(cite="src/duo/allele.rs:248-364", fn=fn, quick_mode=_quick_mode(qm), a=sets[a], b=sets[bb],
                          expect=(neg == "")))
    assert len(cases) == 8 and [c["expect"] for c in cases] == [True, True, False, False, False, False, False, True]
    out["duo_quick_mode"] = cases
    # ---- src/allele.rs:523-580 build_allele_seqs ----
    a = open(os.path.join(REF, "allele.rs")).read()
    cases = []
    for test in ("test_build_allele_seqs_sorts_indices", "test_build_allele_seqs_with_homozygous",
                 "test_build_allele_seqs_with_three_alleles"):
        b = fn_body(a, test)
        vcf = re.findall(r'b"(\w+)"\.to_vec\(\)', re.search(r"let vcf_allele_seqs = vec!\[(.*?)\];", b).group(1))
        lf = re.search(r'let left_flank = b"(\w+)";', b).group(1)
        rf = re.search(r'let right_flank = b"(\w+)";', b).group(1)
        for gi in re.findall(r"build_allele_seqs\(&vcf_allele_seqs, vec!\[([\d, ]+)\], left_flank, right_flank\)", b):
            idx = sorted(int(x) for x in gi.split(","))
            cases.append(dict(cite="src/allele.rs:523-580 " + test, vcf_allele_seqs=vcf, genotype_indices=[int(x) for x in gi.split(",")],
                              left_flank=lf, right_flank=rf, expect_indices=idx,
                              expect_alleles=[lf + vcf[i] + rf for i in idx]))
    # the literal assertions of the reference, checked against what was just derived
    b = fn_body(a, "test_build_allele_seqs_sorts_indices")
    assert 'b"LFREFRF"' in b and 'b"LFALT1RF"' in b and "vec![0, 1]" in b
    assert cases[0]["expect_alleles"] == ["LFREFRF", "LFALT1RF"] and cases[0]["expect_indices"] == [0, 1]
    b = fn_body(a, "test_build_allele_seqs_with_homozygous")
    assert "vec![1, 1]" in b and 'b"LFALT1RF"' in b and cases[2]["expect_alleles"] == ["LFALT1RF", "LFALT1RF"]
    assert "vec![0, 2]" in fn_body(a, "test_build_allele_seqs_with_three_alleles") and cases[3]["expect_indices"] == [0, 2]
    out["build_allele_seqs"] = cases
    # ---- src/locus.rs:262-277 decode_info_field ----
    l = open(os.path.join(REF, "locus.rs")).read()
    b = fn_body(l, "can_decode_info_field")
    enc = re.search(r'decode_info_field\("([^"]+)"\)', b).group(1)
    name, value = re.findall(r'assert_eq!\("(\w+)", decoded\.\d\)', b)
    bad = re.search(r'decode_info_field\("([^"]+)"\)', fn_body(l, "panic_invalid_decode_info_field")).group(1)
    out["decode_info_field"] = [dict(cite="src/locus.rs:266-271", encoding=enc, expect=[name, value]),
                                dict(cite="src/locus.rs:273-277", encoding=bad, expect=None)]
    return out


if __name__ == "__main__":
    sys.exit(main())
