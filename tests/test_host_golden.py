"""The reference's own unit tests for the host code either side of the hot path, as data
(tests/golden/host_vectors.json, extracted by tests/golden/make_golden.py from
/root/reference/src/trio/allele.rs:351-416, src/duo/allele.rs:248-364, src/allele.rs:523-580, src/locus.rs:262-277),
asserted against BOTH host mirrors: trgt_denovo_b200.rows (Python) and include/trgt_denovo_b200*.hpp (C++).
CPU only: none of these functions touches the device."""
import json
import os
import subprocess

import pytest

import __graft_entry__ as entry

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def vectors(golden_dir):
    return json.load(open(os.path.join(golden_dir, "host_vectors.json")))


@pytest.fixture(scope="module")
def R():
    from importlib import import_module
    import trgt_denovo_b200  # noqa: F401
    return import_module("trgt_denovo_b200.rows")


def _aset(R, pairs):
    return R.AlleleSet(alleles=[R.Allele(seq=b"", motif_count=mc, allele_length=al, genotype=0, index=i)
                                for i, (mc, al) in enumerate(pairs)], hp_counts=(0, 0, 0))


def test_python_mirror_reproduces_reference_host_tests(vectors, R):
    for c in vectors["trio_quick_mode"]:
        qm = tuple(c["quick_mode"])
        f, m, ch = (_aset(R, c[k]) for k in ("father", "mother", "child"))
        fn = R.trio_check_field_equivalence if c["fn"] == "check_field_equivalence" else R.trio_check_field_similarity
        assert fn(f, m, ch, qm) == c["expect"], c
    for c in vectors["duo_quick_mode"]:
        qm = tuple(c["quick_mode"])
        fn = R.duo_check_field_equivalence if c["fn"] == "check_allele_field_equivalence" else R.duo_check_field_similarity
        assert fn(_aset(R, c["a"]), _aset(R, c["b"]), qm) == c["expect"], c
    for c in vectors["build_allele_seqs"]:
        alleles, idx = R.build_allele_seqs([s.encode() for s in c["vcf_allele_seqs"]], c["genotype_indices"],
                                           c["left_flank"].encode(), c["right_flank"].encode())
        assert [a.decode() for a in alleles] == c["expect_alleles"] and idx == c["expect_indices"], c
    for c in vectors["decode_info_field"]:
        if c["expect"] is None:
            with pytest.raises(ValueError):
                R.decode_info_field(c["encoding"])
        else:
            assert list(R.decode_info_field(c["encoding"])) == c["expect"]


def test_cpp_mirror_reproduces_reference_host_tests(vectors, tmp_path):
    exe = str(tmp_path / "golden_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(entry.ROOT, "include"),
                           os.path.join(HERE, "cpp", "golden_check.cpp"), "-o", exe, "-lz", "-pthread"])
    enc = lambda pairs: ",".join(f"{mc}:{al}" for mc, al in pairs)  # noqa: E731
    tol = lambda qm: "-" if qm[1] is None else repr(qm[1])  # noqa: E731
    lines, want = [], []
    for c in vectors["trio_quick_mode"]:
        lines.append("\t".join(["T", "equiv" if c["fn"] == "check_field_equivalence" else "sim", c["quick_mode"][0],
                                tol(c["quick_mode"]), enc(c["father"]), enc(c["mother"]), enc(c["child"])]))
        want.append("1" if c["expect"] else "0")
    for c in vectors["duo_quick_mode"]:
        lines.append("\t".join(["D", "equiv" if c["fn"] == "check_allele_field_equivalence" else "sim", c["quick_mode"][0],
                                tol(c["quick_mode"]), enc(c["a"]), enc(c["b"])]))
        want.append("1" if c["expect"] else "0")
    for c in vectors["build_allele_seqs"]:
        lines.append("\t".join(["B", ",".join(c["vcf_allele_seqs"]), ",".join(map(str, c["genotype_indices"])),
                                c["left_flank"], c["right_flank"]]))
        want.append(",".join(c["expect_alleles"]) + "\t" + ",".join(map(str, c["expect_indices"])))
    for c in vectors["decode_info_field"]:
        lines.append("I\t" + c["encoding"])
        want.append("ERR" if c["expect"] is None else "\t".join(c["expect"]))
    r = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True)
    got = r.stdout.rstrip("\n").split("\n")
    assert got == want
