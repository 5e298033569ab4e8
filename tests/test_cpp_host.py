"""The C++ host side (include/trgt_denovo_b200.hpp): compiled-language mirror of the reference interface above
the C ABI.  CPU: it compiles, links against the library and refuses to run without a device.  GPU: WFAligner
reproduces the reference's known-answer vectors, and trio / duo process_alleles produce TSV + read-id text
identical to the loop-for-loop CPU restatement (oracle/tdo_rows.py)."""
import json
import os
import random
import subprocess

import pytest

import __graft_entry__ as entry

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    entry.build()
    out = str(tmp_path_factory.mktemp("cpp") / "host_tsv")
    libdir = os.path.dirname(entry.LIB)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(entry.ROOT, "include"),
                           os.path.join(HERE, "cpp", "host_tsv.cpp"), "-o", out, "-L", libdir, "-ltrgt_denovo_b200",
                           "-pthread", "-Wl,-rpath," + libdir])
    return out


def test_cpp_host_builds_and_fails_loudly_without_device(exe):
    import trgt_denovo_b200 as tdb
    if tdb.device_count() > 0:
        pytest.skip("GPU present")
    r = subprocess.run([exe, "align", "ACGT", "ACGT", "0"], capture_output=True, text=True)
    assert r.returncode == 42 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_wfaligner_known_answers(exe, golden_dir):
    for x in json.load(open(os.path.join(golden_dir, "aligner_vectors.json"))):
        if tuple(x["penalties"]) != (8, 4, 2, 24, 1) or x["heuristic"] != "default":
            continue  # the C++ mirror constructs the production configuration only
        clips = list((x.get("clipped") or {"0": x["score"]}).items())
        for clip, exp in clips:
            r = subprocess.run([exe, "align", x["pattern"], x["text"], str(clip)], capture_output=True, text=True, check=True)
            st, score, cigar, clipped = r.stdout.split()[:4]
            assert int(st) == 0 and int(score) == x["score"], x["id"]
            if "cigar" in x:
                assert cigar == x["cigar"]
            assert int(clipped) == exp
            if str(clip) in (x.get("clipped_cigar") or {}):
                assert r.stdout.split()[4] == x["clipped_cigar"][str(clip)]


def _write_family(path, params, loci, groups, roles):
    qm = params.quick_mode
    with open(path, "w") as o:
        o.write("\t".join(["P", str(params.clip_len), repr(params.parent_quantile), "1" if params.partition_by_alignment else "0",
                           qm[0] if qm else "-", "-" if (not qm or qm[1] is None) else repr(qm[1])]) + "\n")
        for i, l in enumerate(loci):
            o.write("\t".join(["L", l.chrom, str(l.start), str(l.end), l.trid, ",".join(l.motifs)]) + "\n")
            for role, g in zip(roles, groups):
                s = g[i]
                o.write("\t".join(["S", role, "0" if s is None else "1", "XX" if s is None else s.karyotype]) + "\n")
                if s is None:
                    continue
                for k in range(len(s.allele_seqs)):
                    o.write("\t".join(["A", s.allele_seqs[k].decode(), s.motif_counts[k], s.allele_lengths[k],
                                       str(s.genotype_indices[k])]) + "\n")
                for r in s.reads:
                    o.write("\t".join(["R", r.name, r.bases.decode(), str(-1 if r.classification is None else r.classification),
                                       str(0 if r.haplotype is None else r.haplotype)]) + "\n")


CASES = [dict(), dict(partition_by_alignment=True, parent_quantile=0.9), dict(quick_mode=("AL", None)),
         dict(quick_mode=("MC", 0.1), clip_len=0)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("mode", ["trio", "duo"])
def test_cpp_process_alleles_tsv_identical(exe, tmp_path, mode, case):
    from importlib import import_module
    import trgt_denovo_b200  # noqa: F401
    from oracle import tdo, tdo_rows
    from tests.util import make_family_loci
    R = import_module("trgt_denovo_b200.rows")
    params = R.Params(**CASES[case])
    rng = random.Random(300 + case + (50 if mode == "duo" else 0))
    loci, f, m, c = make_family_loci(rng, 80, duo=(mode == "duo"))
    fam = str(tmp_path / "family.txt")
    aligner = tdo.Aligner()
    if mode == "trio":
        _write_family(fam, params, loci, (f, m, c), "FMC")
        want = [tdo_rows.process_alleles_trio(loci[i], f[i], m[i], c[i], params, aligner) for i in range(len(loci))]
        wtsv, wids = tdo_rows.write_outputs(want, R.TRIO_COLUMNS)
    else:
        _write_family(fam, params, loci, (c, m), "AB")
        want = [tdo_rows.process_alleles_duo(loci[i], c[i], m[i], params, aligner) for i in range(len(loci))]
        wtsv, wids = tdo_rows.write_outputs(want, R.DUO_COLUMNS)
    out_tsv, out_ids = str(tmp_path / "out.tsv"), str(tmp_path / "ids.txt")
    subprocess.run([exe, mode, fam, out_tsv, out_ids], check=True)
    assert open(out_tsv).read() == wtsv
    assert open(out_ids).read() == wids
