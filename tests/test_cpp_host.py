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


def test_cpp_packing_at_the_source_equals_host_pack(exe, tmp_path):
    """tdbh::PackedBases (from ASCII, and straight from BAM 4-bit SEQ bytes) + detail::Batch (shift-append of
    packed reads, ASCII targets packed in place, shards merged) produce the stream, lengths and exception list of
    tdb_packed_seqs -- word for word what tdb_host_pack makes of the same ASCII blob.  No device needed."""
    import ctypes as C
    import numpy as np
    import trgt_denovo_b200 as tdb
    rng = random.Random(7)
    code = "=ACMGRSVTWYHKDBN"
    seqs, lines = [], []
    for i in range(400):
        n = rng.choice([0, 1, 2, 15, 16, 17, 31, 33, rng.randint(1, 400)])
        alpha = "ACGT" if rng.random() < 0.9 else "ACGTNRYacgt"
        kind = rng.choice("APN")
        if kind == "N":
            alpha = "ACGT" if rng.random() < 0.9 else "ACGTNRYKM="  # what BAM can hold
        s = "".join(rng.choice(alpha) for _ in range(n))
        seqs.append(s)
        if kind == "N":
            nib = [code.index(c) for c in s] + ([0] if n % 2 else [])
            hexs = "".join(f"{(nib[k] << 4) | nib[k + 1]:02x}" for k in range(0, len(nib), 2))
            lines.append(f"N\t{hexs}\t{n}")
        else:
            lines.append(f"{kind}\t{s}")
        if i % 97 == 96:
            lines.append("B")
    (tmp_path / "in.txt").write_text("\n".join(lines) + "\n")
    subprocess.run([exe, "pack", str(tmp_path / "in.txt"), str(tmp_path / "out.txt")], check=True)
    out = (tmp_path / "out.txt").read_text().split("\n")
    blob = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    ln = np.array([len(s) for s in seqs], dtype=np.uint32)
    pk = tdb.pack_sequences(blob, None, ln)
    assert int(out[0]) == blob.size
    words = np.array([int(x, 16) for x in out[1].split()], dtype=np.uint32)
    nw = (blob.size + 15) // 16
    assert len(words) == len(pk.packed) and (words[:nw] == pk.packed[:nw]).all() and (words[nw:] == 0).all()
    assert [int(x) for x in out[2].split()] == list(ln)
    assert [int(x) for x in out[3].split()] == list(pk.exc_seq) and len(pk.exc_seq) > 5
    assert out[4].encode() == pk.exc_bytes.tobytes()


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


def _multi_gpu_case(exe, tmp_path, mode, n_gpus):
    """tdbh::MultiGpu: one ctx + host thread per GPU, loci in balanced contiguous ranges, rows merged in locus order
    -> the same TSV and read-id text as one GPU (and as the CPU restatement)."""
    from importlib import import_module
    import trgt_denovo_b200  # noqa: F401
    from oracle import tdo, tdo_rows
    from tests.util import make_family_loci
    R = import_module("trgt_denovo_b200.rows")
    params = R.Params(partition_by_alignment=(mode == "duo"))
    rng = random.Random(900 + n_gpus)
    loci, f, m, c = make_family_loci(rng, 150, duo=(mode == "duo"), denovo_rate=0.3)
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
    one_tsv, one_ids = str(tmp_path / "one.tsv"), str(tmp_path / "one_ids.txt")
    subprocess.run([exe, mode, fam, one_tsv, one_ids], check=True)
    out_tsv, out_ids = str(tmp_path / "multi.tsv"), str(tmp_path / "multi_ids.txt")
    r = subprocess.run([exe, mode, fam, out_tsv, out_ids, str(n_gpus)], check=True, capture_output=True, text=True)
    bounds = [int(x) for x in r.stdout.split()[1:]]
    assert len(bounds) == n_gpus + 1 and bounds[0] == 0 and bounds[-1] == len(loci)
    assert all(b > a for a, b in zip(bounds, bounds[1:])), bounds  # every GPU got a range
    assert open(out_tsv).read() == open(one_tsv).read() == wtsv
    assert open(out_ids).read() == open(one_ids).read() == wids and len(wids) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["trio", "duo"])
def test_cpp_multi_gpu_driver_shards_and_merges(exe, tmp_path, mode):
    """Three contexts (sharing the visible devices round-robin when there are fewer): the sharding and the merge in
    locus order, on any box."""
    _multi_gpu_case(exe, tmp_path, mode, 3)


@pytest.mark.gpu
def test_cpp_multi_gpu_two_devices(exe, tmp_path):
    """Two real GPUs (gpurun --gpus 2): the merged TSV equals the one-GPU TSV."""
    import trgt_denovo_b200 as tdb
    if tdb.device_count() < 2:
        pytest.skip("needs two GPUs")
    _multi_gpu_case(exe, tmp_path, "trio", 2)
