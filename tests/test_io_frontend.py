"""Real-input front end (include/trgt_denovo_b200_io.hpp): repeat catalog, indexed FASTA,
TRGT VCF and spanning BAM of a synthetic family are written by tests/io_fixtures.py (an independent writer of the
format specifications) and read back by the C++ readers.  CPU: what the readers hand to the batch builder equals
the in-memory description, including the reference's rules (flanks, padding base, genotype order, tag types,
karyotype, version check, per-line catalog errors).  GPU: the TSV / read-id text produced from the files equals
the loop-for-loop CPU restatement on the in-memory inputs."""
import os
import random
import subprocess

import pytest

import __graft_entry__ as entry
from tests import io_fixtures as IOF
from tests.test_cpp_host import _write_family

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    entry.build()
    out = str(tmp_path_factory.mktemp("cppio") / "files_tsv")
    libdir = os.path.dirname(entry.LIB)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(entry.ROOT, "include"),
                           os.path.join(HERE, "cpp", "files_tsv.cpp"), "-o", out, "-L", libdir, "-ltrgt_denovo_b200",
                           "-lz", "-pthread", "-Wl,-rpath," + libdir])
    return out


def _family_files(tmp, seed, n_loci=60, duo=False):
    """-> (ref.fa, catalog.bed, prefixes [f, m, c], loci, (f, m, c) in-memory inputs, n bad catalog lines)"""
    from tests.util import make_family_loci
    rng = random.Random(seed)
    loci, f, m, c = make_family_loci(rng, n_loci, duo=duo)
    for l in loci:  # room for the left flank of the first locus
        l.start += 1000
        l.end += 1000
    groups = {"f": f, "m": m, "c": c}
    kary = {"f": "XY", "m": "XX", "c": rng.choice(["XX", "XY"])}
    versions = {"f": "1.2.0", "m": "0.9.1", "c": "1.5.1"}  # major 0: alleles without the padding base
    for s, g in groups.items():
        for x in g:
            if x is not None:
                x.karyotype = kary[s]
    # ---- reference: random sequence in mixed case, the locus (flank + reference allele + flank) written over it
    chroms = ["chr1", "chr7", "chrX", "chrY", "chrUn"]
    length = 1000 * n_loci + 4000
    contig = {ch: [rng.choice("acgtACGTn") for _ in range(length)] for ch in chroms}
    flanks = []
    for i, l in enumerate(loci):
        src = next((g[i] for g in (f, m, c) if g[i] is not None), None)
        lf = src.allele_seqs[0][:50].decode() if src else "".join(rng.choice("ACGT") for _ in range(50))
        rf = src.allele_seqs[0][-50:].decode() if src else "".join(rng.choice("ACGT") for _ in range(50))
        ref_allele = (l.motifs[0] * ((l.end - l.start) // len(l.motifs[0]) + 1))[: l.end - l.start]
        piece = lf + ref_allele + rf
        piece = "".join(ch.lower() if rng.random() < 0.3 else ch for ch in piece)  # soft-masked reference
        contig[l.chrom][l.start - 50: l.end + 50] = piece
        flanks.append((lf, rf))
    ref = str(tmp / "ref.fa")
    IOF.write_fasta(ref, [(ch, "".join(contig[ch])) for ch in chroms], width=rng.choice([50, 60, 73]))
    # ---- catalog, with lines the reference rejects
    lines, bad = [], 0
    for i, l in enumerate(loci):
        if i % 17 == 3:
            lines.append(rng.choice(["chr1\t10\t20", "chrNope\t100\t120\tID=x;MOTIFS=A", "chr1\t500\t400\tID=y;MOTIFS=AC",
                                     "chr1\t100\t120\tMOTIFS=A", "chr1\t100\t120\tID=z", "chr1\t1e3\t2000\tID=w;MOTIFS=A",
                                     "chr1\t100\t120\tID=q;ID=r;MOTIFS=A"]))
            bad += 1
        sep = rng.choice(["\t", " "])
        lines.append(sep.join([l.chrom, str(l.start), str(l.end), f"ID={l.trid};MOTIFS={','.join(l.motifs)};STRUC=({l.motifs[0]})n"]))
    bed = str(tmp / "catalog.bed")
    IOF.write_text(bed, "\n".join(lines) + "\n", bgzf=seed % 2 == 0)
    # ---- per sample VCF + BAM
    prefixes = []
    for s in ("f", "m", "c"):
        pad = 1 if int(versions[s].split(".")[0]) > 0 else 0
        vcf = ["##fileformat=VCFv4.2", "##trgtVersion=0.0.1", f"##trgtVersion={versions[s]}", "##contig=<ID=chr1>",
               "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsample"]
        recs = []
        refs = [(ch, length) for ch in chroms]
        hdr = ("@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{ch}\tLN:{length}\n" for ch in chroms) +
               "@PG\tID:other\tPN:other\tVN:9.9\tCL:other --karyotype ZZ\n" +
               f"@PG\tID:trgt\tPN:trgt\tVN:{versions[s]}\tCL:trgt genotype --genome g.fa --karyotype {kary[s]} --reads r.bam\n")
        for i, l in enumerate(loci):
            x = groups[s][i]
            lf, rf = flanks[i]
            padb = "N" * pad
            if x is None:
                how = rng.randrange(3)
                if how == 0:  # a record whose genotype is missing
                    vcf.append("\t".join([l.chrom, str(l.start), ".", padb + "A", ".", ".", ".", f"TRID={l.trid};END={l.end}",
                                          "GT:AL:MC", ".:.:."]))
                elif how == 1:  # reads but no VCF record
                    recs.append(IOF.bam_record(f"orphan{i}", chroms.index(l.chrom), l.start, "ACGT", [("TR", "Z", l.trid)]))
                continue  # how == 2: nothing at all
            alleles = [a[50:len(a) - 50].decode() for a in x.allele_seqs]
            n = len(alleles)
            gt = "1" if n == 1 else rng.choice(["1/2", "2|1", "1|2"])
            vcf.append("\t".join([l.chrom, str(l.start), ".", padb + "ACAC", ",".join(padb + a for a in alleles), ".", ".",
                                  f"TRID={l.trid};END={l.end};MOTIFS={l.motifs[0]};STRUC=({l.motifs[0]})n",
                                  "GT:AL:ALLR:SD:MC:MS:AP:AM",
                                  ":".join([gt, ",".join(x.allele_lengths), "1-2,3-4", "10,11", ",".join(x.motif_counts),
                                            "0(0-1)", "1.0,1.0", ".,."])]))
            for k, r in enumerate(x.reads):
                tags = [("TR", "Z", l.trid), ("rq", "f", 0.999), ("XA", "A", "q"), ("XS", "s", -3)]
                tags.append(("AL", "C", r.classification) if k % 3 else ("AL", "i", r.classification))
                if r.haplotype is not None:
                    tags.append(("HP", "C" if k % 2 else "i", r.haplotype))
                tags += [("SO", "i", 12), ("EO", "i", -7), ("MO", "B", ("i", [1, -5, 9])), ("XB", "B", ("C", [1, 2, 3])),
                         ("XH", "H", "1AE3")]
                rng.shuffle(tags)
                recs.append(IOF.bam_record(r.name, chroms.index(l.chrom), l.start - 50, r.bases.decode(), tags))
            if i % 11 == 0:
                recs.append(IOF.bam_record(f"untagged{i}", chroms.index(l.chrom), l.start, "ACGTN", [("rq", "f", 0.5)]))
        prefix = str(tmp / f"sample_{s}")
        IOF.write_text(prefix + ".sorted.vcf.gz", "\n".join(vcf) + "\n", bgzf=True)
        IOF.write_bam(prefix + ".spanning.sorted.bam", hdr, refs, recs)
        prefixes.append(prefix)
    return ref, bed, prefixes, loci, (f, m, c), bad, kary, versions


def test_readers_reproduce_the_in_memory_inputs(exe, tmp_path):
    from importlib import import_module
    import trgt_denovo_b200  # noqa: F401
    R = import_module("trgt_denovo_b200.rows")
    for seed in (1, 2):
        d = tmp_path / f"s{seed}"
        d.mkdir()
        ref, bed, prefixes, loci, (f, m, c), bad, kary, versions = _family_files(d, seed)
        want = str(d / "want.txt")
        _write_family(want, R.Params(), loci, (f, m, c), "FMC")
        got = str(d / "got.txt")
        subprocess.run([exe, "dump", ref, bed, "50", got] + prefixes, check=True)
        lines = open(got).read().splitlines()
        body = [x for x in lines if x[0] not in "EV"]
        assert "\n".join(body) + "\n" == open(want).read()
        errs = [x for x in lines if x.startswith("E\tError at BED line")]
        assert len(errs) == bad
        vs = [x.split("\t") for x in lines if x.startswith("V")]
        for row, s in zip(vs, "fmc"):
            assert row[2] == versions[s] and row[3] == ("1" if versions[s][0] != "0" else "0") and row[4] == kary[s]
            assert int(row[5]) > 0  # reads without a TR tag were counted, not assigned
        missing = sum(x is None for g in (f, m, c) for x in g)
        why = [x for x in lines if x.startswith("E") and not x.startswith("E\tError at BED line")]
        assert len(why) == missing and all(("missing" in x or "Missing" in x or "No reads" in x) for x in why)


def test_version_mismatch_and_bad_files_fail(exe, tmp_path):
    ref, bed, prefixes, *_ = _family_files(tmp_path, 3, n_loci=8)
    vcf = prefixes[0] + ".sorted.vcf.gz"
    IOF.write_text(vcf, "##fileformat=VCFv4.2\n##trgtVersion=7.7.7\n#CHROM\n", bgzf=True)
    r = subprocess.run([exe, "dump", ref, bed, "50", str(tmp_path / "o.txt")] + prefixes, capture_output=True, text=True)
    assert r.returncode == 1 and "TRGT version mismatch" in r.stderr
    with open(prefixes[1] + ".spanning.sorted.bam", "wb") as o:
        o.write(IOF.bgzf_bytes(b"NOPE" + b"\0" * 64))
    r = subprocess.run([exe, "dump", ref, bed, "50", str(tmp_path / "o.txt")] + prefixes[1:], capture_output=True, text=True)
    assert r.returncode == 1 and "not a BAM file" in r.stderr
    os.remove(ref + ".fai")
    r = subprocess.run([exe, "dump", ref, bed, "50", str(tmp_path / "o.txt")] + prefixes[2:], capture_output=True, text=True)
    assert r.returncode == 1 and "Reference index file not found" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["trio", "duo"])
def test_tsv_from_files_identical(exe, tmp_path, mode):
    from importlib import import_module
    import trgt_denovo_b200  # noqa: F401
    from oracle import tdo, tdo_rows
    R = import_module("trgt_denovo_b200.rows")
    ref, bed, prefixes, loci, (f, m, c), *_ = _family_files(tmp_path, 4, n_loci=80)
    params = R.Params()
    aligner = tdo.Aligner()
    out_tsv, out_ids = str(tmp_path / "out.tsv"), str(tmp_path / "ids.txt")
    if mode == "trio":
        want = [tdo_rows.process_alleles_trio(loci[i], f[i], m[i], c[i], params, aligner) for i in range(len(loci))]
        wtsv, wids = tdo_rows.write_outputs(want, R.TRIO_COLUMNS)
        subprocess.run([exe, "trio", ref, bed, out_tsv, out_ids] + prefixes, check=True)
    else:
        want = [tdo_rows.process_alleles_duo(loci[i], c[i], m[i], params, aligner) for i in range(len(loci))]
        wtsv, wids = tdo_rows.write_outputs(want, R.DUO_COLUMNS)
        subprocess.run([exe, "duo", ref, bed, out_tsv, out_ids, prefixes[2], prefixes[1]], check=True)
    assert open(out_tsv).read() == wtsv
    assert open(out_ids).read() == wids
    assert wtsv.count("\n") > len(loci) // 2


def test_damaged_files_fail_cleanly(exe, tmp_path):
    """Truncated or bit-flipped BAM / VCF / catalog files: the readers report an error (exit code 1) or read what is
    still well-formed (0); they never crash."""
    ref, bed, prefixes, *_ = _family_files(tmp_path, 5, n_loci=12)
    rng = random.Random(5)
    targets = [prefixes[0] + ".spanning.sorted.bam", prefixes[0] + ".sorted.vcf.gz", bed]
    for trial in range(40):
        path = targets[trial % len(targets)]
        good = open(path, "rb").read()
        if trial % 2 == 0:
            bad = good[: rng.randrange(1, len(good))]
        else:
            b = bytearray(good)
            for _ in range(rng.randint(1, 6)):
                b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            bad = bytes(b)
        with open(path, "wb") as o:
            o.write(bad)
        r = subprocess.run([exe, "dump", ref, bed, "50", str(tmp_path / "o.txt")] + prefixes, capture_output=True, text=True)
        assert r.returncode in (0, 1), (trial, path, r.returncode, r.stderr[-300:])
        with open(path, "wb") as o:
            o.write(good)
    # undamaged again
    r = subprocess.run([exe, "dump", ref, bed, "50", str(tmp_path / "o.txt")] + prefixes, capture_output=True, text=True)
    assert r.returncode == 0


def test_damaged_bam_payload_fails_cleanly(exe, tmp_path):
    """Damage INSIDE the decompressed BAM stream (re-compressed, so the gzip layer is intact): record sizes, tag
    types and lengths that point outside the record."""
    ref, bed, prefixes, *_ = _family_files(tmp_path, 6, n_loci=10)
    import zlib
    path = prefixes[2] + ".spanning.sorted.bam"
    raw = b""
    d = zlib.decompressobj(31)
    data = open(path, "rb").read()
    while data:
        raw += d.decompress(data)
        data = d.unused_data
        d = zlib.decompressobj(31)
    rng = random.Random(6)
    for trial in range(60):
        b = bytearray(raw)
        for _ in range(rng.randint(1, 4)):
            b[rng.randrange(8, len(b))] = rng.randrange(256)
        with open(path, "wb") as o:
            o.write(IOF.bgzf_bytes(bytes(b)))
        r = subprocess.run([exe, "dump", ref, bed, "50", str(tmp_path / "o.txt")] + prefixes, capture_output=True, text=True)
        assert r.returncode in (0, 1), (trial, r.returncode, r.stderr[-300:])
