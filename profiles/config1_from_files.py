#!/usr/bin/env python3
"""BASELINE config 1 the way the reference runs it: the synthetic 1 000-locus trio (seed 20250001) written out as a
repeat catalog, an indexed reference FASTA and, per sample, a TRGT-style VCF and a spanning-read BAM; then
files -> TSV + read ids through the C++ front end and host side (tests/cpp/files_tsv.cpp), checked against (a) the
CPU restatement on the same inputs (oracle/tdo_rows.py) and (b) the Python in-memory GPU pipeline.

  python profiles/config1_from_files.py [n_loci]      (GPU box; needs g++ and zlib)"""
import json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
entry.build()
from importlib import import_module
import trgt_denovo_b200 as tdb
from oracle import tdo, tdo_rows
from tests import io_fixtures as IOF
S = import_module("trgt_denovo_b200.synth")
R = import_module("trgt_denovo_b200.rows")

n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
b = S.generate(S.config("trio_1k"), 0, n_loci, os.cpu_count() or 1)
SAMPLES = ("mother", "father", "child")  # the generator's sample order
KARY = {"mother": "XX", "father": "XY", "child": "XX"}
loci, fam, ref_alleles, flanks = [], {s: [] for s in SAMPLES}, [], []
for i in range(n_loci):
    m = b.meta[i]
    motif = m.motif.decode()
    n_all = [sum(1 for a in range(2) if m.units[s][a] > 0) for s in range(3)]
    child0 = b.seq(m.seq_begin)
    lf, rf = child0[:50].decode(), child0[-50:].decode()
    si = m.seq_begin + n_all[2]
    ref_units = m.units[0][0]
    start = 2000 * i + 1000
    loci.append(R.Locus(chrom="chr1", start=start, end=start + ref_units * len(motif), motifs=[motif], trid=f"SYN{i:06d}"))
    ref_alleles.append(motif * ref_units), flanks.append((lf, rf))
    for s, name in enumerate(SAMPLES):
        seqs, reads = [], []
        for a in range(n_all[s]):
            seqs.append((lf + motif * m.units[s][a] + rf).encode())
            if s == 2:
                assert seqs[-1] == b.seq(m.seq_begin + a)
            for k in range(m.n_reads[s][a]):
                reads.append(R.Read(name=f"{name}/{i}/{a}/{k}/ccs", bases=b.seq(si), classification=a, haplotype=None))
                si += 1
        fam[name].append(R.SampleInput(allele_seqs=seqs, motif_counts=[str(m.units[s][a]) for a in range(n_all[s])],
                                       allele_lengths=[str(m.units[s][a] * len(motif)) for a in range(n_all[s])],
                                       genotype_indices=list(range(1, n_all[s] + 1)), reads=reads, karyotype=KARY[name])
                         if reads else None)
    assert si == m.seq_end
tmp = tempfile.mkdtemp(prefix="tdb_config1_")
length = 2000 * n_loci + 4000
contig = bytearray(b"N" * length)
for l, ra, (lf, rf) in zip(loci, ref_alleles, flanks):
    contig[l.start - 50: l.end + 50] = (lf + ra + rf).encode()
ref = os.path.join(tmp, "ref.fa")
IOF.write_fasta(ref, [("chr1", contig.decode())])
bed = os.path.join(tmp, "catalog.bed.gz")
IOF.write_text(bed, "".join(f"chr1\t{l.start}\t{l.end}\tID={l.trid};MOTIFS={l.motifs[0]};STRUC=({l.motifs[0]})n\n" for l in loci), bgzf=True)
prefix = {}
bytes_bam = 0
for name in SAMPLES:
    vcf = ["##fileformat=VCFv4.2", "##trgtVersion=1.2.0", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + name]
    recs = []
    hdr = f"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chr1\tLN:{length}\n@PG\tID:trgt\tPN:trgt\tVN:1.2.0\tCL:trgt genotype --karyotype {KARY[name]}\n"
    for l, x, ra in zip(loci, fam[name], ref_alleles):
        if x is None:
            continue
        al = [a[50:len(a) - 50].decode() for a in x.allele_seqs]
        vcf.append("\t".join([l.chrom, str(l.start), ".", "N" + ra, ",".join("N" + a for a in al), ".", ".",
                              f"TRID={l.trid};END={l.end};MOTIFS={l.motifs[0]}", "GT:AL:MC",
                              ":".join(["/".join(str(g) for g in x.genotype_indices), ",".join(x.allele_lengths), ",".join(x.motif_counts)])]))
        for r in x.reads:
            recs.append(IOF.bam_record(r.name, 0, l.start - 50, r.bases.decode(),
                                       [("TR", "Z", l.trid), ("rq", "f", 0.999), ("AL", "C", r.classification), ("SO", "i", 0), ("EO", "i", 0)]))
    prefix[name] = os.path.join(tmp, name)
    IOF.write_text(prefix[name] + ".sorted.vcf.gz", "\n".join(vcf) + "\n", bgzf=True)
    IOF.write_bam(prefix[name] + ".spanning.sorted.bam", hdr, [("chr1", length)], recs)
    bytes_bam += os.path.getsize(prefix[name] + ".spanning.sorted.bam")
exe = os.path.join(tmp, "files_tsv")
libdir = os.path.dirname(entry.LIB)
subprocess.check_call(["g++", "-std=c++17", "-O2"] + (["-DTDBH_TIMING"] if os.environ.get("TDBH_TIMING") else []) + ["-I", os.path.join(entry.ROOT, "include"), os.path.join(entry.ROOT, "tests", "cpp", "files_tsv.cpp"),
                       "-o", exe, "-L", libdir, "-ltrgt_denovo_b200", "-lz", "-pthread", "-Wl,-rpath," + libdir])
out_tsv, out_ids = os.path.join(tmp, "out.tsv"), os.path.join(tmp, "ids.txt")
cmd = [exe, "trio", ref, bed, out_tsv, out_ids, prefix["father"], prefix["mother"], prefix["child"]]
subprocess.run(cmd, check=True, capture_output=True)  # warm the page cache and the CUDA context cache
t0 = time.perf_counter()
r = subprocess.run(cmd, check=True, capture_output=True, text=True)
wall = time.perf_counter() - t0
if os.environ.get("TDBH_TIMING"):
    sys.stderr.write(r.stderr)
timing = {k: float(v) for k, v in (x.split("=") for x in r.stderr.strip().splitlines()[-1].split()[1:])}
params = R.Params()
t0 = time.perf_counter()
aligner = tdo.Aligner()
want = [tdo_rows.process_alleles_trio(loci[i], fam["father"][i], fam["mother"][i], fam["child"][i], params, aligner) for i in range(n_loci)]
wtsv, wids = tdo_rows.write_outputs(want, R.TRIO_COLUMNS)
cpu_s = time.perf_counter() - t0
ctx = tdb.Context(device_id=0)
got = R.process_trio_batch(ctx, loci, fam["father"], fam["mother"], fam["child"], params)
print(json.dumps({"config": "1: 1000-locus trio from VCF + BAM files", "loci": n_loci,
                  "reads": sum(len(x.reads) for s in SAMPLES for x in fam[s] if x), "bam_bytes": bytes_bam,
                  "files_to_tsv_wall_s": wall, "files_to_tsv_loci_per_s": n_loci / wall, "stages_ms": timing,
                  "tsv_identical_to_cpu_restatement": open(out_tsv).read() == wtsv,
                  "read_ids_identical_to_cpu_restatement": open(out_ids).read() == wids,
                  "tsv_identical_to_python_gpu_pipeline": R.format_tsv(got, R.TRIO_COLUMNS) == wtsv,
                  "cpu_restatement_python_rows_s": cpu_s, "tsv_rows": wtsv.count("\n") - 1}))
