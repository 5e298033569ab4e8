"""Shared helpers for the test-suite: seeded synthetic reads / alleles (ASCII)."""
import random

import numpy as np


def rand_seq(rng: random.Random, n: int, alphabet="ACGT") -> str:
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng: random.Random, s: str, div: float, alphabet="ACGT") -> str:
    out = []
    for ch in s:
        r = rng.random()
        if r < div / 3:
            out.append(rng.choice(alphabet))
        elif r < 2 * div / 3:
            pass
        elif r < div:
            out.append(ch)
            out.append(rng.choice(alphabet))
        else:
            out.append(ch)
    return "".join(out) or "A"


def str_locus(rng: random.Random, min_units=3, max_units=40, flank=50):
    lf, rf = rand_seq(rng, flank), rand_seq(rng, flank)
    motif = rand_seq(rng, rng.randint(1, 6))
    return lf, motif, rf


def make_pairs(rng: random.Random, n_targets: int, reads_per_target: int, max_units=40, err=0.01,
               alphabet="ACGT"):
    """STR-like (read, target) pairs: reads carry +-few motif units and sequencing errors."""
    seqs, pp, pt = [], [], []
    for _ in range(n_targets):
        lf, motif, rf = str_locus(rng)
        units = rng.randint(3, max_units)
        t = len(seqs)
        seqs.append((lf + motif * units + rf).encode())
        for _ in range(reads_per_target):
            u = max(1, units + rng.choice([0, 0, 0, 1, -1, 2, -2, 5, -5, 12]))
            read = mutate(rng, lf + motif * u + rf, err, alphabet)
            pp.append(len(seqs))
            pt.append(t)
            seqs.append(read.encode())
    return seqs, np.array(pp, dtype=np.uint32), np.array(pt, dtype=np.uint32)


def to_arrays(seqs):
    blob = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    offs = np.zeros(len(seqs), dtype=np.uint64)
    if len(seqs):
        offs[1:] = np.cumsum(lens.astype(np.uint64))[:-1]
    return blob, offs, lens


def make_family_loci(rng: random.Random, n_loci: int, duo=False, missing_rate=0.05, denovo_rate=0.15,
                     reads_per_allele=8, err=0.004):
    """Synthetic per-locus inputs as `load_alleles` would see them (TRGT VCF alleles + spanning reads with AL/HP
    tags): -> (loci, father, mother, child) lists (duo: father entries are None and unused)."""
    from importlib import import_module
    import trgt_denovo_b200  # noqa: F401
    R = import_module("trgt_denovo_b200.rows")
    loci, fam = [], {"f": [], "m": [], "c": []}
    for li in range(n_loci):
        lf, motif, rf = str_locus(rng)
        base = rng.randint(3, 30)
        chrom = rng.choice(["chr1", "chr7", "chrX", "chrY"])
        loci.append(R.Locus(chrom=chrom, start=1000 * li + 17, end=1000 * li + 17 + base * len(motif),
                            motifs=[motif], trid=f"TR{li:05d}_{motif}"))
        units = {}
        for s in "fm":
            n_all = 1 if rng.random() < 0.15 else 2
            units[s] = [max(1, base + rng.randint(-2, 2)) for _ in range(n_all)]
        cu = [rng.choice(units["m"]), rng.choice(units["f"])]
        if rng.random() < 0.15:
            cu = cu[:1]
        if rng.random() < denovo_rate:
            k = rng.randrange(len(cu))
            cu[k] = max(1, cu[k] + rng.choice([-1, 1]) * rng.randint(1, 8))
        units["c"] = cu
        for s in "fmc":
            if rng.random() < missing_rate:
                fam[s].append(None)
                continue
            seqs = [(lf + motif * u + rf).encode() for u in units[s]]
            reads = []
            for ai, u in enumerate(units[s]):
                for k in range(max(1, int(rng.gauss(reads_per_allele, 2)))):
                    uu = max(1, u + (rng.choice([-1, 1]) if rng.random() < 0.05 else 0))
                    bases = mutate(rng, lf + motif * uu + rf, err).encode()
                    reads.append(R.Read(name=f"{s}{li}_{ai}_{k}/ccs", bases=bases, classification=ai,
                                        haplotype=rng.choice([1, 2, None, None])))
            rng.shuffle(reads)
            fam[s].append(R.SampleInput(allele_seqs=seqs, motif_counts=[str(u) for u in units[s]],
                                        allele_lengths=[str(u * len(motif)) for u in units[s]],
                                        genotype_indices=list(range(1, len(seqs) + 1)), reads=reads,
                                        karyotype=rng.choice(["XX", "XY"])))
    return loci, fam["f"], fam["m"], fam["c"]
