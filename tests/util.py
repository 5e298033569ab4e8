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
