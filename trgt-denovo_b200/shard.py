"""Locus sharding across the GPUs of one box: independent units, no collective.

The reference treats loci as independent tasks (src/commands/shared.rs:136-145: rayon par_bridge over
the locus stream, one thread-local aligner each).  Here the host splits the locus list into one
contiguous range per GPU, balanced on the alignment work proxy sum(n + m) over the pairs of each
locus, runs one ctx per GPU, and concatenates the per-locus results in locus (BED) order.
"""
import numpy as np


def even_bounds(n_units: int, world: int):
    """world+1 boundaries of near-equal contiguous ranges."""
    return [(n_units * r) // world for r in range(world + 1)]


def balanced_bounds(weights, world: int):
    """Contiguous ranges whose weight sums are as equal as prefix sums allow (weights >= 0 per locus)."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if n == 0:
        return [0] * (world + 1)
    pre = np.concatenate([[0.0], np.cumsum(w)])
    total = pre[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(pre, target, side="left"))
        # the boundary whose prefix sum is closest to the target, never moving backwards
        if i > 0 and abs(pre[i - 1] - target) <= abs(pre[min(i, n)] - target):
            i -= 1
        bounds.append(min(max(i, bounds[-1]), n))
    bounds.append(n)
    return bounds


def locus_weights(loci, seq_len, pair_pattern, pair_text):
    """sum over the pairs of each locus of (n + m): list_off[c][0] .. list_off[c][5] are pair ranges."""
    sl = np.asarray(seq_len, dtype=np.int64)
    per_pair = sl[np.asarray(pair_pattern)] + sl[np.asarray(pair_text)]
    pre = np.concatenate([[0], np.cumsum(per_pair)])
    out = np.zeros(len(loci), dtype=np.int64)
    for i, loc in enumerate(loci):
        for c in range(loc.n_child):
            out[i] += pre[loc.list_off[c][5]] - pre[loc.list_off[c][0]]
    return out


def merge_in_locus_order(parts):
    """parts: list of (first_locus, rows) from the ranks -> rows in locus order."""
    out = []
    for _, rows in sorted(parts, key=lambda p: p[0]):
        out.extend(rows)
    return out
