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


def locus_weights_from_meta(batch):
    """The same weights for a generated batch, vectorised over its per-locus pair ranges (SynthMeta.pair_begin/end)."""
    import ctypes as C
    n = batch.n_loci
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    sl = batch.seq_len.astype(np.int64)
    per_pair = sl[batch.pair_pattern] + sl[batch.pair_text]
    pre = np.concatenate([[0], np.cumsum(per_pair)])
    meta = np.frombuffer((C.c_uint8 * (C.sizeof(batch.meta[0]) * n)).from_address(C.addressof(batch.meta)), dtype=np.uint8)
    meta = meta.reshape(n, -1)
    rec = C.sizeof(batch.meta[0])
    pb = meta[:, rec - 8:rec - 4].copy().view(np.uint32)[:, 0].astype(np.int64)  # pair_begin, pair_end: the last two fields
    pe = meta[:, rec - 4:rec].copy().view(np.uint32)[:, 0].astype(np.int64)
    return pre[pe] - pre[pb]


def merge_in_locus_order(parts):
    """parts: list of (first_locus, rows) from the ranks -> rows in locus order."""
    out = []
    for _, rows in sorted(parts, key=lambda p: p[0]):
        out.extend(rows)
    return out


class MultiGpu:
    """One Context (= tdb_ctx) and one host thread per GPU; the locus list runs in contiguous ranges balanced on
    sum(n + m), rows come back in locus (BED) order.  Mirrors src/commands/shared.rs:136-145 (loci are independent
    tasks; rayon par_bridge there, one GPU per range here).  n_gpus may exceed the visible devices (contexts then
    share devices round-robin)."""

    def __init__(self, n_gpus=None, **ctx_kw):
        from . import aligner, binding
        visible = binding.device_count()
        if visible <= 0:
            raise binding.TdbError(binding.E_NO_DEVICE, "no CUDA device available (this library has no CPU path)")
        n = n_gpus or visible
        self.ctxs = [aligner.Context(device_id=i % visible, **ctx_kw) for i in range(n)]
        import os
        for c in self.ctxs:
            c.set_host_threads(max(1, (os.cpu_count() or 1) // n))
        self.bounds = None

    def close(self):
        for c in self.ctxs:
            c.close()

    @staticmethod
    def locus_weight(parent1, parent2, child):
        if child is None or parent1 is None:
            return 1
        reads = bases = 0
        for p in (parent1, parent2):
            if p is not None:
                reads += len(p.reads)
                bases += sum(len(r.bases) for r in p.reads)
        own_bases = sum(len(r.bases) for r in child.reads)
        tgt = sum(len(a) for a in child.allele_seqs)
        n_t = max(1, len(child.allele_seqs))
        return 1 + n_t * bases + reads * tgt + own_bases + len(child.reads) * (tgt // n_t)

    def _run(self, n, weights, fn):
        import threading
        self.bounds = balanced_bounds(weights, len(self.ctxs))
        parts, errs = [None] * len(self.ctxs), []

        def work(g):
            try:
                lo, hi = self.bounds[g], self.bounds[g + 1]
                parts[g] = (lo, fn(self.ctxs[g], lo, hi) if hi > lo else [])
            except Exception as e:  # noqa: BLE001 -- re-raised on the calling thread
                errs.append(e)
        th = [threading.Thread(target=work, args=(g,)) for g in range(len(self.ctxs))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return merge_in_locus_order(parts)

    def process_trio_batch(self, loci, father, mother, child, params):
        from . import rows
        w = [self.locus_weight(mother[i], father[i], child[i]) for i in range(len(loci))]
        return self._run(len(loci), w, lambda ctx, lo, hi: rows.process_trio_batch(
            ctx, loci[lo:hi], father[lo:hi], mother[lo:hi], child[lo:hi], params))

    def process_duo_batch(self, loci, sample_a, sample_b, params):
        from . import rows
        w = [self.locus_weight(sample_b[i], None, sample_a[i]) for i in range(len(loci))]
        return self._run(len(loci), w, lambda ctx, lo, hi: rows.process_duo_batch(
            ctx, loci[lo:hi], sample_a[lo:hi], sample_b[lo:hi], params))
