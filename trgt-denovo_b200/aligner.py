"""Host-side mirror of the aligner surface of the reference (src/aligner.rs, src/commands/shared.rs).

Names, argument meaning and error behaviour follow the reference so that parity tests read like the
reference's own tests:
  AlignmentStatus              src/aligner.rs:64-86
  WFAligner.align_end_to_end   src/aligner.rs:248-260
  WFAligner.score              src/aligner.rs:290-292
  WFAligner.cigar              src/aligner.rs:404-436
  WFAligner.cigar_wfa          src/aligner.rs:438-448
  WFAligner.cigar_score_clipped src/aligner.rs:359-376
  WFAligner.set_heuristic      src/aligner.rs:552-605 (None / WFadaptive)
  create_aligner_with_scoring  src/commands/shared.rs:39-51
All arithmetic happens on the GPU behind the C ABI.
"""
import ctypes as C
import enum

import numpy as np

from . import binding as B


class AlignmentStatus(enum.IntEnum):
    StatusAlgCompleted = 0
    StatusAlgPartial = 1
    StatusMaxStepsReached = -100
    StatusOOM = -200
    StatusUnattainable = -300


def _ptr(a):
    return None if a is None else a.ctypes.data


class Context:
    """One tdb_ctx: bound to one GPU, single owner (like a thread-local WFAligner)."""

    def __init__(self, device_id=0, penalties=(8, 4, 2, 24, 1), heuristic=(1, 10, 50, 1), clip_len=50,
                 scratch_bytes=0, count_work=False):
        L = B.lib()
        cfg = B.Config()
        L.tdb_default_config(C.byref(cfg))
        cfg.device_id = device_id
        cfg.penalties = B.Penalties(*penalties)
        cfg.heuristic = B.Heuristic(*heuristic)
        cfg.clip_len = clip_len
        cfg.scratch_bytes = scratch_bytes
        h = C.c_void_p()
        r = L.tdb_create(C.byref(cfg), C.byref(h))
        if r != B.OK:
            raise B.TdbError(r, L.tdb_last_error(None).decode())
        self._h = h
        if count_work:
            L.tdb_set_count_work(h, 1)
        self.penalties = tuple(penalties)
        self.clip_len = clip_len
        self.device_id = device_id

    def close(self):
        if getattr(self, "_h", None):
            B.lib().tdb_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, r):
        if r != B.OK:
            raise B.TdbError(r, B.lib().tdb_last_error(self._h).decode())

    def set_clip_len(self, clip_len):
        self._check(B.lib().tdb_set_clip_len(self._h, clip_len))
        self.clip_len = clip_len

    def set_host_threads(self, n_threads):
        self._check(B.lib().tdb_set_host_threads(self._h, int(n_threads)))

    def set_heuristic(self, heuristic):
        self._check(B.lib().tdb_set_heuristic(self._h, C.byref(B.Heuristic(*heuristic))))

    def set_count_work(self, on):
        """Exact algorithmic C/E/T counting (tdb_set_count_work): every content goes through the wavefronts."""
        self._check(B.lib().tdb_set_count_work(self._h, int(bool(on))))

    def counters(self):
        c = B.Counters()
        self._check(B.lib().tdb_get_counters(self._h, C.byref(c)))
        return c

    # ---- batched alignment (host buffers) ---------------------------------------------------------
    def align_batch(self, blob, seq_off, seq_len, pair_pattern, pair_text):
        """-> (scores int32[n_pairs], status int32[n_pairs]); src/denovo.rs:22-54 for a pair list."""
        blob, seq_off, seq_len, pp, pt = _canon(blob, seq_off, seq_len, pair_pattern, pair_text)
        n = len(pp)
        scores = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        self._check(B.lib().tdb_align_batch(self._h, _ptr(blob), blob.size, _ptr(seq_off), _ptr(seq_len),
                                            len(seq_len), _ptr(pp), _ptr(pt), n, _ptr(scores), _ptr(status)))
        return scores, status

    def align_batch_cigar(self, blob, seq_off, seq_len, pair_pattern, pair_text):
        """-> (scores, status, penalty, [cigar bytes per pair])."""
        blob, seq_off, seq_len, pp, pt = _canon(blob, seq_off, seq_len, pair_pattern, pair_text)
        n = len(pp)
        caps = (seq_len[pp].astype(np.uint64) + seq_len[pt].astype(np.uint64)) if n else np.zeros(0, np.uint64)
        coff = np.zeros(n, dtype=np.uint64)
        if n:
            coff[1:] = np.cumsum(caps)[:-1]
        total = int(caps.sum()) if n else 0
        cbuf = np.zeros(total + 16, dtype=np.uint8)
        clen = np.zeros(n, dtype=np.uint32)
        scores = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        pen = np.zeros(n, dtype=np.int32)
        self._check(B.lib().tdb_align_batch_cigar(self._h, _ptr(blob), blob.size, _ptr(seq_off), _ptr(seq_len),
                                                  len(seq_len), _ptr(pp), _ptr(pt), n, _ptr(scores), _ptr(status),
                                                  _ptr(pen), _ptr(cbuf), _ptr(coff), _ptr(clen)))
        cigs = [cbuf[int(coff[i]):int(coff[i]) + int(clen[i])].tobytes() for i in range(n)]
        return scores, status, pen, cigs

    # ---- packed input (the caller packed at load time; include/trgt_denovo_b200.h tdb_packed_seqs) ----------
    def align_batch_packed(self, packed: "PackedInput", pair_pattern, pair_text):
        pp = np.ascontiguousarray(pair_pattern, dtype=np.uint32)
        pt = np.ascontiguousarray(pair_text, dtype=np.uint32)
        n = len(pp)
        scores = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        self._check(B.lib().tdb_align_batch_packed(self._h, C.byref(packed.struct), _ptr(pp), _ptr(pt), n,
                                                   _ptr(scores), _ptr(status)))
        return scores, status

    def assess_batch_packed(self, packed: "PackedInput", pair_pattern, pair_text, loci, q=1.0, duo=False):
        pp = np.ascontiguousarray(pair_pattern, dtype=np.uint32)
        pt = np.ascontiguousarray(pair_text, dtype=np.uint32)
        n, n_loci = len(pp), len(loci)
        out = (B.AlleleOut * (2 * max(1, n_loci)))()
        scores = np.zeros(max(1, n), dtype=np.int32)
        mask = np.zeros(max(1, n), dtype=np.uint8)
        fn = B.lib().tdb_duo_batch_packed if duo else B.lib().tdb_trio_batch_packed
        self._check(fn(self._h, C.byref(packed.struct), _ptr(pp), _ptr(pt), n, C.cast(loci, C.c_void_p), n_loci, q,
                       C.cast(out, C.c_void_p), _ptr(scores), _ptr(mask)))
        return out, scores[:n], mask[:n]

    # ---- reductions / fused ---------------------------------------------------------------------------
    def reduce(self, loci, scores, q=1.0, duo=False):
        """loci: ctypes array of TrioLocus.  -> (AlleleOut array [2*n_loci], mask uint8[n_scores])."""
        scores = np.ascontiguousarray(scores, dtype=np.int32)
        n_loci = len(loci)
        out = (B.AlleleOut * (2 * max(1, n_loci)))()
        mask = np.zeros(max(1, len(scores)), dtype=np.uint8)
        fn = B.lib().tdb_duo_reduce if duo else B.lib().tdb_trio_reduce
        self._check(fn(self._h, C.cast(loci, C.c_void_p), n_loci, _ptr(scores), len(scores), q,
                       C.cast(out, C.c_void_p), _ptr(mask)))
        return out, mask

    def assess_batch(self, blob, seq_off, seq_len, pair_pattern, pair_text, loci, q=1.0, duo=False):
        """Fused alignment + reduction (trio/duo assess_denovo for many loci).
        -> (AlleleOut array [2*n_loci], scores int32[n_pairs], mask uint8[n_pairs])"""
        blob, seq_off, seq_len, pp, pt = _canon(blob, seq_off, seq_len, pair_pattern, pair_text)
        n = len(pp)
        n_loci = len(loci)
        out = (B.AlleleOut * (2 * max(1, n_loci)))()
        scores = np.zeros(max(1, n), dtype=np.int32)
        mask = np.zeros(max(1, n), dtype=np.uint8)
        fn = B.lib().tdb_duo_batch if duo else B.lib().tdb_trio_batch
        self._check(fn(self._h, _ptr(blob), blob.size, _ptr(seq_off), _ptr(seq_len), len(seq_len), _ptr(pp),
                       _ptr(pt), n, C.cast(loci, C.c_void_p), n_loci, q, C.cast(out, C.c_void_p), _ptr(scores),
                       _ptr(mask)))
        return out, scores[:n], mask[:n]


class PackedInput:
    """The arrays behind a tdb_packed_seqs (kept alive with the struct that points at them)."""

    def __init__(self, packed, n_bases, seq_off, seq_len, exc_seq, exc_bytes, len16=False):
        self.packed = np.ascontiguousarray(packed, dtype=np.uint32)
        self.seq_off = None if seq_off is None else np.ascontiguousarray(seq_off, dtype=np.uint64)
        self.seq_len = np.ascontiguousarray(seq_len, dtype=np.uint32)
        # len16: hand the lengths over as uint16 (tdb_packed_seqs.seq_len16), every sequence < 65 536 bases
        self.seq_len16 = self.seq_len.astype(np.uint16) if len16 else None
        if len16 and len(self.seq_len) and int(self.seq_len.max()) > 0xffff:
            raise ValueError("len16 needs every sequence shorter than 65 536 bases")
        self.exc_seq = np.ascontiguousarray(exc_seq, dtype=np.uint32)
        self.exc_bytes = np.ascontiguousarray(exc_bytes, dtype=np.uint8)
        self.n_bases = int(n_bases)
        self.struct = B.PackedSeqs(_ptr(self.packed), self.n_bases, _ptr(self.seq_off),
                                   None if len16 else _ptr(self.seq_len),
                                   len(self.seq_len), _ptr(self.exc_seq) if len(self.exc_seq) else None,
                                   _ptr(self.exc_bytes) if len(self.exc_seq) else None, len(self.exc_seq),
                                   _ptr(self.seq_len16))


def pack_sequences(blob, seq_off, seq_len, n_threads=1, dense=None, len16=False) -> PackedInput:
    """ASCII blob -> the packed form of the *_packed entry points, through tdb_host_pack (what a loader that holds
    ASCII would run once; a loader that holds BAM 4-bit bases emits the same stream directly, tdbh::PackedBatch).
    dense: hand the library seq_off = NULL (default: when the sequences are back to back)."""
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    seq_len = np.ascontiguousarray(seq_len, dtype=np.uint32)
    n = len(seq_len)
    ends = np.cumsum(seq_len, dtype=np.uint64)
    back_to_back = np.concatenate([[0], ends[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
    off = back_to_back if seq_off is None else np.ascontiguousarray(seq_off, dtype=np.uint64)
    L = B.lib()
    packed = np.zeros(L.tdb_packed_words(blob.size), dtype=np.uint32)
    flag = np.zeros(max(1, n), dtype=np.uint8)
    r = L.tdb_host_pack(_ptr(blob), blob.size, _ptr(off), _ptr(seq_len), n, _ptr(packed), _ptr(flag), int(n_threads))
    if r != B.OK:
        raise B.TdbError(r, "tdb_host_pack: layout contract violated")
    exc = np.nonzero(flag[:n])[0].astype(np.uint32)
    raw = b"".join(blob[int(off[i]):int(off[i]) + int(seq_len[i])].tobytes() for i in exc)
    if dense is None:
        dense = n == 0 or np.array_equal(off, back_to_back)
    return PackedInput(packed, blob.size, None if dense else off, seq_len, exc, np.frombuffer(raw, dtype=np.uint8), len16=len16)


def _canon(blob, seq_off, seq_len, pp, pt):
    # seq_off None: dense layout (sequences back to back), the library derives the offsets
    return (np.ascontiguousarray(blob, dtype=np.uint8),
            None if seq_off is None else np.ascontiguousarray(seq_off, dtype=np.uint64),
            np.ascontiguousarray(seq_len, dtype=np.uint32), np.ascontiguousarray(pp, dtype=np.uint32),
            np.ascontiguousarray(pt, dtype=np.uint32))


class WFAligner:
    """Single-pair shim with the reference's WFAligner call shape (gap-affine 2-piece only)."""

    def __init__(self, ctx: Context):
        self._ctx = ctx

    def align_end_to_end(self, pattern: bytes, text: bytes) -> AlignmentStatus:
        r = B.lib().tdb_align_end_to_end(self._ctx._h, pattern, len(pattern), text, len(text))
        try:
            return AlignmentStatus(r)
        except ValueError:
            # src/aligner.rs:83: unknown status -> panic; a TDB_E_* code surfaces as an exception here
            raise B.TdbError(r, B.lib().tdb_last_error(self._ctx._h).decode())

    def score(self) -> int:
        return B.lib().tdb_score(self._ctx._h)

    def cigar_wfa(self) -> bytes:
        n = C.c_int32(0)
        p = B.lib().tdb_cigar_ops(self._ctx._h, C.byref(n))
        return C.string_at(p, n.value) if n.value else b""

    def cigar(self, flank_len=None) -> str:
        ops = self.cigar_wfa()
        off = flank_len or 0
        begin, end = off, len(ops) - off
        if begin >= end:
            return ""
        out, last, run = [], ops[begin], 1
        for i in range(begin + 1, end):
            if ops[i] == last:
                run += 1
            else:
                out.append(f"{run}{chr(last)}")
                last, run = ops[i], 1
        out.append(f"{run}{chr(last)}")
        return "".join(out)

    def cigar_score_clipped(self, flank_len: int) -> int:
        return B.lib().tdb_cigar_score_clipped(self._ctx._h, flank_len)

    def cigar_score(self) -> int:
        return self.cigar_score_clipped(0)

    def set_heuristic(self, heuristic):
        """heuristic: None, or ("WFadaptive", min_wavefront_length, max_distance_threshold, steps)."""
        if heuristic is None:
            self._ctx.set_heuristic((0, 0, 0, 0))
        else:
            kind, a, b, c = heuristic
            if kind != "WFadaptive":
                raise B.TdbError(B.E_UNSUPPORTED, f"heuristic {kind} is not on the hot path")
            self._ctx.set_heuristic((1, a, b, c))


def create_aligner_with_scoring(device_id=0, clip_len=50) -> WFAligner:
    """src/commands/shared.rs:39-51: GapAffine2p 8,4,2,24,1, Alignment scope, default heuristic."""
    return WFAligner(Context(device_id=device_id, clip_len=clip_len))
