"""ctypes binding of the CPU ORACLE (oracle/libtdo_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under trgt-denovo_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libtdo_oracle.so")

STATUS_COMPLETED = 0
STATUS_UNATTAINABLE = -300

DEFAULT_PENALTIES = (8, 4, 2, 24, 1)  # src/model.rs:13-23
DEFAULT_HEURISTIC = (1, 10, 50, 1)  # WFA2 default wf-adaptive(10,50,1), SURVEY.md A.0
NO_HEURISTIC = (0, 0, 0, 0)


class Penalties(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("x", "o1", "e1", "o2", "e2")]


class Heuristic(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("kind", "min_wf_len", "max_dist", "steps")]


class Stats(C.Structure):
    _fields_ = [("cells", C.c_int64), ("ext16", C.c_int64), ("columns", C.c_int64),
                ("max_width", C.c_int32), ("steps", C.c_int32)]


class AlleleOut(C.Structure):
    _fields_ = [("denovo_coverage", C.c_int32), ("mean_diff_father", C.c_float),
                ("mean_diff_mother", C.c_float), ("mother_overlap", C.c_int32 * 2),
                ("father_overlap", C.c_int32 * 2), ("allele_origin", C.c_int32),
                ("denovo_status", C.c_int32), ("error", C.c_int32), ("top_mother", C.c_double),
                ("top_father", C.c_double), ("child_median", C.c_double), ("mean_col", C.c_double * 4)]


class TrioLocus(C.Structure):
    _fields_ = [("n_child", C.c_int32), ("n_mother", C.c_int32), ("n_father", C.c_int32),
                ("list_off", (C.c_uint32 * 6) * 2), ("child_len", C.c_int32 * 2),
                ("mother_len", C.c_int32 * 2), ("father_len", C.c_int32 * 2)]


def build(force=False):
    """Compile the oracle with gcc (building the checker is not using it)."""
    srcs = [os.path.join(_HERE, f) for f in ("tdo_wfa.c", "tdo_reduce.c", "tdo_batch.c", "tdo_oracle.h")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtdo_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.tdo_aligner_new.restype = C.c_void_p
        L.tdo_aligner_new.argtypes = [C.POINTER(Penalties), C.POINTER(Heuristic)]
        L.tdo_aligner_delete.argtypes = [C.c_void_p]
        L.tdo_aligner_set_heuristic.argtypes = [C.c_void_p, C.POINTER(Heuristic)]
        L.tdo_align_end_to_end.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.tdo_score.argtypes = [C.c_void_p]
        L.tdo_cigar_ops.restype = C.POINTER(C.c_char)
        L.tdo_cigar_ops.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.tdo_cigar_string.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.tdo_cigar_score_clipped.argtypes = [C.c_void_p, C.c_int]
        L.tdo_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.tdo_gotoh2p_penalty.restype = C.c_int64
        L.tdo_gotoh2p_penalty.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(Penalties)]
        L.tdo_align_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.POINTER(Penalties), C.POINTER(Heuristic), C.c_int,
                                      C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        L.tdo_quantile.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double)]
        L.tdo_median.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.tdo_get_top_other_score.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                              C.POINTER(C.c_double)]
        L.tdo_get_score_count_diff.restype = None
        L.tdo_get_score_count_diff.argtypes = [C.c_double, C.c_void_p, C.c_int, C.POINTER(C.c_int64),
                                               C.POINTER(C.c_float)]
        L.tdo_get_overlap_coverage.restype = C.c_int32
        L.tdo_get_overlap_coverage.argtypes = [C.c_double, C.c_void_p, C.c_int]
        L.tdo_trio_assess.restype = None
        L.tdo_trio_assess.argtypes = [C.POINTER(TrioLocus), C.c_void_p, C.c_double, C.POINTER(AlleleOut),
                                      C.c_void_p]
        L.tdo_duo_assess.restype = None
        L.tdo_duo_assess.argtypes = L.tdo_trio_assess.argtypes
        L.tdo_assess_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class Aligner:
    """Mirror of the slice of WFAligner (src/aligner.rs:174-606) the hot path uses."""

    def __init__(self, penalties=DEFAULT_PENALTIES, heuristic=DEFAULT_HEURISTIC):
        self._pen = Penalties(*penalties)
        self._h = lib().tdo_aligner_new(C.byref(self._pen), C.byref(Heuristic(*heuristic)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().tdo_aligner_delete(self._h)
            self._h = None

    def set_heuristic(self, heuristic):
        lib().tdo_aligner_set_heuristic(self._h, C.byref(Heuristic(*heuristic)))

    def align_end_to_end(self, pattern: bytes, text: bytes) -> int:
        return lib().tdo_align_end_to_end(self._h, pattern, len(pattern), text, len(text))

    def score(self) -> int:
        return lib().tdo_score(self._h)

    def cigar_ops(self) -> bytes:
        n = C.c_int(0)
        p = lib().tdo_cigar_ops(self._h, C.byref(n))
        return C.string_at(p, n.value)

    def cigar(self, clip=None) -> str:
        buf = C.create_string_buffer(1 << 20)
        n = lib().tdo_cigar_string(self._h, clip or 0, buf, len(buf))
        assert n >= 0
        return buf.value.decode()

    def cigar_score_clipped(self, clip: int) -> int:
        return lib().tdo_cigar_score_clipped(self._h, clip)

    def stats(self) -> Stats:
        st = Stats()
        lib().tdo_get_stats(self._h, C.byref(st))
        return st


def gotoh2p_penalty(pattern: bytes, text: bytes, penalties=DEFAULT_PENALTIES) -> int:
    return lib().tdo_gotoh2p_penalty(pattern, len(pattern), text, len(text), C.byref(Penalties(*penalties)))


def align_batch(blob, seq_off, seq_len, pair_pattern, pair_text, clip, penalties=DEFAULT_PENALTIES,
                heuristic=DEFAULT_HEURISTIC, n_threads=1):
    """Returns (scores int32[n], status int32[n], Stats total)."""
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    seq_len = np.ascontiguousarray(seq_len, dtype=np.uint32)
    pp = np.ascontiguousarray(pair_pattern, dtype=np.uint32)
    pt = np.ascontiguousarray(pair_text, dtype=np.uint32)
    n = len(pp)
    scores = np.zeros(n, dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    st = Stats()
    lib().tdo_align_batch(blob.ctypes.data, seq_off.ctypes.data, seq_len.ctypes.data, pp.ctypes.data,
                          pt.ctypes.data, n, C.byref(Penalties(*penalties)), C.byref(Heuristic(*heuristic)),
                          int(clip), int(n_threads), scores.ctypes.data, status.ctypes.data, C.byref(st))
    return scores, status, st


def quantile(xs, q):
    a = np.array(xs, dtype=np.float64)
    out = C.c_double()
    r = lib().tdo_quantile(a.ctypes.data, len(a), q, C.byref(out))
    return None if r else out.value


def median(xs):
    a = np.array(xs, dtype=np.int32)
    out = C.c_double()
    r = lib().tdo_median(a.ctypes.data, len(a), C.byref(out))
    return None if r else out.value


def get_score_count_diff(top, aligns):
    a = np.array(aligns, dtype=np.int32)
    c, d = C.c_int64(), C.c_float()
    lib().tdo_get_score_count_diff(top, a.ctypes.data, len(a), C.byref(c), C.byref(d))
    return c.value, d.value


def get_overlap_coverage(thr, lists):
    return [lib().tdo_get_overlap_coverage(thr, np.array(l, dtype=np.int32).ctypes.data, len(l)) for l in lists]


def make_locus(mother, father, child_lists, child_len=(0, 0), mother_len=(0, 0), father_len=(0, 0)):
    """Build a TrioLocus + flat score array from python lists.

    mother/father: per child allele c, a list of per-parent-allele score lists; child_lists[c] = own scores.
    """
    loc = TrioLocus()
    loc.n_child = len(child_lists)
    loc.n_mother = len(mother[0])
    loc.n_father = len(father[0]) if father else 0
    flat = []
    for c in range(loc.n_child):
        lists = [mother[c][0], mother[c][1] if loc.n_mother > 1 else [],
                 father[c][0] if father else [], father[c][1] if father and loc.n_father > 1 else [],
                 child_lists[c]]
        for j, l in enumerate(lists):
            loc.list_off[c][j] = len(flat)
            flat.extend(l)
        loc.list_off[c][5] = len(flat)
    for i in range(2):
        loc.child_len[i], loc.mother_len[i], loc.father_len[i] = child_len[i], mother_len[i], father_len[i]
    return loc, np.array(flat, dtype=np.int32)


def trio_assess(loc, scores, q=1.0, duo=False):
    out = (AlleleOut * 2)()
    mask = np.zeros(max(1, len(scores)), dtype=np.uint8)
    fn = lib().tdo_duo_assess if duo else lib().tdo_trio_assess
    fn(C.byref(loc), scores.ctypes.data, q, out, mask.ctypes.data)
    return [out[c] for c in range(loc.n_child)], mask


def assess_batch(loci_ptr, n_loci, scores, q=1.0, duo=False, n_threads=1):
    """loci_ptr: address of an array of TrioLocus (same layout as tdb_trio_locus).
    -> (AlleleOut array [2*n_loci], mask uint8[n_scores])"""
    scores = np.ascontiguousarray(scores, dtype=np.int32)
    out = (AlleleOut * (2 * max(1, n_loci)))()
    mask = np.zeros(max(1, len(scores)), dtype=np.uint8)
    lib().tdo_assess_batch(loci_ptr, n_loci, scores.ctypes.data, q, int(duo), int(n_threads),
                           C.cast(out, C.c_void_p), mask.ctypes.data)
    return out, mask
