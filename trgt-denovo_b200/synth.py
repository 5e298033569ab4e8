"""Synthetic trio / duo workloads (include/tdb_synth.h).

Data generation only -- no alignment arithmetic.  Arrays are numpy views over the C allocation
(pinned host memory when pinned=True) and stay valid while the SynthBatch object is alive.
"""
import ctypes as C
import os

import numpy as np

from . import binding as B

# seeds fixed by SURVEY.md 8(d)
SEED_CONFIG1 = 20250001
SEED_CONFIG2 = 20250002
SEED_CONFIG4 = 20250004
SEED_CONFIG5 = 20250005
GENOME_WIDE_LOCI = 937_000


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("duo", C.c_int32), ("flank_len", C.c_int32),
                ("reads_per_allele", C.c_double), ("denovo_rate", C.c_double), ("sub_rate", C.c_double),
                ("indel_rate", C.c_double), ("stutter_prob", C.c_double), ("len_median", C.c_double),
                ("len_sigma", C.c_double), ("len_min", C.c_int32), ("len_max", C.c_int32),
                ("long_expansion", C.c_int32), ("pinned", C.c_int32)]


class SynthMeta(C.Structure):
    _fields_ = [("motif", C.c_char * 8), ("units", (C.c_int32 * 2) * 3), ("n_reads", (C.c_int32 * 2) * 3),
                ("denovo", C.c_int32), ("denovo_allele", C.c_int32), ("seq_begin", C.c_uint32),
                ("seq_end", C.c_uint32), ("pair_begin", C.c_uint32), ("pair_end", C.c_uint32)]


class _Batch(C.Structure):
    _fields_ = [("blob", C.c_void_p), ("blob_bytes", C.c_size_t), ("seq_off", C.c_void_p),
                ("seq_len", C.c_void_p), ("n_seqs", C.c_size_t), ("pair_pattern", C.c_void_p),
                ("pair_text", C.c_void_p), ("n_pairs", C.c_size_t), ("loci", C.c_void_p), ("meta", C.c_void_p),
                ("n_loci", C.c_size_t), ("sum_nm", C.c_uint64), ("sum_bases", C.c_uint64), ("pinned", C.c_int32)]


SYNTH_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtdb_synth.so")
_lib = None


def _L():
    """The generator's own library (trgt-denovo_b200/synth_src): loading it maps no product code."""
    global _lib
    if _lib is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise B.TdbError(B.E_INVALID, f"{SYNTH_LIB_PATH} not built (run __graft_entry__.build())")
        _lib = C.CDLL(SYNTH_LIB_PATH)
    L = _lib
    if not getattr(L, "_synth_ready", False):
        L.tdb_synth_default_params.restype = None
        L.tdb_synth_default_params.argtypes = [C.POINTER(SynthParams)]
        L.tdb_synth_generate.argtypes = [C.POINTER(SynthParams), C.c_uint64, C.c_uint64, C.c_int, C.POINTER(_Batch)]
        L.tdb_synth_free.restype = None
        L.tdb_synth_free.argtypes = [C.POINTER(_Batch)]
        L.tdb_synth_micro.argtypes = [C.c_uint64, C.c_uint32, C.c_double, C.c_size_t, C.c_int, C.c_int,
                                      C.POINTER(_Batch)]
        L._synth_ready = True
    return L


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class SynthBatch:
    def __init__(self, raw: _Batch):
        self._raw = raw
        r = raw
        self.blob = _view(r.blob, r.blob_bytes, np.uint8)
        self.seq_off = _view(r.seq_off, r.n_seqs, np.uint64)
        self.seq_len = _view(r.seq_len, r.n_seqs, np.uint32)
        self.pair_pattern = _view(r.pair_pattern, r.n_pairs, np.uint32)
        self.pair_text = _view(r.pair_text, r.n_pairs, np.uint32)
        self.n_loci, self.n_seqs, self.n_pairs = r.n_loci, r.n_seqs, r.n_pairs
        self.sum_nm, self.sum_bases = r.sum_nm, r.sum_bases
        self.loci = (B.TrioLocus * max(1, r.n_loci)).from_address(r.loci) if r.loci else None
        self.meta = (SynthMeta * max(1, r.n_loci)).from_address(r.meta) if r.meta else None
        self.loci_ptr = r.loci

    def seq(self, i) -> bytes:
        o, n = int(self.seq_off[i]), int(self.seq_len[i])
        return self.blob[o:o + n].tobytes()

    def close(self):
        if self._raw is not None:
            self.blob = self.seq_off = self.seq_len = self.pair_pattern = self.pair_text = None
            self.loci = self.meta = None
            _L().tdb_synth_free(C.byref(self._raw))
            self._raw = None

    __del__ = close


def default_params(**kw) -> SynthParams:
    p = SynthParams()
    _L().tdb_synth_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def generate(params: SynthParams, locus_begin: int, locus_end: int, n_threads=None) -> SynthBatch:
    raw = _Batch()
    nt = n_threads or os.cpu_count() or 1
    r = _L().tdb_synth_generate(C.byref(params), locus_begin, locus_end, nt, C.byref(raw))
    if r != 0:
        raise B.TdbError(r, "tdb_synth_generate failed")
    return SynthBatch(raw)


def micro(seed: int, length: int, divergence: float, n_pairs: int, n_threads=None, pinned=False) -> SynthBatch:
    raw = _Batch()
    r = _L().tdb_synth_micro(seed, length, divergence, n_pairs, n_threads or os.cpu_count() or 1, int(pinned),
                             C.byref(raw))
    if r != 0:
        raise B.TdbError(r, "tdb_synth_micro failed")
    return SynthBatch(raw)


def config(name: str, **overrides) -> SynthParams:
    """Named workloads of BASELINE.json / SURVEY.md 8(d)."""
    if name == "trio_1k":  # config 1
        p = default_params(seed=SEED_CONFIG1, denovo_rate=0.01)
    elif name == "trio_genome":  # config 2
        p = default_params(seed=SEED_CONFIG2, denovo_rate=1e-4)
    elif name == "duo_genome":  # config 3
        p = default_params(seed=SEED_CONFIG2, denovo_rate=1e-4, duo=1)
    elif name == "long_expansion":  # config 4
        p = default_params(seed=SEED_CONFIG4, long_expansion=1, reads_per_allele=30.0)
    else:
        raise ValueError(name)
    for k, v in overrides.items():
        setattr(p, k, v)
    return p
