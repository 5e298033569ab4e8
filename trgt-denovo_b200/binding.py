"""ctypes declarations for include/trgt_denovo_b200.h (one-to-one with the header)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TDB_LIB: load another build of the same library (tuning sweeps); the default is the in-tree build
LIB_PATH = os.environ.get("TDB_LIB") or os.path.join(_HERE, "libtrgt_denovo_b200.so")

OK = 0
E_INVALID, E_NO_DEVICE, E_CUDA, E_OOM, E_UNSUPPORTED = -1, -2, -3, -4, -5


class TdbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tdb error {code}: {msg}")
        self.code = code


class Penalties(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("mismatch", "gap_opening1", "gap_extension1", "gap_opening2",
                                         "gap_extension2")]


class Heuristic(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("kind", "min_wavefront_length", "max_distance_threshold",
                                         "steps_between_cutoffs")]


class Config(C.Structure):
    _fields_ = [("device_id", C.c_int32), ("penalties", Penalties), ("heuristic", Heuristic),
                ("clip_len", C.c_int32), ("reserved_flags", C.c_int32), ("scratch_bytes", C.c_uint64)]


class TrioLocus(C.Structure):
    _fields_ = [("n_child", C.c_int32), ("n_mother", C.c_int32), ("n_father", C.c_int32),
                ("list_off", (C.c_uint32 * 6) * 2), ("child_len", C.c_int32 * 2),
                ("mother_len", C.c_int32 * 2), ("father_len", C.c_int32 * 2)]


class AlleleOut(C.Structure):
    _fields_ = [("denovo_coverage", C.c_int32), ("mean_diff_father", C.c_float),
                ("mean_diff_mother", C.c_float), ("mother_overlap", C.c_int32 * 2),
                ("father_overlap", C.c_int32 * 2), ("allele_origin", C.c_int32),
                ("denovo_status", C.c_int32), ("error", C.c_int32), ("top_mother", C.c_double),
                ("top_father", C.c_double), ("child_median", C.c_double), ("mean_col", C.c_double * 4)]


class Counters(C.Structure):
    _fields_ = [("pairs", C.c_uint64), ("tier_pairs", C.c_uint64 * 4), ("cells", C.c_uint64),
                ("ext16", C.c_uint64), ("columns", C.c_uint64), ("exec_cells", C.c_uint64),
                ("exec_ext16", C.c_uint64), ("exec_columns", C.c_uint64), ("tier_cells", C.c_uint64 * 4),
                ("tier_ext16", C.c_uint64 * 4), ("tier_columns", C.c_uint64 * 4), ("pairs_identical", C.c_uint64),
                ("pairs_duplicate", C.c_uint64), ("pairs_closed_form", C.c_uint64), ("kernel_launches", C.c_uint64), ("chunks", C.c_uint64), ("host_packed", C.c_uint64), ("bytes_h2d", C.c_uint64), ("bytes_raw", C.c_uint64),
                ("ms_pack", C.c_float), ("ms_dedup", C.c_float), ("ms_align", C.c_float),
                ("ms_scatter", C.c_float), ("ms_reduce", C.c_float), ("ms_total_device", C.c_float),
                ("ms_host_plan", C.c_float), ("ms_host_pack_wait", C.c_float), ("ms_host_call", C.c_float),
                ("ms_host_packed", C.c_float), ("ms_host_enqueued", C.c_float),
                ("ms_tier", C.c_float * 4), ("ms_thread_tier", C.c_float), ("thread_pairs", C.c_uint64),
                ("thread_cells", C.c_uint64), ("thread_ext16", C.c_uint64), ("thread_columns", C.c_uint64),
                ("bytes_d2h", C.c_uint64), ("pairs_failed", C.c_uint64),
                ("ms_fused_call", C.c_float), ("ms_fused_tail", C.c_float),
                ("long_pairs", C.c_uint64), ("long_cells", C.c_uint64), ("long_ext16", C.c_uint64),
                ("long_columns", C.c_uint64), ("ms_long_tier", C.c_float), ("reserved0", C.c_float)]


class PackedSeqs(C.Structure):
    _fields_ = [("packed", C.c_void_p), ("n_bases", C.c_size_t), ("seq_off", C.c_void_p), ("seq_len", C.c_void_p),
                ("n_seqs", C.c_size_t), ("exc_seq", C.c_void_p), ("exc_bytes", C.c_void_p), ("n_exc", C.c_size_t),
                ("seq_len16", C.c_void_p)]


SCORE_FAILED = -2147483648


# every symbol include/trgt_denovo_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "tdb_default_config", "tdb_abi_version", "tdb_device_count", "tdb_create", "tdb_destroy",
    "tdb_last_error", "tdb_set_clip_len", "tdb_set_host_threads", "tdb_set_heuristic", "tdb_align_batch",
    "tdb_align_batch_device", "tdb_align_batch_cigar", "tdb_trio_reduce", "tdb_duo_reduce",
    "tdb_trio_batch", "tdb_duo_batch", "tdb_assess_batch_device", "tdb_align_end_to_end", "tdb_score",
    "tdb_cigar_ops", "tdb_cigar_score_clipped", "tdb_get_counters", "tdb_set_count_work",
    "tdb_host_alloc", "tdb_host_free", "tdb_packed_words", "tdb_host_pack", "tdb_host_pack_isa",
    "tdb_align_batch_packed", "tdb_trio_batch_packed", "tdb_duo_batch_packed",
]

_lib = None


def lib():
    """Load the product library.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TdbError(E_NO_DEVICE, f"{LIB_PATH} not built (run __graft_entry__.build()); no CPU fallback exists")
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, u64p = C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p
    L.tdb_default_config.restype = None
    L.tdb_default_config.argtypes = [C.POINTER(Config)]
    L.tdb_abi_version.restype = C.c_int
    L.tdb_device_count.restype = C.c_int
    L.tdb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.tdb_destroy.restype = None
    L.tdb_destroy.argtypes = [vp]
    L.tdb_last_error.restype = C.c_char_p
    L.tdb_last_error.argtypes = [vp]
    L.tdb_set_clip_len.argtypes = [vp, i32]
    L.tdb_set_host_threads.argtypes = [vp, i32]
    L.tdb_set_heuristic.argtypes = [vp, C.POINTER(Heuristic)]
    batch = [vp, vp, sz, u64p, vp, sz, vp, vp, sz]
    L.tdb_align_batch.argtypes = batch + [vp, vp]
    L.tdb_align_batch_device.argtypes = batch + [vp, vp]
    L.tdb_align_batch_cigar.argtypes = batch + [vp, vp, vp, vp, vp, vp]
    red = [vp, vp, sz, vp, sz, C.c_double, vp, vp]
    L.tdb_trio_reduce.argtypes = red
    L.tdb_duo_reduce.argtypes = red
    fused = batch + [vp, sz, C.c_double, vp, vp, vp]
    L.tdb_trio_batch.argtypes = fused
    L.tdb_duo_batch.argtypes = fused
    L.tdb_assess_batch_device.argtypes = [vp, i32] + batch[1:] + [vp, sz, C.c_double, vp, vp, vp]
    L.tdb_align_batch_packed.argtypes = [vp, C.POINTER(PackedSeqs), vp, vp, sz, vp, vp]
    pfused = [vp, C.POINTER(PackedSeqs), vp, vp, sz, vp, sz, C.c_double, vp, vp, vp]
    L.tdb_trio_batch_packed.argtypes = pfused
    L.tdb_duo_batch_packed.argtypes = pfused
    L.tdb_align_end_to_end.argtypes = [vp, C.c_char_p, i32, C.c_char_p, i32]
    L.tdb_score.restype = i32
    L.tdb_score.argtypes = [vp]
    L.tdb_cigar_ops.restype = C.POINTER(C.c_uint8)
    L.tdb_cigar_ops.argtypes = [vp, C.POINTER(i32)]
    L.tdb_cigar_score_clipped.restype = i32
    L.tdb_cigar_score_clipped.argtypes = [vp, i32]
    L.tdb_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.tdb_set_count_work.argtypes = [vp, i32]
    L.tdb_packed_words.restype = sz
    L.tdb_packed_words.argtypes = [sz]
    L.tdb_host_pack.argtypes = [vp, sz, vp, vp, sz, vp, vp, C.c_int]
    L.tdb_host_pack_isa.restype = C.c_char_p
    L.tdb_host_alloc.restype = vp
    L.tdb_host_alloc.argtypes = [sz]
    L.tdb_host_free.restype = None
    L.tdb_host_free.argtypes = [vp]
    _lib = L
    return L


def device_count():
    return lib().tdb_device_count()
