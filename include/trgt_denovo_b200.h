/*
 * trgt_denovo_b200.h -- C ABI of the B200-native read<->allele alignment hot path of trgt-denovo.
 *
 * This library replaces what sits behind the reference's `src/wfa2.rs` (`pub use wfa2_sys::*`, the
 * bindgen'd WFA2-lib C API) for the ONE configuration the binary uses, and the per-read score
 * reductions of `src/denovo.rs`, `src/trio/denovo.rs`, `src/duo/denovo.rs`, `src/math.rs`.
 * Plain pointers and sizes only; no CUDA or torch types in any signature.
 *
 * Reference interface replaced by each entry point (paths relative to /root/reference):
 *
 *   tdb_create / tdb_destroy        wfa2::wavefront_aligner_new / wavefront_aligner_delete through
 *                                   WFAlignerGapAffine2Pieces::create_aligner_with_match
 *                                   (src/aligner.rs:712-740) and Drop (src/aligner.rs:237-245);
 *                                   configuration = create_aligner_with_scoring()
 *                                   (src/commands/shared.rs:39-51) + Params.clip_len (:59-70)
 *   tdb_align_batch[_device]        the loops of align_alleleset / align_allele (src/denovo.rs:22-54):
 *                                   per pair WFAligner::align_end_to_end (src/aligner.rs:248-260, i.e.
 *                                   wavefront_aligner_set_alignment_end_to_end + wavefront_align)
 *                                   followed by WFAligner::cigar_score_clipped (src/aligner.rs:320-376)
 *   tdb_trio_reduce                 trio assess_denovo minus the alignment calls
 *                                   (src/trio/denovo.rs:83-186: get_denovo_coverage :326-362,
 *                                   math::median, get_overlap_coverage, update_mean_matrix :199-229,
 *                                   combination ranking :149-173, get_denovo_type :245-308)
 *   tdb_duo_reduce                  duo assess_denovo minus the alignment calls (src/duo/denovo.rs:46-119)
 *   tdb_trio_batch / tdb_duo_batch  trio::denovo::assess_denovo / duo::denovo::assess_denovo for many
 *                                   loci at once (alignment + fused on-device reduction)
 *   tdb_align_end_to_end, tdb_score, tdb_cigar_ops, tdb_cigar_score_clipped
 *                                   single-pair shim with WFAligner's own call shape
 *                                   (src/aligner.rs:248, :290, :438, :359)
 *   tdb_set_heuristic               WFAligner::set_heuristic (src/aligner.rs:552-605), None and
 *                                   WFadaptive only (the others are never reached from main)
 *   tdb_last_error                  AlignmentStatus / panics of src/aligner.rs:64-86
 *
 * Error model: every call returns TDB_OK (0) or a negative TDB_E_* code and records a message
 * retrievable with tdb_last_error().  Nothing aborts.  There is NO CPU fallback: without a CUDA
 * device (or with the kernels missing) tdb_create fails with TDB_E_NO_DEVICE.
 * Per-pair status uses WFA2-lib's integer convention (0 completed, <0 failed), so the host keeps
 * dropping non-completed scores exactly as src/denovo.rs:31-35 does.
 *
 * Ownership/threading: a ctx owns its device buffers and streams, is bound to one GPU and is
 * single-owner like WFAligner (`&mut self`, one per worker thread, src/commands/trio.rs:113-116).
 * Calls are synchronous from the caller's view.  One ctx per GPU may be driven concurrently from
 * different host threads.  No caller pointer is retained after a call returns.
 */
#ifndef TRGT_DENOVO_B200_H
#define TRGT_DENOVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDB_ABI_VERSION 5 /* 5: tdb_counters grew (long tier) */

/* return codes */
#define TDB_OK 0
#define TDB_E_INVALID (-1)     /* bad argument */
#define TDB_E_NO_DEVICE (-2)   /* no usable CUDA device: the product has no CPU path */
#define TDB_E_CUDA (-3)        /* CUDA runtime error (message in tdb_last_error) */
#define TDB_E_OOM (-4)         /* device or host allocation failed */
#define TDB_E_UNSUPPORTED (-5) /* configuration outside what the kernels implement */

/* per-pair alignment status = WFA2-lib WF_STATUS_* (src/aligner.rs:64-86) */
#define TDB_STATUS_ALG_COMPLETED 0
#define TDB_STATUS_ALG_PARTIAL 1
#define TDB_STATUS_MAX_STEPS_REACHED (-100)
#define TDB_STATUS_OOM (-200)
#define TDB_STATUS_UNATTAINABLE (-300)
/* out_score of a pair whose status is not 0: there is no score.  The fused calls (tdb_trio_batch, tdb_duo_batch,
 * tdb_assess_batch_device, *_packed) keep such pairs out of every score list, which is what dropping non-completed
 * alignments does in the reference (src/denovo.rs:31-35, src/allele.rs:220); tdb_trio_reduce / tdb_duo_reduce skip
 * entries holding this value for the same reason.  tdb_counters.pairs_failed counts them. */
#define TDB_SCORE_FAILED (-2147483647 - 1)

/* AlnScoring, src/model.rs:5-23 (match is fixed to 0 as in create_aligner). */
typedef struct {
    int32_t mismatch;       /* 8  */
    int32_t gap_opening1;   /* 4  */
    int32_t gap_extension1; /* 2  */
    int32_t gap_opening2;   /* 24 */
    int32_t gap_extension2; /* 1  */
} tdb_penalties;

/* Heuristic, src/aligner.rs:32-41: only None and WFadaptive are implemented. */
#define TDB_HEURISTIC_NONE 0
#define TDB_HEURISTIC_WFADAPTIVE 1
typedef struct {
    int32_t kind;
    int32_t min_wavefront_length;   /* 10 (WFA2-lib default) */
    int32_t max_distance_threshold; /* 50 */
    int32_t steps_between_cutoffs;  /* 1  */
} tdb_heuristic;

typedef struct {
    int32_t device_id;       /* CUDA device ordinal */
    tdb_penalties penalties; /* see tdb_default_config */
    tdb_heuristic heuristic;
    int32_t clip_len;          /* Params.clip_len: flank_len, or 0 with --no-clip-aln */
    int32_t reserved_flags;    /* 0 */
    uint64_t scratch_bytes;    /* cap for the long-alignment tiers' device scratch; 0 = default */
} tdb_config;

typedef struct tdb_ctx tdb_ctx;

/* Fills the reference's production configuration: penalties 8,4,2,24,1, wf-adaptive(10,50,1),
 * clip_len 50 (src/commands/shared.rs:39-63, src/cli.rs flank_len default). */
void tdb_default_config(tdb_config *cfg);
int tdb_abi_version(void);
/* Number of CUDA devices visible (0 when none; never fails). */
int tdb_device_count(void);

int tdb_create(const tdb_config *cfg, tdb_ctx **out);
void tdb_destroy(tdb_ctx *ctx);
/* Message of the last failed call on this ctx ("" if none); with ctx == NULL, of the last failed
 * tdb_create on this thread.  Valid until the next call. */
const char *tdb_last_error(const tdb_ctx *ctx);

int tdb_set_clip_len(tdb_ctx *ctx, int32_t clip_len);
/* CPU threads the host-buffer calls may use for 2-bit packing and chunk planning (default: all hardware
 * threads but one; with one ctx per GPU on a shared host give each its share, e.g. cores / n_gpus). */
int tdb_set_host_threads(tdb_ctx *ctx, int32_t n_threads);
int tdb_set_heuristic(tdb_ctx *ctx, const tdb_heuristic *h);

/*
 * Batched alignment.  Sequences are raw ASCII bytes exactly as the reference passes them
 * (`read.bases`, `allele.seq`; raw byte equality: 'N'=='N', case matters), concatenated in
 * `seq_blob`; sequence i is seq_blob[seq_off[i] .. seq_off[i]+seq_len[i]).  Layout contract (checked
 * on the device, TDB_E_INVALID otherwise): sequences are stored in index order and do not overlap
 * (seq_off[i] + seq_len[i] <= seq_off[i+1] <= blob_bytes) -- what appending reads and alleles locus
 * by locus produces.  seq_off == NULL (every batch entry point, host or device) declares the DENSE
 * layout that appending to one byte vector gives: seq_off[0] = 0, seq_off[i+1] = seq_off[i] +
 * seq_len[i]; the offsets are then derived on the device by a prefix sum of seq_len and never cross
 * PCIe (8 bytes per sequence less to upload: prefer it).  Pair j aligns
 * pattern = sequence pair_pattern[j] (the read) against text = sequence pair_text[j] (the target).
 * out_score[j]  = cigar_score_clipped(clip_len) (<= 0), written when out_status[j] == 0;
 * out_status[j] = WFA2 status.  out_status may be NULL.
 * All pointers are HOST pointers (pinned memory -- tdb_host_alloc -- lets the copies overlap the
 * kernels; pageable memory works but serialises them).  Large batches are cut into pair chunks
 * whose uploads are pipelined against the alignment of the previous chunk; this needs the pair list
 * to walk the sequence table roughly in order (locus by locus), otherwise the call still works but
 * uploads everything first.
 * Equal (pattern bytes, text bytes) pairs are aligned once and the result copied to the duplicates
 * (the alignment is a pure function of the two byte strings).
 */
int tdb_align_batch(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off,
                    const uint32_t *seq_len, size_t n_seqs, const uint32_t *pair_pattern,
                    const uint32_t *pair_text, size_t n_pairs, int32_t *out_score, int32_t *out_status);

/* Same, with every pointer a DEVICE pointer on the ctx's GPU (inputs already resident in HBM).
 * seq_blob must be readable 16 bytes past blob_bytes.  Synchronous. */
int tdb_align_batch_device(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes,
                           const uint64_t *seq_off, const uint32_t *seq_len, size_t n_seqs,
                           const uint32_t *pair_pattern, const uint32_t *pair_text, size_t n_pairs,
                           int32_t *out_score, int32_t *out_status);

/* As tdb_align_batch, additionally returning every alignment's CIGAR (one ASCII byte per column,
 * 'M','X','I','D', forward order = wfa2 cigar_t.operations[begin_offset..end_offset),
 * src/aligner.rs:438-448).  Pair j's columns go to cigar_buf[cigar_off[j] ..); the caller sizes
 * each slot >= plen+tlen; cigar_len[j] receives the column count.  Host pointers.  Diagnostic /
 * test entry point: it runs every pair through the general kernel tier. */
int tdb_align_batch_cigar(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes,
                          const uint64_t *seq_off, const uint32_t *seq_len, size_t n_seqs,
                          const uint32_t *pair_pattern, const uint32_t *pair_text, size_t n_pairs,
                          int32_t *out_score, int32_t *out_status, int32_t *out_penalty,
                          uint8_t *cigar_buf, const uint64_t *cigar_off, uint32_t *cigar_len);

/*
 * Per-locus reduction descriptors.  For child (duo: sample A) allele c < n_child the scores of
 *   j=0: mother (duo: sample B) allele 0 reads vs child allele c     (align_alleleset, denovo.rs:22)
 *   j=1: mother allele 1 reads          j=2: father allele 0 reads   j=3: father allele 1 reads
 *   j=4: the child's own allele-c reads vs child allele c            (align_allele, denovo.rs:41)
 * are scores[list_off[c][j] .. list_off[c][j+1]).  Lists of absent alleles are empty.
 * Only pairs with status 0 may be listed (src/denovo.rs:31-35 drops the others).
 */
typedef struct {
    int32_t n_child, n_mother, n_father; /* allele counts, 1 or 2 (duo: n_father = 0) */
    uint32_t list_off[2][6];
    int32_t child_len[2], mother_len[2], father_len[2]; /* allele sequence lengths (get_denovo_type) */
} tdb_trio_locus;

/* allele_origin codes (src/model.rs:87-138 Display): 0 "M:1", 1 "M:2", 2 "F:1", 3 "F:2", 4 "?", 5 "." */
/* denovo_status codes (src/model.rs:46-83): 0 "X", 1 "Y:+", 2 "Y:-", 3 "Y:=", 4 "Y:?", 5 "." */
typedef struct {
    int32_t denovo_coverage;
    float mean_diff_father; /* duo: unused */
    float mean_diff_mother; /* duo: mean_diff_b */
    int32_t mother_overlap[2]; /* duo: b_overlap_coverage */
    int32_t father_overlap[2];
    int32_t allele_origin;
    int32_t denovo_status;
    int32_t error; /* 1: all parental score lists empty -- the reference panics here
                      (Option::unwrap, src/trio/denovo.rs:333-334, src/duo/denovo.rs:107) */
    double top_mother, top_father, child_median;
    double mean_col[4]; /* this allele's column of the 4x2 mean matrix (trio/denovo.rs:199-229) */
} tdb_allele_out;

/* Reductions on scores that are already on the host.  out has 2 slots per locus (slot c of locus l
 * at out[2*l+c]); denovo_mask[i] (may be NULL) = 1 where own-read score i exceeds the threshold
 * (the read ids of src/trio/denovo.rs:339-343). */
int tdb_trio_reduce(tdb_ctx *ctx, const tdb_trio_locus *loci, size_t n_loci, const int32_t *scores,
                    size_t n_scores, double parent_quantile, tdb_allele_out *out, uint8_t *denovo_mask);
int tdb_duo_reduce(tdb_ctx *ctx, const tdb_trio_locus *loci, size_t n_loci, const int32_t *scores,
                   size_t n_scores, double parent_quantile, tdb_allele_out *out, uint8_t *denovo_mask);

/* Fused alignment + reduction for a batch of loci: what assess_denovo does, for thousands of loci
 * per call.  scores index space == pair index space (list_off indexes pairs).  out_score and
 * denovo_mask may be NULL.  Host pointers. */
int tdb_trio_batch(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off,
                   const uint32_t *seq_len, size_t n_seqs, const uint32_t *pair_pattern,
                   const uint32_t *pair_text, size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci,
                   double parent_quantile, tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask);
int tdb_duo_batch(tdb_ctx *ctx, const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off,
                  const uint32_t *seq_len, size_t n_seqs, const uint32_t *pair_pattern,
                  const uint32_t *pair_text, size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci,
                  double parent_quantile, tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask);
/* Device-resident variants of the fused call (all pointers device pointers; is_duo selects mode). */
int tdb_assess_batch_device(tdb_ctx *ctx, int32_t is_duo, const uint8_t *seq_blob, size_t blob_bytes,
                            const uint64_t *seq_off, const uint32_t *seq_len, size_t n_seqs,
                            const uint32_t *pair_pattern, const uint32_t *pair_text, size_t n_pairs,
                            const tdb_trio_locus *loci, size_t n_loci, double parent_quantile,
                            tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask);

/*
 * Packed input: the caller packed at load time (north_star: "locus.rs and read.rs pack the reads and alleles ...
 * into 2-bit/byte-packed device batches"; the reference expands BAM's 4-bit bases to ASCII at src/read.rs:57 and
 * builds allele targets at src/allele.rs:406-426 -- a host that emits 2-bit codes there never materialises ASCII
 * and the batch call does no per-base work on the CPU).  The layout is the one tdb_host_pack produces:
 *   packed    ONE 2-bit stream over all sequences: base at stream position x -> word x / 16, bits 2 (x % 16),
 *             code (ASCII byte >> 1) & 3  (A 0, C 1, T 2, G 3); tdb_packed_words(n_bases) words (the last two are
 *             padding and must be readable; their content is ignored)
 *   seq_off   stream position of base 0 of each sequence, or NULL for the dense layout (back to back)
 *   seq_len   bases per sequence (or seq_len16, see below)
 *   exc_seq   ascending indices of the sequences that hold a byte outside {A,C,G,T} ('N', IUPAC codes, lower
 *             case): WFA2 compares raw bytes, so those are compared byte-wise from exc_bytes, where their raw
 *             ASCII is stored back to back in exc_seq order (seq_len[exc_seq[i]] bytes each).  Their 2-bit codes
 *             in `packed` are still (byte >> 1) & 3 of every byte.  n_exc == 0: all three may be NULL.
 * Host pointers; pinned memory (tdb_host_alloc) lets the upload of chunk k+1 overlap the alignment of chunk k.
 * Results are bit-identical to the ASCII entry points on the same sequences.
 */
typedef struct {
    const uint32_t *packed;
    size_t n_bases;
    const uint64_t *seq_off;
    const uint32_t *seq_len;
    size_t n_seqs;
    const uint32_t *exc_seq;
    const uint8_t *exc_bytes;
    size_t n_exc;
    const uint16_t *seq_len16; /* alternative to seq_len (which must then be NULL) when every sequence is shorter than
                                  65 536 bases: the length table that crosses PCIe is half as large */
} tdb_packed_seqs;
int tdb_align_batch_packed(tdb_ctx *ctx, const tdb_packed_seqs *seqs, const uint32_t *pair_pattern,
                           const uint32_t *pair_text, size_t n_pairs, int32_t *out_score, int32_t *out_status);
int tdb_trio_batch_packed(tdb_ctx *ctx, const tdb_packed_seqs *seqs, const uint32_t *pair_pattern,
                          const uint32_t *pair_text, size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci,
                          double parent_quantile, tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask);
int tdb_duo_batch_packed(tdb_ctx *ctx, const tdb_packed_seqs *seqs, const uint32_t *pair_pattern,
                         const uint32_t *pair_text, size_t n_pairs, const tdb_trio_locus *loci, size_t n_loci,
                         double parent_quantile, tdb_allele_out *out, int32_t *out_score, uint8_t *denovo_mask);

/* Single-pair shim keeping WFAligner's call shape (src/aligner.rs:248-260, :290, :438, :359). */
int tdb_align_end_to_end(tdb_ctx *ctx, const uint8_t *pattern, int32_t pattern_len, const uint8_t *text,
                         int32_t text_len); /* returns the WFA2 status, or a TDB_E_* (< -1000 never) */
int32_t tdb_score(const tdb_ctx *ctx);                          /* cigar.score = -penalty */
const uint8_t *tdb_cigar_ops(const tdb_ctx *ctx, int32_t *len); /* host copy, valid until next align */
int32_t tdb_cigar_score_clipped(const tdb_ctx *ctx, int32_t flank_len);

/* Counters of the last batch call (for roofline accounting and tier diagnostics). */
typedef struct {
    uint64_t pairs;         /* pairs in the call */
    uint64_t tier_pairs[4]; /* pairs ALIGNED by tier 0 (four pairs per warp, up to 32 diagonals, shared memory), 1 (four
                               pairs per warp, up to 64 diagonals; other penalty sets: one pair per warp), 2 and 3 (one
                               pair per warp, global scratch).  See long_pairs / thread_pairs for the other two tiers. */
    uint64_t cells;         /* C: sum over computed score steps of wavefront width, ALGORITHMIC: */
    uint64_t ext16;         /* E: 16-base extension compares         every pair counted, duplicates included */
    uint64_t columns;       /* T: CIGAR columns replayed             (= what the CPU oracle counts) */
    uint64_t exec_cells, exec_ext16, exec_columns; /* the same for the work the kernels actually executed */
    uint64_t tier_cells[4], tier_ext16[4], tier_columns[4]; /* executed work per tier */
    uint64_t pairs_identical; /* pattern bytes == text bytes: resolved without an alignment */
    uint64_t pairs_duplicate; /* copies of an already aligned (pattern, text) content */
    uint64_t pairs_closed_form; /* one substitution / one gap of <= 20 bases: scored in closed form (0 while
                                   exact work counting is on) */
    uint64_t kernel_launches;
    uint64_t chunks;        /* pipeline chunks the call was cut into */
    uint64_t host_packed;   /* 1: the caller's CPU threads 2-bit packed the sequences before the upload;
                               2: the caller handed over packed sequences (*_packed entry points) */
    uint64_t bytes_h2d;     /* bytes the alignment stage actually copied host -> device (host-buffer calls) */
    uint64_t bytes_raw;     /* ... of which ASCII blocks the calling thread took over from the host packer while the
                               copy engine was idle (packed on the GPU instead) */
    float ms_pack, ms_dedup, ms_align, ms_scatter, ms_reduce, ms_total_device; /* CUDA-event times */
    float ms_host_plan, ms_host_pack_wait, ms_host_call; /* host wall clock: chunk planning, time the calling
                         thread waited for the packer threads, whole run of pack+align (host-buffer calls) */
    float ms_host_packed, ms_host_enqueued; /* since the start of the call: last block packed by the worker threads,
                         last chunk queued by the calling thread */
    float ms_tier[4]; /* per alignment tier (events on the launching stream); [2] covers both global tiers.
                         Stage times are recorded for single-chunk calls only (ms_total_device always). */
    float ms_thread_tier;   /* the thread-per-pair tier that runs in front of tier 0 */
    uint64_t thread_pairs;  /* pairs ALIGNED by the thread-per-pair tier (final penalty <= 24), and the work */
    uint64_t thread_cells, thread_ext16, thread_columns; /* it executed (also included in exec_*) */
    uint64_t bytes_d2h;     /* result bytes copied device -> host by the fused host-buffer calls (tdb_trio_batch /
                               tdb_duo_batch): rows, scores if asked for, and the read mask -- as a list of its set
                               positions for calls of >= 8 Mi pairs, where a few bytes in a million are set */
    uint64_t pairs_failed;  /* pairs no tier could finish (status TDB_STATUS_OOM, score TDB_SCORE_FAILED) */
    float ms_fused_call;    /* host wall clock of the whole fused host-buffer call (tdb_trio_batch[_packed] ...), entry to return */
    float ms_fused_tail;    /* ... of which after the alignment stage: reduction, row and mask download */
    uint64_t long_pairs;    /* pairs ALIGNED by the long tier (one pair per warp, wavefronts of up to 192 diagonals in shared
                               memory, provenance in global scratch) that runs in front of tiers 2 and 3, and the work */
    uint64_t long_cells, long_ext16, long_columns; /* it executed (also included in exec_*) */
    float ms_long_tier;     /* its share of ms_tier[2] */
    float reserved0;
} tdb_counters;
int tdb_get_counters(const tdb_ctx *ctx, tdb_counters *out);
/* Exact algorithmic work counting (cells / ext16 / columns of tdb_counters: every pair counted as the CPU
 * oracle counts it, duplicates included).  Off by default: it routes every non-identical content through
 * the wavefront kernels (no closed form) and keeps per-pair work records.  The exec_* / tier_* counters of
 * the work the kernels executed are always maintained. */
int tdb_set_count_work(tdb_ctx *ctx, int32_t on);

/* The host half of the packing role (north_star: "locus.rs and read.rs pack the reads and alleles ...
 * into 2-bit packed device batches"): ASCII blob -> 2-bit stream, exactly what the host-buffer batch
 * calls run on the caller's CPU threads before the upload.  Data re-encoding only, exposed so that a
 * host can inspect / pre-compute the packed form; needs no GPU.  Layout: the byte at blob position x
 * goes to word x / 16, bits 2 (x % 16), code (byte >> 1) & 3 -- independent of sequence boundaries; a
 * sequence is the window [seq_off, seq_off + seq_len) of the stream.  packed must hold
 * tdb_packed_words(blob_bytes) words; flag[i] = 1 when sequence i has a byte outside {A,C,G,T}.
 * Returns TDB_OK or TDB_E_INVALID (layout contract violated). */
size_t tdb_packed_words(size_t blob_bytes);
int tdb_host_pack(const uint8_t *seq_blob, size_t blob_bytes, const uint64_t *seq_off, const uint32_t *seq_len,
                  size_t n_seqs, uint32_t *packed, uint8_t *flag, int n_threads);
/* "avx512bw", "avx2" or "scalar": the instruction set the host packer uses on this CPU. */
const char *tdb_host_pack_isa(void);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost) for callers without a CUDA binding. */
void *tdb_host_alloc(size_t bytes);
void tdb_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* TRGT_DENOVO_B200_H */
