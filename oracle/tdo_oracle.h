/*
 * tdo_oracle.h -- CPU ORACLE for the trgt-denovo read<->allele alignment hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (trgt-denovo_b200/csrc) never links or calls anything in oracle/.
 *
 * What it restates (plain C, scalar, one aligner state per thread):
 *   - WFA2-lib's gap-affine-2-piece, end-to-end, alignment-scope wavefront aligner with the
 *     library-default wf-adaptive(10,50,1) heuristic, as reached from the reference through
 *     src/aligner.rs:248-260 (align_end_to_end) and src/commands/shared.rs:41-51.
 *     WFA2-lib itself is NOT under /root/reference: it is the un-vendored dependency
 *     wfa2-sys 0.1.0 (git https://github.com/ctsa/rust-wfa2.git rev 4342b3b0..., Cargo.toml:39,
 *     Cargo.lock:1543-1550).  Its published algorithm (WFA2-lib 2.3.x: wavefront_compute_affine2p,
 *     wavefront_extend_end2end, wavefront_heuristic_wfadaptive, piggy-back backtrace + pcigar unpack)
 *     is restated from SURVEY.md Appendix A.
 *   - src/aligner.rs:303-357  cigar_score_gap_affine2_clipped
 *   - src/denovo.rs:68-129, src/math.rs:82-155, src/trio/denovo.rs:83-362, src/duo/denovo.rs:46-119
 *
 * PARITY PINNING: pinned on the reference's own known-answer vectors (src/aligner.rs tests
 * B4,B5,B6,B7,B9 of SURVEY.md Appendix B; reduction tests of src/trio/denovo.rs and src/math.rs).
 * The wf-adaptive cut-off in MemoryLow mode is NOT asserted by any reference test: for pairs whose
 * path is changed by the heuristic the oracle is "parity unpinned" (restates Appendix A.3).
 */
#ifndef TDO_ORACLE_H
#define TDO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* WFA2-lib status codes (src/aligner.rs:64-86 maps them to AlignmentStatus). */
#define TDO_STATUS_ALG_COMPLETED 0
#define TDO_STATUS_ALG_PARTIAL 1
#define TDO_STATUS_MAX_STEPS_REACHED (-100)
#define TDO_STATUS_OOM (-200)
#define TDO_STATUS_UNATTAINABLE (-300)

/* Penalties, src/model.rs:5-23 (defaults 8,4,2,24,1).  o2 < 0 selects single-piece gap-affine
 * (used only to pin the restatement on src/aligner.rs:820-838). */
typedef struct {
    int32_t x, o1, e1, o2, e2;
} tdo_penalties;

/* kind: 0 = none, 1 = wf-adaptive(min_wf_len, max_dist, steps).  WFA2 default = {1,10,50,1}. */
typedef struct {
    int32_t kind, min_wf_len, max_dist, steps;
} tdo_heuristic;

/* Work counters, SURVEY.md section 8(d): C (cells), E (16-base extension compares), T (columns). */
typedef struct {
    int64_t cells;     /* sum over computed score steps of (hi-lo+1) */
    int64_t ext16;     /* sum over extended diagonals of (matched/16 + 1) */
    int64_t columns;   /* CIGAR columns emitted */
    int32_t max_width; /* widest M wavefront seen */
    int32_t steps;     /* final penalty score */
} tdo_stats;

typedef struct tdo_aligner tdo_aligner;

tdo_aligner *tdo_aligner_new(const tdo_penalties *pen, const tdo_heuristic *heur);
void tdo_aligner_delete(tdo_aligner *a);
void tdo_aligner_set_heuristic(tdo_aligner *a, const tdo_heuristic *heur);

/* src/aligner.rs:248-260.  pattern = read, text = target.  Returns a TDO_STATUS_*. */
int tdo_align_end_to_end(tdo_aligner *a, const uint8_t *pattern, int plen, const uint8_t *text, int tlen);
/* src/aligner.rs:290 : cigar.score (= -penalty for match=0). */
int tdo_score(const tdo_aligner *a);
/* One ASCII byte per column ('M','X','I','D'), forward order; src/aligner.rs:438-448. */
const char *tdo_cigar_ops(const tdo_aligner *a, int *len);
/* src/aligner.rs:404-436 : run-length string of [begin+clip, end-clip). Returns length written. */
int tdo_cigar_string(const tdo_aligner *a, int clip, char *buf, int buflen);
/* src/aligner.rs:320-357 / 359-376. */
int tdo_cigar_score_clipped(const tdo_aligner *a, int clip);
void tdo_get_stats(const tdo_aligner *a, tdo_stats *st);

/* Independent check of the optimal score: plain O(n*m) Gotoh DP with two affine pieces.
 * Returns the optimal penalty (>= 0). */
int64_t tdo_gotoh2p_penalty(const uint8_t *pattern, int plen, const uint8_t *text, int tlen,
                            const tdo_penalties *pen);

/* Batched, multi-threaded: pairs index into a sequence blob.  One aligner per thread, mirroring
 * the thread-local aligner of src/commands/trio.rs:113-116.  out_stats may be NULL.
 * Semantics per pair = align_end_to_end + cigar_score_clipped(clip) (src/denovo.rs:29-37). */
int tdo_align_batch(const uint8_t *blob, const uint64_t *seq_off, const uint32_t *seq_len,
                    const uint32_t *pair_pattern, const uint32_t *pair_text, size_t n_pairs,
                    const tdo_penalties *pen, const tdo_heuristic *heur, int clip, int n_threads,
                    int32_t *out_score, int32_t *out_status, tdo_stats *out_total);

/* ---- reductions ------------------------------------------------------------------------- */

/* src/math.rs:82-98.  Sorts xs in place.  Returns 0 and *out on success, -1 when empty. */
int tdo_quantile(double *xs, int n, double q, double *out);
/* src/math.rs:137-155.  Returns 0 / -1 (empty). */
int tdo_median(const int32_t *data, int n, double *out);
/* src/denovo.rs:93-101.  lists given as concatenated scores + n_lists+1 offsets. */
int tdo_get_top_other_score(const int32_t *scores, const int32_t *list_off, int n_lists, double q,
                            double *out);
/* src/denovo.rs:117-129. */
void tdo_get_score_count_diff(double top, const int32_t *aligns, int n, int64_t *count, float *mean_diff);
/* src/denovo.rs:68-78 for one list. */
int32_t tdo_get_overlap_coverage(double thr, const int32_t *aligns, int n);

/* One child (or sample-A) allele worth of reduced output. */
typedef struct {
    int32_t denovo_coverage;
    float mean_diff_father; /* duo: mean_diff_b */
    float mean_diff_mother;
    int32_t mother_overlap[2]; /* duo: b_overlap */
    int32_t father_overlap[2];
    int32_t allele_origin; /* 0..3 = M:1,M:2,F:1,F:2 (src/model.rs:94-112), 4 = '?', 5 = '.' */
    int32_t denovo_status; /* 0 = X, 1 = Y:+, 2 = Y:-, 3 = Y:=, 4 = Y:?, 5 = '.' */
    int32_t error;         /* 1 = reference would panic (all parental lists empty, trio/denovo.rs:333-334) */
    double top_mother, top_father, child_median;
    double mean_col[4]; /* column of the 4x2 mean matrix, src/trio/denovo.rs:199-229 */
} tdo_allele_out;

/* Locus descriptor shared with the product ABI (same field meaning as tdb_trio_locus). */
typedef struct {
    int32_t n_child, n_mother, n_father; /* allele counts (1 or 2) */
    /* For child allele c: scores of lists mother a0, mother a1, father a0, father a1, own reads are
     * scores[list_off[c][j] .. list_off[c][j+1]) for j = 0..4. */
    uint32_t list_off[2][6];
    int32_t child_len[2], mother_len[2], father_len[2]; /* allele sequence lengths (denovo type) */
} tdo_trio_locus;

/* src/trio/denovo.rs:83-186 on precomputed scores.  out[c] for c < n_child;
 * denovo_mask (may be NULL) gets 1 for every own-read score index > threshold (read ids). */
void tdo_trio_assess(const tdo_trio_locus *loc, const int32_t *scores, double q, tdo_allele_out *out,
                     uint8_t *denovo_mask);
/* src/duo/denovo.rs:46-119 : father_* fields unused, lists j=0,1 = B a0,a1 ; j=4 = own. */
void tdo_duo_assess(const tdo_trio_locus *loc, const int32_t *scores, double q, tdo_allele_out *out,
                    uint8_t *denovo_mask);

/* Whole-batch reductions threaded over loci; out has 2 slots per locus. */
int tdo_assess_batch(const tdo_trio_locus *loci, size_t n_loci, const int32_t *scores, double q, int duo,
                     int n_threads, tdo_allele_out *out, uint8_t *mask);

#ifdef __cplusplus
}
#endif
#endif
