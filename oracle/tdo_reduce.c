/*
 * tdo_reduce.c -- ORACLE (test infrastructure only, see tdo_oracle.h).
 *
 * Line-by-line restatement of the reference's score reductions:
 *   src/math.rs:82-98    quantile            src/math.rs:100-155  partition/select/median
 *   src/denovo.rs:68-78  get_overlap_coverage
 *   src/denovo.rs:93-101 get_top_other_score src/denovo.rs:117-129 get_score_count_diff
 *   src/trio/denovo.rs:83-186 assess_denovo, :199-229 update_mean_matrix, :245-308 denovo type,
 *   src/trio/denovo.rs:326-362 get_denovo_coverage
 *   src/duo/denovo.rs:46-119
 * Pinned on the reference unit tests (src/trio/denovo.rs:369-475,592-612; src/math.rs:157-201).
 */
#include "tdo_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int cmp_f64(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* src/math.rs:82-98 */
int tdo_quantile(double *xs, int n, double q, double *out) {
    if (n == 0) return -1;
    qsort(xs, (size_t)n, sizeof(double), cmp_f64);
    double index = q * ((double)n - 1.0);
    double fl = floor(index);
    size_t lower = fl <= 0.0 ? 0 : (size_t)fl; /* Rust `as usize` saturates */
    size_t upper = lower + 1;
    double fraction = index - (double)lower;
    if (upper >= (size_t)n) {
        *out = xs[n - 1];
        return 0;
    }
    *out = xs[lower] + (xs[upper] - xs[lower]) * fraction;
    return 0;
}

/* src/math.rs:100-135 : quick-select through partition(); the value selected is the k-th order
 * statistic, so a direct restatement of partition/select is kept to mirror tie handling. */
static int select_k(const int32_t *data, int n, int k, int32_t *out) {
    if (n == 0) return -1;
    int32_t pivot = data[0];
    int32_t *left = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *right = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int nl = 0, nr = 0;
    for (int i = 1; i < n; ++i) {
        if (data[i] < pivot) left[nl++] = data[i];
        else right[nr++] = data[i];
    }
    int r;
    if (nl == k) {
        *out = pivot;
        r = 0;
    } else if (nl > k) r = select_k(left, nl, k, out);
    else r = select_k(right, nr, k - (nl + 1), out);
    free(left);
    free(right);
    return r;
}

/* src/math.rs:137-155 */
int tdo_median(const int32_t *data, int n, double *out) {
    if (n == 0) return -1;
    if (n % 2 == 0) {
        int32_t a, b;
        if (select_k(data, n, n / 2 - 1, &a) || select_k(data, n, n / 2, &b)) return -1;
        *out = (double)(int32_t)((uint32_t)a + (uint32_t)b) / 2.0; /* i32 add (wrapping in release) */
        return 0;
    }
    int32_t a;
    if (select_k(data, n, n / 2, &a)) return -1;
    *out = (double)a;
    return 0;
}

/* src/denovo.rs:93-101 */
int tdo_get_top_other_score(const int32_t *scores, const int32_t *list_off, int n_lists, double q,
                            double *out) {
    int have = 0;
    double best = 0.0;
    for (int l = 0; l < n_lists; ++l) {
        int n = list_off[l + 1] - list_off[l];
        if (n <= 0) continue;
        double *xs = (double *)malloc(sizeof(double) * (size_t)n);
        for (int i = 0; i < n; ++i) xs[i] = (double)scores[list_off[l] + i];
        double v;
        tdo_quantile(xs, n, q, &v);
        free(xs);
        /* Iterator::max_by keeps the last maximum; values are equal then */
        if (!have || v >= best) best = v, have = 1;
    }
    if (!have) return -1;
    *out = best;
    return 0;
}

/* src/denovo.rs:117-129 */
void tdo_get_score_count_diff(double top, const int32_t *aligns, int n, int64_t *count, float *mean_diff) {
    int64_t c = 0;
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        double a = (double)aligns[i];
        if (a > top) ++c, sum += a;
    }
    *count = c;
    *mean_diff = c > 0 ? (float)fabs(sum / (double)c - top) : 0.0f;
}

/* src/denovo.rs:68-78 */
int32_t tdo_get_overlap_coverage(double thr, const int32_t *aligns, int n) {
    int32_t c = 0;
    for (int i = 0; i < n; ++i)
        if ((double)aligns[i] >= thr) ++c;
    return c;
}

static double list_mean(const int32_t *s, uint32_t b, uint32_t e) {
    if (e <= b) return -INFINITY; /* src/trio/denovo.rs:214-216 */
    int32_t sum = 0;
    for (uint32_t i = b; i < e; ++i) sum = (int32_t)((uint32_t)sum + (uint32_t)s[i]); /* i32 sum */
    return (double)sum / (double)(e - b);
}

static int top_of(const int32_t *scores, const uint32_t *off, int first, int n_lists, double q, double *out) {
    int32_t lo[3];
    /* repack the (possibly 2) lists into a local offset table relative to scores */
    for (int l = 0; l <= n_lists; ++l) lo[l] = (int32_t)off[first + l];
    return tdo_get_top_other_score(scores, lo, n_lists, q, out);
}

static const int COMBS_1D[4][2][2] = {{{0, 0}, {0, 0}}, {{1, 0}, {1, 0}}, {{2, 0}, {2, 0}}, {{3, 0}, {3, 0}}};
static const int COMBS_2D[8][2][2] = {{{0, 0}, {2, 1}}, {{0, 0}, {3, 1}}, {{1, 0}, {2, 1}}, {{1, 0}, {3, 1}},
                                      {{2, 0}, {0, 1}}, {{3, 0}, {0, 1}}, {{2, 0}, {1, 1}}, {{3, 0}, {1, 1}}};

/* src/trio/denovo.rs:83-186 */
void tdo_trio_assess(const tdo_trio_locus *loc, const int32_t *scores, double q, tdo_allele_out *out,
                     uint8_t *denovo_mask) {
    double matrix[4][2];
    for (int i = 0; i < 4; ++i) matrix[i][0] = matrix[i][1] = -DBL_MAX; /* f64::MIN */
    for (int c = 0; c < loc->n_child; ++c) {
        const uint32_t *off = loc->list_off[c];
        tdo_allele_out *o = &out[c];
        memset(o, 0, sizeof(*o));
        o->allele_origin = 4;
        const int32_t *child = scores + off[4];
        int nchild = (int)(off[5] - off[4]);
        /* get_denovo_coverage, src/trio/denovo.rs:326-362 */
        double top_m = 0, top_f = 0;
        int em = top_of(scores, off, 0, loc->n_mother, q, &top_m);
        /* father lists start at slot 2 regardless of n_mother (empty slot 1 when n_mother==1) */
        int ef = top_of(scores, off, 2, loc->n_father, q, &top_f);
        if (em || ef) {
            o->error = 1; /* reference: Option::unwrap() on None panics */
            continue;
        }
        o->top_mother = top_m, o->top_father = top_f;
        double thr = top_m > top_f ? top_m : top_f; /* f64::max */
        if (denovo_mask)
            for (int i = 0; i < nchild; ++i) denovo_mask[off[4] + i] = ((double)child[i] > thr) ? 1 : 0;
        int64_t cm, cf;
        tdo_get_score_count_diff(top_m, child, nchild, &cm, &o->mean_diff_mother);
        tdo_get_score_count_diff(top_f, child, nchild, &cf, &o->mean_diff_father);
        o->denovo_coverage = (int32_t)(top_m >= top_f ? cm : cf);
        double med;
        if (tdo_median(child, nchild, &med)) med = DBL_MAX; /* unwrap_or(f64::MAX) */
        o->child_median = med;
        for (int a = 0; a < loc->n_mother; ++a)
            o->mother_overlap[a] = tdo_get_overlap_coverage(med, scores + off[a], (int)(off[a + 1] - off[a]));
        for (int a = 0; a < loc->n_father; ++a)
            o->father_overlap[a] =
                tdo_get_overlap_coverage(med, scores + off[2 + a], (int)(off[3 + a] - off[2 + a]));
        /* update_mean_matrix, src/trio/denovo.rs:199-229 */
        matrix[0][c] = list_mean(scores, off[0], off[1]);
        matrix[1][c] = loc->n_mother > 1 ? list_mean(scores, off[1], off[2]) : -DBL_MAX;
        matrix[2][c] = list_mean(scores, off[2], off[3]);
        matrix[3][c] = loc->n_father > 1 ? list_mean(scores, off[3], off[4]) : -DBL_MAX;
        for (int i = 0; i < 4; ++i) o->mean_col[i] = matrix[i][c];
    }
    for (int c = 0; c < loc->n_child; ++c)
        if (out[c].error) return;
    /* combos, src/trio/denovo.rs:149-173 */
    double cs[8];
    int idx[8], ncomb;
    if (loc->n_child > 1) {
        ncomb = 8;
        for (int i = 0; i < 8; ++i) {
            double sum = 0.0; /* Iterator::sum::<f64>() starts from 0.0 (well, -0.0 + x == x anyway) */
            for (int j = 0; j < 2; ++j) sum += matrix[COMBS_2D[i][j][0]][COMBS_2D[i][j][1]];
            cs[i] = sum, idx[i] = i;
        }
    } else {
        ncomb = 4;
        for (int i = 0; i < 4; ++i) cs[i] = matrix[COMBS_1D[i][0][0]][COMBS_1D[i][0][1]], idx[i] = i;
    }
    /* sort_unstable_by(desc) on <= 20 elements is an insertion sort in Rust's std: stable. */
    for (int i = 1; i < ncomb; ++i) {
        double v = cs[i];
        int vi = idx[i], j = i - 1;
        while (j >= 0 && cs[j] < v) cs[j + 1] = cs[j], idx[j + 1] = idx[j], --j;
        cs[j + 1] = v, idx[j + 1] = vi;
    }
    double comb_diff = fabs(cs[0] - cs[1]);
    for (int c = 0; c < loc->n_child; ++c) {
        tdo_allele_out *o = &out[c];
        if (comb_diff < 1.0) o->allele_origin = 4;
        else o->allele_origin = loc->n_child > 1 ? COMBS_2D[idx[0]][c][0] : COMBS_1D[idx[0]][c][0];
        if (o->denovo_coverage == 0) o->denovo_status = 0;
        else {
            /* get_denovo_type, src/trio/denovo.rs:245-308 */
            int plen = -1;
            switch (o->allele_origin) {
            case 0: plen = loc->mother_len[0]; break;
            case 1: plen = loc->mother_len[1]; break;
            case 2: plen = loc->father_len[0]; break;
            case 3: plen = loc->father_len[1]; break;
            default: break;
            }
            if (plen < 0) o->denovo_status = 4;
            else if (plen < loc->child_len[c]) o->denovo_status = 1;
            else if (plen > loc->child_len[c]) o->denovo_status = 2;
            else o->denovo_status = 3;
        }
    }
}

/* src/duo/denovo.rs:46-119 */
void tdo_duo_assess(const tdo_trio_locus *loc, const int32_t *scores, double q, tdo_allele_out *out,
                    uint8_t *denovo_mask) {
    for (int c = 0; c < loc->n_child; ++c) {
        const uint32_t *off = loc->list_off[c];
        tdo_allele_out *o = &out[c];
        memset(o, 0, sizeof(*o));
        o->allele_origin = 5;
        const int32_t *own = scores + off[4];
        int nown = (int)(off[5] - off[4]);
        double top_b = 0;
        if (top_of(scores, off, 0, loc->n_mother, q, &top_b)) {
            o->error = 1;
            continue;
        }
        o->top_mother = top_b;
        if (denovo_mask)
            for (int i = 0; i < nown; ++i) denovo_mask[off[4] + i] = ((double)own[i] > top_b) ? 1 : 0;
        int64_t cnt;
        tdo_get_score_count_diff(top_b, own, nown, &cnt, &o->mean_diff_mother);
        o->denovo_coverage = (int32_t)cnt;
        double med;
        if (tdo_median(own, nown, &med)) med = DBL_MAX;
        o->child_median = med;
        for (int a = 0; a < loc->n_mother; ++a)
            o->mother_overlap[a] = tdo_get_overlap_coverage(med, scores + off[a], (int)(off[a + 1] - off[a]));
        o->denovo_status = cnt == 0 ? 0 : 4;
    }
}
