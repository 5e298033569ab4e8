/*
 * tdo_batch.c -- ORACLE (test infrastructure only, see tdo_oracle.h).
 *
 * Multi-threaded batch driver over tdo_align_end_to_end + tdo_cigar_score_clipped, i.e. the body of
 * src/denovo.rs:22-54 (align_alleleset / align_allele) for a list of (read, target) pairs.
 * One aligner per thread, as the reference's thread_local ALIGNER (src/commands/trio.rs:113-116).
 * Used as the checker in tests and as bench.py's cpu_baseline ("port": it is a scalar restatement,
 * not WFA2-lib, which ships vectorised kernels).
 */
#include "tdo_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint8_t *blob;
    const uint64_t *seq_off;
    const uint32_t *seq_len;
    const uint32_t *pp, *pt;
    size_t n_pairs;
    const tdo_penalties *pen;
    const tdo_heuristic *heur;
    int clip;
    int32_t *out_score, *out_status;
    size_t *next; /* shared work counter */
    pthread_mutex_t *mu;
    tdo_stats total;
} job_t;

#define CHUNK 256

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    tdo_aligner *a = tdo_aligner_new(j->pen, j->heur);
    memset(&j->total, 0, sizeof(j->total));
    for (;;) {
        size_t b = __atomic_fetch_add(j->next, (size_t)CHUNK, __ATOMIC_RELAXED);
        if (b >= j->n_pairs) break;
        size_t e = b + CHUNK < j->n_pairs ? b + CHUNK : j->n_pairs;
        for (size_t i = b; i < e; ++i) {
            uint32_t p = j->pp[i], t = j->pt[i];
            int st = tdo_align_end_to_end(a, j->blob + j->seq_off[p], (int)j->seq_len[p],
                                          j->blob + j->seq_off[t], (int)j->seq_len[t]);
            if (j->out_status) j->out_status[i] = st;
            j->out_score[i] = st == TDO_STATUS_ALG_COMPLETED ? tdo_cigar_score_clipped(a, j->clip) : 0;
            tdo_stats s;
            tdo_get_stats(a, &s);
            j->total.cells += s.cells;
            j->total.ext16 += s.ext16;
            j->total.columns += s.columns;
            j->total.steps += 0;
            if (s.max_width > j->total.max_width) j->total.max_width = s.max_width;
        }
    }
    tdo_aligner_delete(a);
    return NULL;
}

int tdo_align_batch(const uint8_t *blob, const uint64_t *seq_off, const uint32_t *seq_len,
                    const uint32_t *pair_pattern, const uint32_t *pair_text, size_t n_pairs,
                    const tdo_penalties *pen, const tdo_heuristic *heur, int clip, int n_threads,
                    int32_t *out_score, int32_t *out_status, tdo_stats *out_total) {
    if (n_threads < 1) n_threads = 1;
    size_t next = 0;
    job_t *jobs = (job_t *)calloc((size_t)n_threads, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int i = 0; i < n_threads; ++i) {
        job_t *j = &jobs[i];
        j->blob = blob, j->seq_off = seq_off, j->seq_len = seq_len;
        j->pp = pair_pattern, j->pt = pair_text, j->n_pairs = n_pairs;
        j->pen = pen, j->heur = heur, j->clip = clip;
        j->out_score = out_score, j->out_status = out_status, j->next = &next;
    }
    if (n_threads == 1) worker(&jobs[0]);
    else {
        for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, worker, &jobs[i]);
        for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    }
    if (out_total) {
        memset(out_total, 0, sizeof(*out_total));
        for (int i = 0; i < n_threads; ++i) {
            out_total->cells += jobs[i].total.cells;
            out_total->ext16 += jobs[i].total.ext16;
            out_total->columns += jobs[i].total.columns;
            if (jobs[i].total.max_width > out_total->max_width) out_total->max_width = jobs[i].total.max_width;
        }
    }
    free(jobs);
    free(th);
    return 0;
}

/* Per-locus reductions for a whole batch (trio or duo), threaded over loci: the CPU side of
 * tdb_trio_batch's second half, used as checker and cpu_baseline. */
typedef struct {
    const tdo_trio_locus *loci;
    size_t n_loci;
    const int32_t *scores;
    double q;
    int duo;
    tdo_allele_out *out;
    uint8_t *mask;
    size_t *next;
} rjob_t;

static void *rworker(void *arg) {
    rjob_t *j = (rjob_t *)arg;
    for (;;) {
        size_t b = __atomic_fetch_add(j->next, (size_t)1024, __ATOMIC_RELAXED);
        if (b >= j->n_loci) break;
        size_t e = b + 1024 < j->n_loci ? b + 1024 : j->n_loci;
        for (size_t i = b; i < e; ++i) {
            tdo_allele_out o[2];
            memset(o, 0, sizeof(o));
            if (j->duo) tdo_duo_assess(&j->loci[i], j->scores, j->q, o, j->mask);
            else tdo_trio_assess(&j->loci[i], j->scores, j->q, o, j->mask);
            j->out[2 * i] = o[0];
            j->out[2 * i + 1] = o[1];
        }
    }
    return NULL;
}

int tdo_assess_batch(const tdo_trio_locus *loci, size_t n_loci, const int32_t *scores, double q, int duo,
                     int n_threads, tdo_allele_out *out, uint8_t *mask) {
    if (n_threads < 1) n_threads = 1;
    size_t next = 0;
    rjob_t job = {loci, n_loci, scores, q, duo, out, mask, &next};
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int i = 1; i < n_threads; ++i) pthread_create(&th[i], NULL, rworker, &job);
    rworker(&job);
    for (int i = 1; i < n_threads; ++i) pthread_join(th[i], NULL);
    free(th);
    return 0;
}
