/*
 * tdb_synth.h -- deterministic synthetic trio / duo workload generator.
 *
 * Not part of the drop-in boundary: this is the data source for tests and bench.py.  The reference
 * has no fixtures and its real inputs need htslib (BAM/VCF), which is out of scope for this path;
 * the generator emits exactly what `load_alleles` (src/allele.rs:146-191) would hand to
 * `assess_denovo`: per sample the allele sequences `left_flank + allele + right_flank`
 * (src/allele.rs:406-426) and the spanning reads partitioned by their AL tag (src/allele.rs:249-259),
 * already laid out as a batch for include/trgt_denovo_b200.h.
 *
 * Every locus is a pure function of (seed, locus index): any range can be generated on any rank.
 */
#ifndef TDB_SYNTH_H
#define TDB_SYNTH_H

#include "trgt_denovo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t seed;
    int32_t duo;            /* 0 trio; 1 duo (A = child, B = mother), SURVEY.md 8(d) config 3 */
    int32_t flank_len;      /* 50 */
    double reads_per_allele; /* Poisson mean: 15 (30x), 30 (60x) */
    double denovo_rate;     /* fraction of loci whose child allele gains/loses U{1..10} motif units */
    double sub_rate;        /* 2e-4 per base */
    double indel_rate;      /* 1e-3 per base, single-base */
    double stutter_prob;    /* 0.02 per read: +-1 motif unit */
    double len_median;      /* allele length ~ LogNormal(median 40 bp, sigma 0.8) clipped [6,1000] */
    double len_sigma;
    int32_t len_min, len_max;
    int32_t long_expansion; /* config 4: child allele U[1000,10000] bp, parents 100..1000 bp or equal */
    int32_t pinned;         /* 1: allocate the batch arrays with cudaHostAlloc (faster H2D) */
} tdb_synth_params;

typedef struct {
    char motif[8];
    int32_t units[3][2];   /* motif units of allele a of sample s (0 mother, 1 father, 2 child) */
    int32_t n_reads[3][2]; /* reads partitioned to that allele */
    int32_t denovo;        /* 1 when a de novo change was injected; child allele index in denovo_allele */
    int32_t denovo_allele;
    uint32_t seq_begin, seq_end;   /* this locus' sequences in the batch tables */
    uint32_t pair_begin, pair_end; /* this locus' pairs */
} tdb_synth_meta;

typedef struct {
    uint8_t *blob;
    size_t blob_bytes;
    uint64_t *seq_off;
    uint32_t *seq_len;
    size_t n_seqs;
    uint32_t *pair_pattern, *pair_text;
    size_t n_pairs;
    tdb_trio_locus *loci;
    tdb_synth_meta *meta;
    size_t n_loci;
    uint64_t sum_nm;   /* sum over pairs of n*m (GCUPS numerator) */
    uint64_t sum_bases; /* sum over pairs of n+m */
    int32_t pinned;
} tdb_synth_batch;

void tdb_synth_default_params(tdb_synth_params *p);
/* Generates loci [locus_begin, locus_end).  Returns 0 or a TDB_E_* code. */
int tdb_synth_generate(const tdb_synth_params *p, uint64_t locus_begin, uint64_t locus_end, int n_threads,
                       tdb_synth_batch *out);
void tdb_synth_free(tdb_synth_batch *b);

/* Microbenchmark pairs: n_pairs pairs of a uniform-random ACGT text of
 * length `len` and a pattern derived from it with `divergence` (sub:ins:del = 1:1:1).  Only
 * blob/seq_off/seq_len/pair_* of `out` are filled (n_loci = 0). */
int tdb_synth_micro(uint64_t seed, uint32_t len, double divergence, size_t n_pairs, int n_threads, int pinned,
                    tdb_synth_batch *out);

#ifdef __cplusplus
}
#endif
#endif
