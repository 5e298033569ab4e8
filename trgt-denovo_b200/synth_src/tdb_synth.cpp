// tdb_synth.cpp -- deterministic synthetic trio/duo batches (include/tdb_synth.h).
// Host-only data generation for tests and bench.py; it performs no alignment arithmetic.
#include "../../include/tdb_synth.h"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    int poisson(double mean) {
        const double L = std::exp(-mean);
        int k = 0;
        double p = 1.0;
        do {
            ++k;
            p *= uni();
        } while (p > L);
        return k - 1;
    }
};

inline uint64_t mix(uint64_t a, uint64_t b) {
    Rng r(a ^ (b * 0xd6e8feb86659fd93ull + 0x2545f4914f6cdd1dull));
    return r.next();
}

const char BASES[4] = {'A', 'C', 'G', 'T'};

struct LocusPlan {
    char motif[8];
    int mlen;
    char lf[64], rf[64];
    int flank;
    int units[3][2];   // 0 mother, 1 father, 2 child
    int n_alleles[3];
    int n_reads[3][2];
    int denovo, denovo_allele;
    uint64_t seed;
};

void plan_locus(const tdb_synth_params &P, uint64_t li, LocusPlan &pl) {
    memset(&pl, 0, sizeof(pl));
    pl.seed = mix(P.seed, li);
    Rng r(pl.seed);
    // motif length ~ {1:.15, 2:.30, 3:.15, 4:.20, 5:.10, 6:.10}
    const double cum[6] = {0.15, 0.45, 0.60, 0.80, 0.90, 1.0};
    const double u = r.uni();
    int ml = 1;
    while (ml < 6 && u >= cum[ml - 1]) ++ml;
    pl.mlen = ml;
    for (;;) {
        bool homo = true;
        for (int i = 0; i < ml; ++i) {
            pl.motif[i] = BASES[r.below(4)];
            if (pl.motif[i] != pl.motif[0]) homo = false;
        }
        if (ml == 1 || !homo) break;
    }
    pl.flank = std::min(P.flank_len, 64);
    for (int i = 0; i < pl.flank; ++i) pl.lf[i] = BASES[r.below(4)];
    for (int i = 0; i < pl.flank; ++i) pl.rf[i] = BASES[r.below(4)];
    auto clampu = [&](int x) { return std::max(1, x); };
    pl.n_alleles[0] = pl.n_alleles[1] = pl.n_alleles[2] = 2;
    if (!P.long_expansion) {
        double len = std::exp(std::log(P.len_median) + P.len_sigma * r.normal());
        len = std::min((double)P.len_max, std::max((double)P.len_min, len));
        const int base = clampu((int)std::lround(len / ml));
        for (int s = 0; s < 2; ++s)
            for (int a = 0; a < 2; ++a) pl.units[s][a] = clampu(base + (int)r.below(5) - 2);
        pl.units[2][0] = pl.units[0][r.below(2)];
        pl.units[2][1] = pl.units[1][r.below(2)];
        if (r.uni() < P.denovo_rate) {
            pl.denovo = 1;
            pl.denovo_allele = (int)r.below(2);
            const int d = 1 + (int)r.below(10);
            const int sign = r.below(2) ? 1 : -1;
            pl.units[2][pl.denovo_allele] = clampu(pl.units[2][pl.denovo_allele] + sign * d);
        }
    } else {
        const int child_len = 1000 + (int)r.below(9001);
        const bool event = r.uni() < 0.5;
        const int pbase = event ? clampu((100 + (int)r.below(901)) / ml) : clampu(child_len / ml);
        for (int s = 0; s < 2; ++s)
            for (int a = 0; a < 2; ++a) pl.units[s][a] = clampu(pbase + (int)r.below(5) - 2);
        pl.units[2][0] = pl.units[0][r.below(2)];
        pl.units[2][1] = pl.units[1][r.below(2)];
        if (event) {
            pl.denovo = 1;
            pl.denovo_allele = (int)r.below(2);
            pl.units[2][pl.denovo_allele] = clampu(child_len / ml);
        }
    }
    if (P.duo) pl.n_alleles[1] = 0;
    for (int s = 0; s < 3; ++s)
        for (int a = 0; a < pl.n_alleles[s]; ++a) pl.n_reads[s][a] = r.poisson(P.reads_per_allele);
    // keep the reference's unwrap() in get_denovo_coverage reachable only through real dropout:
    // every parent gets at least one read so that synthetic benches do not model a panic
    for (int s = 0; s < 2; ++s)
        if (pl.n_alleles[s] && pl.n_reads[s][0] + pl.n_reads[s][1] == 0) pl.n_reads[s][0] = 1;
}

inline int allele_len(const LocusPlan &pl, int units) { return 2 * pl.flank + units * pl.mlen; }

void write_allele(const LocusPlan &pl, int units, uint8_t *dst) {
    memcpy(dst, pl.lf, (size_t)pl.flank);
    uint8_t *q = dst + pl.flank;
    for (int i = 0; i < units; ++i, q += pl.mlen) memcpy(q, pl.motif, (size_t)pl.mlen);
    memcpy(q, pl.rf, (size_t)pl.flank);
}

// One read of allele `units`: optional +-1 unit stutter, then substitutions / single-base indels placed
// by geometric skipping.  dst == nullptr: only the length is computed (same RNG stream).
int make_read(const tdb_synth_params &P, const LocusPlan &pl, int units, uint64_t seed, std::vector<uint8_t> &src,
              uint8_t *dst) {
    Rng r(seed);
    int u = units;
    if (r.uni() < P.stutter_prob) u = std::max(1, u + (r.below(2) ? 1 : -1));
    const int n = allele_len(pl, u);
    if (dst) {
        src.resize((size_t)n);
        write_allele(pl, u, src.data());
    }
    const double perr = P.sub_rate + P.indel_rate;
    const double lg = perr > 0 ? std::log1p(-perr) : 0.0;
    int pos = 0, out = 0;
    while (pos < n) {
        int gap = n;
        if (perr > 0) {
            double uu = r.uni();
            if (uu < 1e-300) uu = 1e-300;
            const double g = std::floor(std::log(uu) / lg);
            gap = g > (double)n ? n : (int)g;
        }
        const int e = std::min(n, pos + gap);
        if (dst) memcpy(dst + out, src.data() + pos, (size_t)(e - pos));
        out += e - pos;
        pos = e;
        if (pos >= n) break;
        const double t = r.uni() * perr;
        if (t < P.sub_rate) {
            if (dst) {
                uint8_t c;
                do c = (uint8_t)BASES[r.below(4)]; while (c == src[(size_t)pos]);
                dst[out] = c;
            } else {
                // keep the RNG stream identical to the writing pass: draw until different
                // (needs the source base: recompute it without materialising the read)
                const int F = pl.flank;
                const uint8_t sb = pos < F ? (uint8_t)pl.lf[pos]
                                 : (pos < F + u * pl.mlen ? (uint8_t)pl.motif[(pos - F) % pl.mlen]
                                                         : (uint8_t)pl.rf[pos - F - u * pl.mlen]);
                uint8_t c;
                do c = (uint8_t)BASES[r.below(4)]; while (c == sb);
            }
            ++out, ++pos;
        } else if (r.below(2)) { // deletion
            ++pos;
        } else { // insertion
            const uint8_t c = (uint8_t)BASES[r.below(4)];
            if (dst) dst[out] = c;
            ++out;
        }
    }
    return out;
}

struct LocusSize {
    uint32_t n_seqs, n_pairs;
    uint64_t bytes;
};

// sequence order inside a locus: child allele sequences, then reads of mother a0, a1, father a0, a1,
// child a0, a1.  pair order: for each child allele c: mother a0, a1, father a0, a1, own reads of c.
void size_locus(const tdb_synth_params &P, const LocusPlan &pl, LocusSize &sz) {
    std::vector<uint8_t> dummy;
    sz.n_seqs = 0, sz.n_pairs = 0, sz.bytes = 0;
    for (int c = 0; c < pl.n_alleles[2]; ++c) sz.bytes += (uint64_t)allele_len(pl, pl.units[2][c]), ++sz.n_seqs;
    for (int s = 0; s < 3; ++s)
        for (int a = 0; a < pl.n_alleles[s]; ++a)
            for (int i = 0; i < pl.n_reads[s][a]; ++i) {
                sz.bytes += (uint64_t)make_read(P, pl, pl.units[s][a], mix(pl.seed, (uint64_t)(s * 2 + a) << 32 | (uint32_t)i),
                                                dummy, nullptr);
                ++sz.n_seqs;
            }
    const int parents = pl.n_reads[0][0] + pl.n_reads[0][1] + pl.n_reads[1][0] + pl.n_reads[1][1];
    for (int c = 0; c < pl.n_alleles[2]; ++c) sz.n_pairs += (uint32_t)(parents + pl.n_reads[2][c]);
}

template <typename T>
T *alloc_arr(size_t n, bool pinned) {
    void *p = nullptr;
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T) + 64;
    if (pinned) {
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    } else {
        p = malloc(bytes);
    }
    return (T *)p;
}
void free_arr(void *p, bool pinned) {
    if (!p) return;
    if (pinned) cudaFreeHost(p);
    else free(p);
}

} // namespace

extern "C" {

void tdb_synth_default_params(tdb_synth_params *p) {
    memset(p, 0, sizeof(*p));
    p->seed = 20250002ull;
    p->flank_len = 50;
    p->reads_per_allele = 15.0;
    p->denovo_rate = 1e-4;
    p->sub_rate = 2e-4;
    p->indel_rate = 1e-3;
    p->stutter_prob = 0.02;
    p->len_median = 40.0;
    p->len_sigma = 0.8;
    p->len_min = 6;
    p->len_max = 1000;
}

void tdb_synth_free(tdb_synth_batch *b) {
    if (!b) return;
    const bool pin = b->pinned != 0;
    free_arr(b->blob, pin), free_arr(b->seq_off, pin), free_arr(b->seq_len, pin);
    free_arr(b->pair_pattern, pin), free_arr(b->pair_text, pin), free_arr(b->loci, pin);
    free(b->meta);
    memset(b, 0, sizeof(*b));
}

int tdb_synth_generate(const tdb_synth_params *pp, uint64_t l0, uint64_t l1, int n_threads, tdb_synth_batch *out) {
    if (!pp || !out || l1 < l0) return TDB_E_INVALID;
    const tdb_synth_params P = *pp;
    memset(out, 0, sizeof(*out));
    const size_t n = (size_t)(l1 - l0);
    n_threads = std::max(1, n_threads);
    std::vector<LocusSize> sizes(n);
    auto par = [&](auto fn) {
        std::vector<std::thread> th;
        const size_t per = (n + (size_t)n_threads - 1) / (size_t)n_threads;
        for (int t = 0; t < n_threads; ++t) {
            const size_t b = std::min(n, per * (size_t)t), e = std::min(n, b + per);
            if (b < e) th.emplace_back(fn, b, e);
        }
        for (auto &x : th) x.join();
    };
    par([&](size_t b, size_t e) {
        LocusPlan pl;
        for (size_t i = b; i < e; ++i) {
            plan_locus(P, l0 + i, pl);
            size_locus(P, pl, sizes[i]);
        }
    });
    std::vector<uint64_t> seq0(n + 1), pair0(n + 1), byte0(n + 1);
    seq0[0] = pair0[0] = byte0[0] = 0;
    for (size_t i = 0; i < n; ++i) {
        seq0[i + 1] = seq0[i] + sizes[i].n_seqs;
        pair0[i + 1] = pair0[i] + sizes[i].n_pairs;
        byte0[i + 1] = byte0[i] + sizes[i].bytes;
    }
    if (seq0[n] >= 0xffffffffull || pair0[n] >= 0xffffffffull) return TDB_E_UNSUPPORTED;
    const bool pin = P.pinned != 0;
    out->pinned = P.pinned;
    out->n_loci = n, out->n_seqs = (size_t)seq0[n], out->n_pairs = (size_t)pair0[n], out->blob_bytes = (size_t)byte0[n];
    out->blob = alloc_arr<uint8_t>(out->blob_bytes + 64, pin);
    out->seq_off = alloc_arr<uint64_t>(out->n_seqs, pin);
    out->seq_len = alloc_arr<uint32_t>(out->n_seqs, pin);
    out->pair_pattern = alloc_arr<uint32_t>(out->n_pairs, pin);
    out->pair_text = alloc_arr<uint32_t>(out->n_pairs, pin);
    out->loci = alloc_arr<tdb_trio_locus>(n, pin);
    out->meta = (tdb_synth_meta *)calloc(std::max<size_t>(n, 1), sizeof(tdb_synth_meta));
    if (!out->blob || !out->seq_off || !out->seq_len || !out->pair_pattern || !out->pair_text || !out->loci || !out->meta) {
        tdb_synth_free(out);
        return TDB_E_OOM;
    }
    std::vector<uint64_t> nm_part((size_t)n_threads + 1, 0), bases_part((size_t)n_threads + 1, 0);
    std::vector<std::thread> th;
    const size_t per = (n + (size_t)n_threads - 1) / (size_t)n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const size_t b = std::min(n, per * (size_t)t), e = std::min(n, b + per);
        if (b >= e) continue;
        th.emplace_back([&, b, e, t]() {
            LocusPlan pl;
            std::vector<uint8_t> src;
            uint64_t nm = 0, bases = 0;
            for (size_t i = b; i < e; ++i) {
                plan_locus(P, l0 + i, pl);
                uint32_t si = (uint32_t)seq0[i], pi = (uint32_t)pair0[i];
                uint64_t off = byte0[i];
                tdb_synth_meta &m = out->meta[i];
                memcpy(m.motif, pl.motif, 8);
                m.denovo = pl.denovo, m.denovo_allele = pl.denovo_allele;
                m.seq_begin = si, m.pair_begin = pi;
                uint32_t child_seq[2] = {0, 0};
                for (int c = 0; c < pl.n_alleles[2]; ++c) {
                    const int len = allele_len(pl, pl.units[2][c]);
                    write_allele(pl, pl.units[2][c], out->blob + off);
                    out->seq_off[si] = off, out->seq_len[si] = (uint32_t)len;
                    child_seq[c] = si++;
                    off += (uint64_t)len;
                }
                uint32_t first_read[3][2] = {{0, 0}, {0, 0}, {0, 0}};
                for (int s = 0; s < 3; ++s)
                    for (int a = 0; a < 2; ++a) {
                        m.units[s][a] = a < pl.n_alleles[s] ? pl.units[s][a] : 0;
                        m.n_reads[s][a] = a < pl.n_alleles[s] ? pl.n_reads[s][a] : 0;
                        first_read[s][a] = si;
                        if (a >= pl.n_alleles[s]) continue;
                        for (int k = 0; k < pl.n_reads[s][a]; ++k) {
                            const int len = make_read(P, pl, pl.units[s][a],
                                                      mix(pl.seed, (uint64_t)(s * 2 + a) << 32 | (uint32_t)k), src,
                                                      out->blob + off);
                            out->seq_off[si] = off, out->seq_len[si] = (uint32_t)len;
                            ++si;
                            off += (uint64_t)len;
                        }
                    }
                tdb_trio_locus &L = out->loci[i];
                memset(&L, 0, sizeof(L));
                L.n_child = pl.n_alleles[2], L.n_mother = pl.n_alleles[0], L.n_father = pl.n_alleles[1];
                for (int a = 0; a < 2; ++a) {
                    L.child_len[a] = a < pl.n_alleles[2] ? allele_len(pl, pl.units[2][a]) : 0;
                    L.mother_len[a] = a < pl.n_alleles[0] ? allele_len(pl, pl.units[0][a]) : 0;
                    L.father_len[a] = a < pl.n_alleles[1] ? allele_len(pl, pl.units[1][a]) : 0;
                }
                for (int c = 0; c < pl.n_alleles[2]; ++c) {
                    const uint32_t tgt = child_seq[c];
                    const uint64_t tl = out->seq_len[tgt];
                    int j = 0;
                    for (int s = 0; s < 2; ++s)
                        for (int a = 0; a < 2; ++a, ++j) {
                            L.list_off[c][j] = pi;
                            const int cnt = a < pl.n_alleles[s] ? pl.n_reads[s][a] : 0;
                            for (int k = 0; k < cnt; ++k) {
                                const uint32_t rd = first_read[s][a] + (uint32_t)k;
                                out->pair_pattern[pi] = rd, out->pair_text[pi] = tgt;
                                nm += (uint64_t)out->seq_len[rd] * tl, bases += out->seq_len[rd] + tl;
                                ++pi;
                            }
                        }
                    L.list_off[c][4] = pi;
                    for (int k = 0; k < pl.n_reads[2][c]; ++k) {
                        const uint32_t rd = first_read[2][c] + (uint32_t)k;
                        out->pair_pattern[pi] = rd, out->pair_text[pi] = tgt;
                        nm += (uint64_t)out->seq_len[rd] * tl, bases += out->seq_len[rd] + tl;
                        ++pi;
                    }
                    L.list_off[c][5] = pi;
                }
                m.seq_end = si, m.pair_end = pi;
            }
            nm_part[(size_t)t] = nm, bases_part[(size_t)t] = bases;
        });
    }
    for (auto &x : th) x.join();
    for (int t = 0; t < n_threads; ++t) out->sum_nm += nm_part[(size_t)t], out->sum_bases += bases_part[(size_t)t];
    return TDB_OK;
}

int tdb_synth_micro(uint64_t seed, uint32_t len, double divergence, size_t n_pairs, int n_threads, int pinned,
                    tdb_synth_batch *out) {
    if (!out) return TDB_E_INVALID;
    memset(out, 0, sizeof(*out));
    const bool pin = pinned != 0;
    out->pinned = pinned;
    n_threads = std::max(1, n_threads);
    // pass 1: pattern lengths
    std::vector<uint32_t> plen(n_pairs);
    auto gen = [&](size_t i, uint8_t *text, uint8_t *pat) -> uint32_t {
        Rng r(mix(seed, i));
        uint32_t o = 0;
        for (uint32_t k = 0; k < len; ++k) {
            const uint8_t c = (uint8_t)BASES[r.below(4)];
            if (text) text[k] = c;
            const double u = r.uni();
            if (u < divergence / 3) { // substitution
                uint8_t d;
                do d = (uint8_t)BASES[r.below(4)]; while (d == c);
                if (pat) pat[o] = d;
                ++o;
            } else if (u < 2 * divergence / 3) { // deletion from the pattern
            } else if (u < divergence) { // insertion into the pattern
                const uint8_t d = (uint8_t)BASES[r.below(4)];
                if (pat) pat[o] = c, pat[o + 1] = d;
                o += 2;
            } else {
                if (pat) pat[o] = c;
                ++o;
            }
        }
        return o;
    };
    auto par = [&](auto fn) {
        std::vector<std::thread> th;
        const size_t per = (n_pairs + (size_t)n_threads - 1) / (size_t)n_threads;
        for (int t = 0; t < n_threads; ++t) {
            const size_t b = std::min(n_pairs, per * (size_t)t), e = std::min(n_pairs, b + per);
            if (b < e) th.emplace_back(fn, b, e);
        }
        for (auto &x : th) x.join();
    };
    par([&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) plen[i] = gen(i, nullptr, nullptr);
    });
    std::vector<uint64_t> off(n_pairs + 1);
    off[0] = 0;
    for (size_t i = 0; i < n_pairs; ++i) off[i + 1] = off[i] + len + plen[i];
    out->n_seqs = 2 * n_pairs, out->n_pairs = n_pairs, out->blob_bytes = (size_t)off[n_pairs];
    out->blob = alloc_arr<uint8_t>(out->blob_bytes + 64, pin);
    out->seq_off = alloc_arr<uint64_t>(out->n_seqs, pin);
    out->seq_len = alloc_arr<uint32_t>(out->n_seqs, pin);
    out->pair_pattern = alloc_arr<uint32_t>(out->n_pairs, pin);
    out->pair_text = alloc_arr<uint32_t>(out->n_pairs, pin);
    if (!out->blob || !out->seq_off || !out->seq_len || !out->pair_pattern || !out->pair_text) {
        tdb_synth_free(out);
        return TDB_E_OOM;
    }
    par([&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            uint8_t *t = out->blob + off[i], *p = t + len;
            gen(i, t, p);
            out->seq_off[2 * i] = off[i], out->seq_len[2 * i] = len;
            out->seq_off[2 * i + 1] = off[i] + len, out->seq_len[2 * i + 1] = plen[i];
            out->pair_pattern[i] = (uint32_t)(2 * i + 1), out->pair_text[i] = (uint32_t)(2 * i);
        }
    });
    for (size_t i = 0; i < n_pairs; ++i) out->sum_nm += (uint64_t)len * plen[i], out->sum_bases += len + plen[i];
    return TDB_OK;
}

} // extern "C"
