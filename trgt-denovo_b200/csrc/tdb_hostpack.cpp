// tdb_hostpack.cpp -- ASCII -> 2-bit packing on the caller's CPU (see tdb_hostpack.h).
#include "tdb_hostpack.h"

#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace tdb {
namespace {

inline bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// Flags the sequence that owns blob position x (seq_off ascending; zero-length sequences may share offsets).
void flag_owner(uint64_t x, const uint64_t *seq_off, const uint32_t *seq_len, uint32_t n_seqs, uint8_t *flag) {
    uint32_t lo = 0, hi = n_seqs;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (seq_off[mid] <= x) lo = mid + 1;
        else hi = mid;
    }
    for (uint32_t i = lo; i > 0;) {
        --i;
        const uint64_t o = seq_off[i];
        if (x < o + seq_len[i]) {
            flag[i] = 1;
            return;
        }
        if (seq_len[i] != 0) return;
    }
}

struct Owner {
    mutable const uint64_t *seq_off; // null for a dense layout until a byte outside ACGT asks for it
    const uint32_t *seq_len;
    uint32_t n_seqs;
    uint8_t *flag;
    const uint64_t *(*make_off)(void *);
    void *arg;
    const uint64_t *off() const {
        if (!seq_off) seq_off = make_off(arg);
        return seq_off;
    }
};

// All three variants pack bytes [x0, x1) of the blob into words [x0/16, ceil(x1/16)); x0 % 64 == 0.
void pack_scalar(const uint8_t *blob, uint64_t x0, uint64_t x1, uint32_t *packed, const Owner &ow) {
    for (uint64_t x = x0; x < x1; x += 16) {
        uint32_t w = 0;
        for (uint32_t k = 0; k < 16u && x + k < x1; ++k) {
            const uint8_t c = blob[x + k];
            if (!is_acgt(c)) flag_owner(x + k, ow.off(), ow.seq_len, ow.n_seqs, ow.flag);
            w |= (uint32_t)((c >> 1) & 3u) << (2 * k);
        }
        packed[x >> 4] = w;
    }
}

__attribute__((target("avx2"))) void pack_avx2(const uint8_t *blob, uint64_t x0, uint64_t x1, uint32_t *packed,
                                               const Owner &ow) {
    const __m256i three = _mm256_set1_epi8(3), w1 = _mm256_set1_epi16(0x0401), w2 = _mm256_set1_epi32(0x00100001);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'),
                  cT = _mm256_set1_epi8('T');
    const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1,
                                          -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    uint64_t x = x0;
    for (; x + 32 <= x1; x += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(blob + x));
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, cA), _mm256_cmpeq_epi8(v, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(v, cG), _mm256_cmpeq_epi8(v, cT)));
        uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(ok);
        const __m256i c = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
        const __m256i u = _mm256_madd_epi16(_mm256_maddubs_epi16(c, w1), w2); // one packed byte per dword
        const __m256i sh = _mm256_shuffle_epi8(u, pick);
        packed[x >> 4] = (uint32_t)_mm256_extract_epi32(sh, 0);
        packed[(x >> 4) + 1] = (uint32_t)_mm256_extract_epi32(sh, 4);
        while (bad) {
            const int k = __builtin_ctz(bad);
            bad &= bad - 1;
            flag_owner(x + (uint64_t)k, ow.off(), ow.seq_len, ow.n_seqs, ow.flag);
        }
    }
    if (x < x1) pack_scalar(blob, x, x1, packed, ow);
}

__attribute__((target("avx512f,avx512bw,avx512vl"))) void pack_avx512(const uint8_t *blob, uint64_t x0, uint64_t x1,
                                                                     uint32_t *packed, const Owner &ow) {
    const __m512i three = _mm512_set1_epi8(3), w1 = _mm512_set1_epi16(0x0401), w2 = _mm512_set1_epi32(0x00100001);
    const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'),
                  cT = _mm512_set1_epi8('T');
    uint8_t *out = reinterpret_cast<uint8_t *>(packed);
    // software prefetch: one thread of the B200 host sustains 8 GB/s without it and 13 GB/s with it (the
    // per-core hardware prefetcher does not keep enough lines in flight); non-temporal stores skip the
    // read-for-ownership of the packed stream (measured by profiles/hostpack_bw.py)
    static const int pf = getenv("TDB_HOSTPACK_PREFETCH") ? atoi(getenv("TDB_HOSTPACK_PREFETCH")) : 4096;
    static const bool nt = getenv("TDB_HOSTPACK_NT") ? atoi(getenv("TDB_HOSTPACK_NT")) != 0 : true;
    const bool aligned_out = ((reinterpret_cast<uintptr_t>(packed)) & 15) == 0;
    for (uint64_t x = x0; x < x1; x += 64) {
        const uint64_t left = x1 - x;
        if (pf) _mm_prefetch(reinterpret_cast<const char *>(blob + x + pf), _MM_HINT_T0);
        const __mmask64 im = left >= 64 ? ~0ull : ((1ull << left) - 1ull);
        const __m512i v = _mm512_maskz_loadu_epi8(im, blob + x);
        const __mmask64 ok = _mm512_cmpeq_epi8_mask(v, cA) | _mm512_cmpeq_epi8_mask(v, cC) |
                             _mm512_cmpeq_epi8_mask(v, cG) | _mm512_cmpeq_epi8_mask(v, cT);
        uint64_t bad = ~ok & im;
        const __m512i c = _mm512_and_si512(_mm512_srli_epi16(v, 1), three);
        const __m512i u = _mm512_madd_epi16(_mm512_maddubs_epi16(c, w1), w2);
        const __m128i o = _mm512_cvtepi32_epi8(u);
        if (left >= 64) {
            if (nt && aligned_out) _mm_stream_si128(reinterpret_cast<__m128i *>(out + (x >> 2)), o);
            else _mm_storeu_si128(reinterpret_cast<__m128i *>(out + (x >> 2)), o);
        } else { // whole words only: ceil(left / 16) of them
            const uint32_t nbytes = (uint32_t)((left + 15) / 16) * 4u;
            _mm_mask_storeu_epi8(out + (x >> 2), (__mmask16)((1u << nbytes) - 1u), o);
        }
        while (bad) {
            const int k = __builtin_ctzll(bad);
            bad &= bad - 1;
            flag_owner(x + (uint64_t)k, ow.off(), ow.seq_len, ow.n_seqs, ow.flag);
        }
    }
    _mm_sfence(); // order the streaming stores before the block is published as done
}

typedef void (*pack_fn)(const uint8_t *, uint64_t, uint64_t, uint32_t *, const Owner &);

pack_fn pick(const char **name) {
    __builtin_cpu_init();
    const char *force = getenv("TDB_HOSTPACK_ISA"); // testing: "avx2" or "scalar" caps the dispatch
    const bool no512 = force && (!strcmp(force, "avx2") || !strcmp(force, "scalar"));
    const bool no256 = force && !strcmp(force, "scalar");
    if (!no512 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) {
        *name = "avx512bw";
        return pack_avx512;
    }
    if (!no256 && __builtin_cpu_supports("avx2")) {
        *name = "avx2";
        return pack_avx2;
    }
    *name = "scalar";
    return pack_scalar;
}

struct Dispatch {
    const char *name;
    pack_fn fn;
    Dispatch() { fn = pick(&name); }
};
const Dispatch &dispatch() {
    static const Dispatch d;
    return d;
}

} // namespace

const char *hostpack_isa() { return dispatch().name; }

void hostpack_stream(const uint8_t *blob, uint64_t x0, uint64_t x1, uint32_t *packed, const uint64_t *seq_off,
                     const uint32_t *seq_len, uint32_t n_seqs, uint8_t *flag, const uint64_t *(*make_off)(void *),
                     void *make_arg) {
    const Owner ow = {seq_off, seq_len, n_seqs, flag, make_off, make_arg};
    if (x1 > x0) dispatch().fn(blob, x0, x1, packed, ow);
}

size_t hostpack_check_layout(const uint64_t *seq_off, const uint32_t *seq_len, uint32_t s0, uint32_t s1, uint32_t n_seqs,
                             size_t blob_bytes) {
    size_t violations = 0;
    for (uint32_t i = s0; i < s1; ++i) {
        const uint64_t off = seq_off[i];
        const uint32_t len = seq_len[i];
        bool bad = off > blob_bytes || (uint64_t)len > blob_bytes - off;
        if (!bad && i + 1 < n_seqs) bad = off + len > seq_off[i + 1];
        violations += bad;
    }
    return violations;
}

} // namespace tdb
