// tdb_hostpack.cpp -- ASCII -> 2-bit packing on the caller's CPU (see tdb_hostpack.h).
#include "tdb_hostpack.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

namespace tdb {
namespace {

inline uint32_t nwords(uint32_t len) { return (len + 15u) / 16u + 1u; }
inline size_t woff(uint64_t off, uint32_t i) { return (size_t)(uint32_t)(off >> 4) + 2u * (size_t)i; }

inline bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// returns true when the sequence holds a byte outside ACGT
bool pack_seq_scalar(const uint8_t *src, uint32_t len, uint32_t *dst) {
    const uint32_t nw = nwords(len);
    bool bad = false;
    for (uint32_t j = 0; j < nw; ++j) {
        uint32_t w = 0;
        const uint32_t base = j * 16u;
        for (uint32_t k = 0; k < 16u && base + k < len; ++k) {
            const uint8_t c = src[base + k];
            bad |= !is_acgt(c);
            w |= (uint32_t)((c >> 1) & 3u) << (2 * k);
        }
        dst[j] = w;
    }
    return bad;
}

__attribute__((target("avx2"))) bool pack_seq_avx2(const uint8_t *src, uint32_t len, uint32_t *dst) {
    const uint32_t nw = nwords(len);
    const __m256i three = _mm256_set1_epi8(3), w1 = _mm256_set1_epi16(0x0401), w2 = _mm256_set1_epi32(0x00100001);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'),
                  cT = _mm256_set1_epi8('T');
    uint32_t full = len / 32u; // whole 32-byte blocks -> 2 words each
    uint32_t bad = 0;
    uint8_t *out = reinterpret_cast<uint8_t *>(dst);
    for (uint32_t b = 0; b < full; ++b) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 32u * b));
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, cA), _mm256_cmpeq_epi8(v, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(v, cG), _mm256_cmpeq_epi8(v, cT)));
        bad |= ~(uint32_t)_mm256_movemask_epi8(ok);
        const __m256i c = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
        const __m256i u = _mm256_madd_epi16(_mm256_maddubs_epi16(c, w1), w2); // one packed byte per dword
        // gather the low byte of the 8 dwords
        const __m256i sh = _mm256_shuffle_epi8(
            u, _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1,
                                -1, -1, -1, -1, -1, -1, -1, -1));
        const uint32_t lo = (uint32_t)_mm256_extract_epi32(sh, 0), hi = (uint32_t)_mm256_extract_epi32(sh, 4);
        memcpy(out + 8u * b, &lo, 4);
        memcpy(out + 8u * b + 4, &hi, 4);
    }
    // tail (< 32 bases) and the trailing zero word
    bool tbad = false;
    for (uint32_t j = full * 2u; j < nw; ++j) {
        uint32_t w = 0;
        const uint32_t base = j * 16u;
        for (uint32_t k = 0; k < 16u && base + k < len; ++k) {
            const uint8_t c = src[base + k];
            tbad |= !is_acgt(c);
            w |= (uint32_t)((c >> 1) & 3u) << (2 * k);
        }
        dst[j] = w;
    }
    return bad != 0 || tbad;
}

__attribute__((target("avx512f,avx512bw,avx512vl"))) bool pack_seq_avx512(const uint8_t *src, uint32_t len,
                                                                         uint32_t *dst) {
    const uint32_t out_bytes = nwords(len) * 4u;
    const __m512i three = _mm512_set1_epi8(3), w1 = _mm512_set1_epi16(0x0401), w2 = _mm512_set1_epi32(0x00100001);
    const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'),
                  cT = _mm512_set1_epi8('T');
    uint8_t *out = reinterpret_cast<uint8_t *>(dst);
    __mmask64 bad = 0;
    for (uint32_t ob = 0, ib = 0; ob < out_bytes; ob += 16u, ib += 64u) {
        const uint32_t in_left = ib < len ? len - ib : 0u;
        const __mmask64 im = in_left >= 64u ? ~0ull : ((1ull << in_left) - 1ull);
        const __m512i v = _mm512_maskz_loadu_epi8(im, src + ib);
        const __mmask64 ok = _mm512_cmpeq_epi8_mask(v, cA) | _mm512_cmpeq_epi8_mask(v, cC) |
                             _mm512_cmpeq_epi8_mask(v, cG) | _mm512_cmpeq_epi8_mask(v, cT);
        bad |= ~ok & im;
        const __m512i c = _mm512_and_si512(_mm512_srli_epi16(v, 1), three);
        const __m512i u = _mm512_madd_epi16(_mm512_maddubs_epi16(c, w1), w2);
        const __m128i o = _mm512_cvtepi32_epi8(u);
        const uint32_t out_left = out_bytes - ob;
        if (out_left >= 16u) _mm_storeu_si128(reinterpret_cast<__m128i *>(out + ob), o);
        else _mm_mask_storeu_epi8(out + ob, (__mmask16)((1u << out_left) - 1u), o);
    }
    return bad != 0;
}

typedef bool (*pack_fn)(const uint8_t *, uint32_t, uint32_t *);

pack_fn pick(const char **name) {
    __builtin_cpu_init();
    const char *force = getenv("TDB_HOSTPACK_ISA"); // testing: "avx2" or "scalar" caps the dispatch
    const bool no512 = force && (!strcmp(force, "avx2") || !strcmp(force, "scalar"));
    const bool no256 = force && !strcmp(force, "scalar");
    if (!no512 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) {
        *name = "avx512bw";
        return pack_seq_avx512;
    }
    if (!no256 && __builtin_cpu_supports("avx2")) {
        *name = "avx2";
        return pack_seq_avx2;
    }
    *name = "scalar";
    return pack_seq_scalar;
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

size_t hostpack_range(const uint8_t *blob, size_t blob_bytes, const uint64_t *seq_off, const uint32_t *seq_len,
                      uint32_t s0, uint32_t s1, uint32_t n_seqs, uint32_t *packed, uint8_t *flag) {
    const pack_fn fn = dispatch().fn;
    size_t violations = 0;
    for (uint32_t i = s0; i < s1; ++i) {
        const uint64_t off = seq_off[i];
        const uint32_t len = seq_len[i];
        bool bad_layout = off > blob_bytes || (uint64_t)len > blob_bytes - off;
        if (!bad_layout && i + 1 < n_seqs) bad_layout = off + len > seq_off[i + 1];
        if (bad_layout) {
            ++violations;
            flag[i] = 2;
            continue;
        }
        flag[i] = fn(blob + off, len, packed + woff(off, i)) ? 1 : 0;
    }
    return violations;
}

} // namespace tdb
