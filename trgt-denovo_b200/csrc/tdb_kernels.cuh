// tdb_kernels.cuh -- kernels of the hot path: sequence packing, tiered WFA alignment, fused reduction.
#pragma once
#include "tdb_wfa.cuh"

#include "../../include/trgt_denovo_b200.h"

namespace tdb {

// ------------------------------------------------------------------------------------------------
// Packing: raw ASCII -> 2-bit words (the role north_star gives locus.rs/read.rs, done at HBM speed).
// The blob is packed as ONE stream, independent of sequence boundaries: the byte at blob position x
// goes to word x/16, bits 2*(x%16), code (byte >> 1) & 3.  A sequence is a window of that stream
// (word seq_off/16, shift seq_off%16); the aligners read it with a funnel shift, so no per-sequence
// padding or offset table exists.  One thread packs 16 bytes into one word: fully coalesced 16-byte
// loads and 4-byte stores.  A byte outside {A,C,G,T} flags the sequence that owns it (binary search
// in seq_off, rare); pairs touching a flagged sequence take the byte-compare path so that raw byte
// equality (N==N, case-sensitive) is preserved exactly as WFA2 sees it.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack4(uint32_t w) { // 4 ASCII bytes -> 8 bits
    const uint32_t c = (w >> 1) & 0x03030303u;
    return (c | (c >> 6) | (c >> 12) | (c >> 18)) & 0xffu;
}
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 33)) * 0xff51afd7ed558ccdull;
    z = (z ^ (z >> 33)) * 0xc4ceb9fe1a85ec53ull;
    return z ^ (z >> 33);
}
constexpr uint64_t HT_EMPTY = ~0ull;

// error bits raised by the kernels into the ctrl block
enum { ERR_LAYOUT = 1, ERR_PAIR_INDEX = 2 };

// Sets flag = 1 on the sequence of [s_lo, s_hi) that owns blob position x (seq_off sorted ascending).
__device__ inline void flag_owner(uint64_t x, const uint64_t *__restrict__ seq_off,
                                  const uint32_t *__restrict__ seq_len, uint32_t s_lo, uint32_t s_hi,
                                  uint8_t *__restrict__ seq_flag) {
    uint32_t lo = s_lo, hi = s_hi; // last i in [s_lo, s_hi) with seq_off[i] <= x
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (seq_off[mid] <= x) lo = mid + 1;
        else hi = mid;
    }
    // zero-length sequences may share an offset with their successor: walk back over them
    for (uint32_t i = lo; i > s_lo;) {
        --i;
        const uint64_t o = seq_off[i];
        if (x < o + seq_len[i]) {
            seq_flag[i] |= 1;
            return;
        }
        if (o + seq_len[i] <= x && seq_len[i] != 0) return;
    }
}

__device__ __forceinline__ void load16(const uint8_t *__restrict__ blob, uint64_t x0, uint32_t cnt, bool aligned,
                                       uint32_t q[4]) {
    if (aligned && cnt == 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(blob + x0));
        q[0] = v.x, q[1] = v.y, q[2] = v.z, q[3] = v.w;
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint32_t w = 0;
            for (int u = 0; u < 4; ++u) {
                const uint32_t bi = (uint32_t)(t * 4 + u);
                if (bi < cnt) w |= (uint32_t)blob[x0 + bi] << (8 * u);
            }
            q[t] = w;
        }
    }
}
// 16 bytes -> one packed word; badm gets one bit per byte outside {A,C,G,T}.  Validity is a round trip: the
// byte the 2-bit code stands for (one PRMT from the table 'A','C','T','G') must equal the byte itself.
template <bool FULL>
__device__ __forceinline__ uint32_t pack16_t(const uint32_t q[4], uint32_t cnt, uint32_t &badm) {
    uint32_t word = 0, anydiff = 0, diff[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int rem = (int)cnt - 4 * t;
        const uint32_t need = FULL || rem >= 4 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (8 * rem)) - 1u));
        const uint32_t p8 = pack4(q[t] & need);
        const uint32_t y = (p8 | (p8 << 4)) & 0x0f0fu;
        const uint32_t sel = (y | (y << 2)) & 0x3333u; // the four codes as PRMT selector nibbles
        diff[t] = (q[t] ^ __byte_perm(0x47544341u, 0u, sel)) & need;
        anydiff |= diff[t];
        word |= p8 << (8 * t);
    }
    badm = 0;
    if (anydiff) { // rare
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t nz = (diff[t] | ((diff[t] & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u; // 0x80 per bad byte
            badm |= (((nz >> 7) & 1u) | ((nz >> 14) & 2u) | ((nz >> 21) & 4u) | ((nz >> 28) & 8u)) << (4 * t);
        }
    }
    return word;
}
// whole groups (all but the last of a range) skip the per-byte masks
__device__ __forceinline__ uint32_t pack16(const uint32_t q[4], uint32_t cnt, uint32_t &badm) {
    return cnt == 16 ? pack16_t<true>(q, 16, badm) : pack16_t<false>(q, cnt, badm);
}

// Packs blob bytes [b0, b1) (b0 a multiple of 16).  seq_flag of [s_lo, s_hi) must be zeroed before.
// Two 16-byte groups per thread (a whole block covers a contiguous 8 KB), both loads issued before use.
__global__ void __launch_bounds__(256) k_pack(const uint8_t *__restrict__ blob, uint64_t b0, uint64_t b1,
                                              const uint64_t *__restrict__ seq_off,
                                              const uint32_t *__restrict__ seq_len, uint32_t s_lo, uint32_t s_hi,
                                              uint32_t *__restrict__ packed, uint8_t *__restrict__ seq_flag) {
    const uint64_t ngroups = (b1 - b0 + 15) / 16;
    const bool aligned = (((uintptr_t)blob) & 15) == 0;
    const uint64_t stride = (uint64_t)gridDim.x * 512;
    for (uint64_t g0 = (uint64_t)blockIdx.x * 512 + threadIdx.x; g0 < ngroups; g0 += stride) {
        const uint64_t g1 = g0 + 256;
        const uint64_t xa = b0 + g0 * 16, xb = b0 + g1 * 16;
        const uint32_t ca = (uint32_t)min((uint64_t)16, b1 - xa);
        const uint32_t cb = g1 < ngroups ? (uint32_t)min((uint64_t)16, b1 - xb) : 0u;
        uint32_t qa[4], qb[4] = {0, 0, 0, 0};
        load16(blob, xa, ca, aligned, qa);
        if (cb) load16(blob, xb, cb, aligned, qb);
        uint32_t bada, badb = 0;
        packed[xa >> 4] = pack16(qa, ca, bada);
        if (cb) packed[xb >> 4] = pack16(qb, cb, badb);
        while (bada) { // rare: flag the owning sequences
            const int k = __ffs(bada) - 1;
            bada &= bada - 1;
            flag_owner(xa + (uint64_t)k, seq_off, seq_len, s_lo, s_hi, seq_flag);
        }
        while (badb) {
            const int k = __ffs(badb) - 1;
            badb &= badb - 1;
            flag_owner(xb + (uint64_t)k, seq_off, seq_len, s_lo, s_hi, seq_flag);
        }
    }
}

// Word j (bases 16j .. 16j+15) of a sequence in the packed stream, bases past the end masked to zero.
__device__ __forceinline__ uint32_t seq_word(const uint32_t *__restrict__ packed, uint64_t off, uint32_t len, uint32_t j) {
    const uint64_t x = off + 16ull * j;
    const uint32_t *w = packed + (x >> 4);
    const uint32_t v = __funnelshift_r(__ldg(w), __ldg(w + 1), (uint32_t)(x & 15) << 1);
    const uint32_t rem = len - 16u * j;
    return rem >= 16u ? v : (v & ((1u << (2u * rem)) - 1u));
}

// Streams of 16-base words of two windows of the packed blob, compared four words per round with all ten loads of a
// round issued before the first use (the loop is latency bound: one thread walks two sequences).  The packed buffer
// is allocated PACK_SLACK words beyond the stream, so reading up to four words past a sequence is safe; bases past
// `n` are masked out of the comparison.
constexpr int PACK_SLACK = 8;
struct Diff16 { // position of the first differing base among the first n of two windows, or n
    __device__ static __forceinline__ uint32_t run(const uint32_t *__restrict__ wa, uint32_t sha, const uint32_t *__restrict__ wb,
                                                   uint32_t shb, uint32_t n) {
        uint32_t a0 = __ldg(wa), b0 = __ldg(wb);
        for (uint32_t base = 0; base < n; base += 64, wa += 4, wb += 4) {
            const uint32_t a1 = __ldg(wa + 1), a2 = __ldg(wa + 2), a3 = __ldg(wa + 3), a4 = __ldg(wa + 4);
            const uint32_t b1 = __ldg(wb + 1), b2 = __ldg(wb + 2), b3 = __ldg(wb + 3), b4 = __ldg(wb + 4);
            const uint32_t x0 = __funnelshift_r(a0, a1, sha) ^ __funnelshift_r(b0, b1, shb);
            const uint32_t x1 = __funnelshift_r(a1, a2, sha) ^ __funnelshift_r(b1, b2, shb);
            const uint32_t x2 = __funnelshift_r(a2, a3, sha) ^ __funnelshift_r(b2, b3, shb);
            const uint32_t x3 = __funnelshift_r(a3, a4, sha) ^ __funnelshift_r(b3, b4, shb);
            if (x0 | x1 | x2 | x3) {
                const uint32_t w = x0 ? 0u : (x1 ? 1u : (x2 ? 2u : 3u));
                const uint32_t x = x0 ? x0 : (x1 ? x1 : (x2 ? x2 : x3));
                return min(n, base + 16u * w + (uint32_t)((__ffs(x) - 1) >> 1));
            }
            a0 = a4, b0 = b4;
        }
        return n;
    }
};
__device__ __forceinline__ bool same_content(const uint32_t *__restrict__ packed, uint64_t oa, uint64_t ob, uint32_t len) {
    return Diff16::run(packed + (oa >> 4), (uint32_t)(oa & 15) << 1, packed + (ob >> 4), (uint32_t)(ob & 15) << 1, len) >= len;
}

// 64-bit content hash per sequence (feeds the duplicate elimination below) + the blob layout contract
// (sequences in index order, non-overlapping, inside the blob: what makes chunked uploads possible).
// One thread per sequence: neighbouring threads walk neighbouring stretches of the packed stream, so a
// warp's loads stay inside a few cache lines that L1 keeps across the ~10 iterations.
__global__ void __launch_bounds__(256) k_hash(const uint64_t *__restrict__ seq_off, const uint32_t *__restrict__ seq_len,
                                              uint32_t s0, uint32_t s1, uint32_t n_seqs, uint64_t blob_bytes,
                                              const uint32_t *__restrict__ packed, uint8_t *__restrict__ seq_flag,
                                              uint64_t *__restrict__ seq_hash, unsigned *__restrict__ err) {
    const uint32_t i = s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s1) return;
    const uint32_t len = seq_len[i];
    const uint64_t off = seq_off[i];
    const bool outside = off > blob_bytes || (uint64_t)len > blob_bytes - off;
    bool bad_layout = outside;
    if (!outside && i + 1 < n_seqs) bad_layout = off + len > seq_off[i + 1];
    if (bad_layout) atomicOr(err, (unsigned)ERR_LAYOUT);
    if (outside) { // 2: unusable, pairs touching it are skipped
        seq_flag[i] = 2, seq_hash[i] = 0;
        return;
    }
    const uint32_t nw = (len + 15u) / 16u;
    const uint32_t *w = packed + (off >> 4);
    const uint32_t sh = (uint32_t)(off & 15) << 1;
    // two independent 32-bit multiplicative chains over the words (order dependent), mixed to 64 bits at the end: four
    // integer instructions per 16 bases instead of two 64-bit multiplies (every table hit is verified anyway)
    uint32_t h1 = 0x811c9dc5u, h2 = len;
    uint32_t lo = nw ? __ldg(w) : 0u;
    for (uint32_t j = 0; j < nw; ++j) {
        const uint32_t hi = __ldg(w + j + 1);
        uint32_t v = __funnelshift_r(lo, hi, sh);
        lo = hi;
        const uint32_t rem = len - 16u * j;
        if (rem < 16u) v &= (1u << (2u * rem)) - 1u;
        h1 = (h1 ^ v) * 0x01000193u;
        h2 = (h2 + v) * 0x9e3779b1u + (h2 >> 15);
    }
    const uint64_t h = mix64((((uint64_t)h1 << 32) | h2) ^ ((uint64_t)len * 0x9e3779b97f4a7c15ull));
    seq_hash[i] = h == HT_EMPTY ? 0 : h;
}

// packed input may hand the sequence lengths over as uint16 (tdb_packed_seqs.seq_len16): widened here, next to the data
__global__ void __launch_bounds__(256) k_widen_len(const uint16_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// Pair lists are runs -- "reads r .. r+c-1 of one allele against target t" is how align_alleleset /
// align_allele (src/denovo.rs:22-54) visit them -- so a host batch ships them run-length encoded
// (12 bytes per run instead of 8 bytes per pair) and this kernel expands them next to the data.
struct PairRun {
    uint32_t begin; // first pair of the run, relative to the chunk
    uint32_t pattern0, text;
};
__global__ void __launch_bounds__(256) k_expand_runs(const PairRun *__restrict__ runs, uint32_t n_runs, uint32_t p0,
                                                     uint32_t n_pairs, uint32_t *__restrict__ pair_p,
                                                     uint32_t *__restrict__ pair_t) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pairs) return;
    uint32_t lo = 0, hi = n_runs; // last run with begin <= j
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (runs[mid].begin <= j) lo = mid;
        else hi = mid;
    }
    const PairRun r = runs[lo];
    pair_p[p0 + j] = r.pattern0 + (j - r.begin);
    pair_t[p0 + j] = r.text;
}

// ------------------------------------------------------------------------------------------------
// Duplicate elimination.  HiFi reads that span a short repeat are mostly error-free copies of the
// allele they come from, so a locus' pair list holds the same (pattern, text) CONTENT many times
// (synthetic 30x trio: 3.5 pairs per distinct content).  The alignment is a pure function of the two
// byte strings, so each distinct content is aligned once and its result copied.
//
// Copies are LOCAL: pairs arrive locus by locus (src/denovo.rs:22-54 visits the reads of one allele set
// against one target) and a content recurs among the ~100 sequences of its locus, not across the genome.
// k_dedup_window therefore works on windows of DW_PAIRS consecutive pairs, one CTA each, with both tables
// in shared memory (round 1 used two global tables larger than L2: 11.5 ms of random probes per 937 k loci):
//   1. the sequences the window references -- a contiguous index range -- go into a table keyed by their
//      64-bit content hash; a hit is VERIFIED word by word (a collision only costs a missed merge, never
//      a wrong one) -> canonical sequence (local index);
//   2. pairs -> representative pair by the exact key (canon pattern, canon text);
//   3. every representative is classified on the spot: pattern == text, closed form, or the tier /
//      work-list bucket it goes to (|n - m|, the dominant term of the penalty, heaviest first);
//   4. after the tiers ran, k_scatter copies score/status/penalty from each representative to its copies.
// A copy whose twin sits in another window is aligned twice (a missed merge, not an error); a window whose
// pairs reference a wider sequence range than the table holds runs without merging.
// ------------------------------------------------------------------------------------------------
// sort keys of representative pairs: 0 quad tier by descending |n-m| ... ; 255 = not in the sorted list
#ifndef TDB_QL_MAX
#define TDB_QL_MAX 12
#endif
constexpr int QL_MAX = TDB_QL_MAX; // |n-m| >= QL_MAX: the wavefront outgrows the quad tier's 32 diagonals before the pair ends
                                    // (swept again at the end of round 2, with tier 1 at four pairs per warp: 10 / 12 / 14 -> quad + tier 1 = 19.8 / 19.1 / 19.3 ms)
constexpr int WL_MAX = 40;  // |n-m| >= WL_MAX: wider than the warp tier's 64 diagonals
constexpr int KEY_WARP0 = 0; // sorted first: heaviest pairs of the shared-memory tiers, by descending |n-m| (the two pairs
                             // of a lockstep warp run to the larger of their penalties: neighbours should be alike)
constexpr int KEY_QUAD0 = KEY_WARP0 + WL_MAX;
constexpr int ML_MAX = 11;      // |n-m| < ML_MAX and both sequences >= ML_MINLEN bases: thread-per-pair tier first
constexpr int ML_MINLEN = 16;
#ifndef TDB_TH_SEQW
#define TDB_TH_SEQW 21
#endif
constexpr int TH_SEQW = TDB_TH_SEQW;            // 2-bit words staged per sequence by the thread-per-pair tier
constexpr int ML_MAXLEN = 16 * TH_SEQW - 32;    // its longest sequence: shift (<= 15) + length + 16 bases fit the window
constexpr int KEY_MINI0 = KEY_QUAD0 + QL_MAX; // sorted last: the lightest pairs
constexpr int KEY_NONE = 255;
constexpr int N_HIST = 72;
static_assert(KEY_MINI0 + ML_MAX <= N_HIST, "histogram size");

// Closed form for the commonest non-identical contents (about two thirds of the aligned pairs of the 30x
// trio, and the heaviest ones): one substitution, or one gap of L <= CF_MAXGAP bases, everything else matching.
// With the production penalties (8,4,2,24,1):
//  * the penalty is 8 resp. 4+2L, which is optimal: two sequences whose lengths differ by L need gaps of net
//    length L, a single first-piece gap is the cheapest way (<= the second piece 24+L for L <= 20), and any
//    further edit costs more;
//  * the path with that penalty is unique: the gap / substitution starts where diagonal 0 stops matching,
//    because wavefront M[0] only keeps its furthest point;
//  * wf-adaptive cannot remove it: at score 4+2j the gap chain sits on the OUTERMOST diagonal of M (+j for an
//    insertion, -j for a deletion; the second-piece chain only overtakes it beyond score 44), and the cut-off
//    never drops the top diagonal while it is <= k_end nor the bottom one while it is >= k_end (SURVEY A.3:
//    the scans stop at min(k_end, hi) resp. max(k_end, lo)).  For L > 20 the optimum runs through the
//    second-piece chain, which is an INNER diagonal for its first 20 steps and can be cut -- measured: the
//    closed form then disagrees with the wavefronts on 1.4 % of such pairs, so it is not used there.
// The clipped score follows from the column positions (src/aligner.rs:320-357).  Checked against the CPU
// checker on > 3*10^5 pairs (tests/test_gpu_parity.py, profiles/closed_form_check.py).  Not used when exact
// work counting is on: C/E/T of a pair are only known by running the wavefronts.
constexpr int CF_MAXGAP = 20;
struct ClosedFormCfg {
    int enabled, clip;
};
__device__ inline bool closed_form(const uint32_t *__restrict__ packed, uint64_t poff, int n, uint64_t toff, int m,
                                   int clip, int &pen, int &clipped) {
    const int L = abs(m - n);
    if (L > CF_MAXGAP) return false;
    const uint32_t *pw = packed + (poff >> 4), *tw = packed + (toff >> 4);
    const int psh = (int)(poff & 15), tsh = (int)(toff & 15);
    const int lim = min(n, m);
    const int p = (int)Diff16::run(pw, (uint32_t)psh << 1, tw, (uint32_t)tsh << 1, (uint32_t)lim); // common prefix
    int v0, h0;
    if (L == 0) {
        if (p >= n) { // same bytes under two canonical ids (duplicate elimination off)
            pen = 0, clipped = 0;
            return true;
        }
        v0 = h0 = p + 1;
    } else if (m > n) {
        v0 = p, h0 = p + L;
    } else {
        v0 = p + L, h0 = p;
    }
    const int room = n - v0; // == m - h0
    if (room > 0) { // everything behind the gap / substitution must match
        const int pa = psh + v0, ta = tsh + h0;
        if ((int)Diff16::run(pw + (pa >> 4), (uint32_t)(pa & 15) << 1, tw + (ta >> 4), (uint32_t)(ta & 15) << 1, (uint32_t)room) < room)
            return false;
    }
    const int cols = max(n, m), wb = clip, we = cols - clip;
    if (L == 0) {
        pen = 8;
        clipped = (p >= wb && p < we) ? -8 : 0;
    } else {
        pen = 4 + 2 * L;
        const int len = min(p + L, we) - max(p, wb);
        clipped = len > 0 ? -min(4 + 2 * len, 24 + len) : 0;
    }
    return true;
}

struct ClassifyOut {
    uint32_t *rep;       // [n_pairs] representative pair of j (itself when unique / identical)
    uint8_t *key;        // [p1-p0] work-list bucket of the pair (KEY_NONE: not aligned)
    uint32_t *glist;     // pairs routed straight to the global-scratch tiers
    unsigned *gcount;
    unsigned *hist;      // [N_HIST]
    unsigned *n_closed;  // pairs resolved by closed_form
    int32_t *out_score, *out_status, *out_penalty;
    uint32_t *work;      // [3 n_pairs] algorithmic cells, ext16, columns per pair (nullptr: not counted)
};

#ifndef TDB_DW_THREADS
#define TDB_DW_THREADS 512
#endif
constexpr int DW_THREADS = TDB_DW_THREADS; // 512 threads x 4 CTAs (52 KB each) = 64 warps per SM: the kernel waits on global loads
constexpr int DW_PAIRS = 2048;   // pairs per window
constexpr int DW_SEQS = 4096;    // widest sequence index range a window may reference
constexpr int DW_SSLOTS = 8192;  // sequence table slots
constexpr int DW_PSLOTS = 4096;  // pair table slots
struct DedupSmem {
    unsigned short stab[DW_SSLOTS]; // local sequence index + 1 of the slot's owner, 0 = empty
    unsigned short canon[DW_SEQS];  // canonical local index of sequence s_lo + i
    uint32_t pkey[DW_PSLOTS];       // canon pattern << 12 | canon text, ~0 = empty
    unsigned short pval[DW_PSLOTS]; // local pair index of the slot's owner
    unsigned short pslot[DW_PAIRS]; // slot of pair jl (DW_PSLOTS: pattern == text, DW_PSLOTS + 1: unusable)
    unsigned hist[N_HIST + 1];
    uint32_t s_lo, s_hi;
    int bad;
    unsigned n_cand; // representatives left for pass B of the classification
};

__global__ void __launch_bounds__(DW_THREADS, 2048 / DW_THREADS) k_dedup_window(
    const uint32_t *__restrict__ pair_p, const uint32_t *__restrict__ pair_t, uint32_t p0, uint32_t p1, uint32_t n_seqs,
    const uint64_t *__restrict__ seq_hash, const uint32_t *__restrict__ seq_len, const uint8_t *__restrict__ seq_flag,
    const uint64_t *__restrict__ seq_off, const uint32_t *__restrict__ packed, int dedup, int quad_ok, int mini_ok,
    int max_len16, ClosedFormCfg cf, ClassifyOut o, unsigned *__restrict__ err) {
    extern __shared__ __align__(16) uint8_t dw_raw[];
    DedupSmem &sm = *reinterpret_cast<DedupSmem *>(dw_raw);
    const uint32_t w0 = p0 + blockIdx.x * (uint32_t)DW_PAIRS, w1 = min(p1, w0 + (uint32_t)DW_PAIRS);
    const int tid = threadIdx.x;
    // ---- the window's sequence range ----
    if (tid == 0) sm.s_lo = 0xffffffffu, sm.s_hi = 0, sm.bad = 0, sm.n_cand = 0;
    for (int i = tid; i <= N_HIST; i += DW_THREADS) sm.hist[i] = 0;
    for (int i = tid; i < DW_SSLOTS; i += DW_THREADS) sm.stab[i] = 0;
    for (int i = tid; i < DW_PSLOTS; i += DW_THREADS) sm.pkey[i] = 0xffffffffu;
    __syncthreads();
    {
        uint32_t lo = 0xffffffffu, hi = 0;
        bool bad = false;
        _Pragma("unroll 4") for (uint32_t j = w0 + tid; j < w1; j += DW_THREADS) {
            const uint32_t ip = pair_p[j], jt = pair_t[j];
            if (ip >= n_seqs || jt >= n_seqs) bad = true;
            else lo = min(lo, min(ip, jt)), hi = max(hi, max(ip, jt));
        }
        lo = __reduce_min_sync(TDB_FULL, lo), hi = __reduce_max_sync(TDB_FULL, hi);
        const bool any_bad = __any_sync(TDB_FULL, bad);
        if ((tid & 31) == 0) {
            atomicMin(&sm.s_lo, lo), atomicMax(&sm.s_hi, hi);
            if (any_bad) sm.bad = 1;
        }
    }
    __syncthreads();
    if (sm.bad && tid == 0) atomicOr(err, (unsigned)ERR_PAIR_INDEX);
    const uint32_t s_lo = sm.s_lo, s_hi = sm.s_hi;
    const uint32_t n_loc = s_hi >= s_lo ? s_hi - s_lo + 1u : 0u;
    const bool local = dedup && n_loc <= (uint32_t)DW_SEQS; // else: no merging in this window
    // ---- sequences -> canonical sequence ----
    if (local) {
        for (uint32_t i = tid; i < n_loc; i += DW_THREADS) {
            const uint32_t g = s_lo + i;
            uint32_t c = i;
            if (!seq_flag[g]) { // sequences with bytes outside ACGT (or outside the blob) are never merged
                const uint64_t h = seq_hash[g];
                const uint32_t len = seq_len[g];
                uint32_t slot = (uint32_t)h & (DW_SSLOTS - 1);
                for (;;) {
                    const unsigned short old = atomicCAS(&sm.stab[slot], (unsigned short)0, (unsigned short)(i + 1));
                    if (old == 0) break; // first of its content: canonical
                    const uint32_t cand = (uint32_t)old - 1u, gc = s_lo + cand;
                    if (seq_hash[gc] == h && seq_len[gc] == len && same_content(packed, seq_off[g], seq_off[gc], len)) {
                        c = cand;
                        break;
                    }
                    slot = (slot + 1) & (DW_SSLOTS - 1);
                }
            }
            sm.canon[i] = (unsigned short)c;
        }
    }
    __syncthreads();
    // ---- pairs -> slot of their (canon pattern, canon text) key ----
    _Pragma("unroll 4") for (uint32_t j = w0 + tid; j < w1; j += DW_THREADS) {
        const uint32_t jl = j - w0;
        const uint32_t ip = pair_p[j], jt = pair_t[j];
        unsigned short ps = (unsigned short)(DW_PSLOTS + 1);
        if (ip < n_seqs && jt < n_seqs && !((seq_flag[ip] | seq_flag[jt]) & 2)) {
            if (!local) {
                ps = ip == jt ? (unsigned short)DW_PSLOTS : (unsigned short)(DW_PSLOTS + 2); // + 2: its own representative
            } else {
                const uint32_t cp = sm.canon[ip - s_lo], ct = sm.canon[jt - s_lo];
                if (cp == ct) {
                    ps = (unsigned short)DW_PSLOTS;
                } else {
                    const uint32_t key = cp << 12 | ct;
                    uint32_t slot = (key * 0x9e3779b1u) >> 20; // 12 bits
                    for (;;) {
                        const uint32_t old = atomicCAS(&sm.pkey[slot], 0xffffffffu, key);
                        if (old == 0xffffffffu) sm.pval[slot] = (unsigned short)jl; // the winner is the representative
                        if (old == 0xffffffffu || old == key) break;
                        slot = (slot + 1) & (DW_PSLOTS - 1);
                    }
                    ps = (unsigned short)slot;
                }
            }
        }
        sm.pslot[jl] = ps;
    }
    __syncthreads();
    // ---- representatives: identical / closed form / tier and bucket ----
    // Two passes.  A: every pair gets its representative; pattern == text is settled on the spot; the representatives
    // that are left -- one pair in six -- are COMPACTED into a candidate list (it reuses the pair-key table, dead by
    // now).  B: the threads walk that list, so the closed form's sequence scans run with full warps (done in place,
    // under the one-in-six predicate, they were half of the kernel's instructions at 1-4 active lanes).
    unsigned short *cand = reinterpret_cast<unsigned short *>(sm.pkey);
    _Pragma("unroll 4") for (uint32_t j = w0 + tid; j < w1; j += DW_THREADS) {
        const uint32_t jl = j - w0;
        const unsigned ps = sm.pslot[jl];
        uint32_t rep = j;
        if (ps == (unsigned)DW_PSLOTS + 1u) {
            // bad pair index (ERR_PAIR_INDEX) or a sequence outside the blob (ERR_LAYOUT, raised by k_hash): the call
            // fails; such pairs never reach an alignment kernel
            o.out_score[j] = TDB_SCORE_FAILED;
            if (o.out_status) o.out_status[j] = TDB_STATUS_UNATTAINABLE;
            if (o.out_penalty) o.out_penalty[j] = INT32_MIN;
            if (o.work) o.work[3 * (size_t)j] = o.work[3 * (size_t)j + 1] = o.work[3 * (size_t)j + 2] = 0;
        } else if (ps == (unsigned)DW_PSLOTS) { // same bytes: penalty 0, every column a match, clipped score 0
            o.out_score[j] = 0;
            if (o.out_status) o.out_status[j] = TDB_STATUS_ALG_COMPLETED;
            if (o.out_penalty) o.out_penalty[j] = 0;
            if (o.work) {
                const int m = (int)seq_len[pair_t[j]];
                o.work[3 * (size_t)j] = 0, o.work[3 * (size_t)j + 1] = (uint32_t)(m >> 4) + 1u;
                o.work[3 * (size_t)j + 2] = (uint32_t)m;
            }
        } else {
            if (ps < (unsigned)DW_PSLOTS) rep = w0 + sm.pval[ps];
            if (rep == j) cand[atomicAdd(&sm.n_cand, 1u)] = (unsigned short)jl;
        }
        o.rep[j] = rep;
        o.key[j - p0] = (uint8_t)KEY_NONE; // pass B gives the candidates that need an alignment their bucket
    }
    __syncthreads();
    const uint32_t n_cand = sm.n_cand;
    for (uint32_t c = tid; c < n_cand; c += DW_THREADS) {
        const uint32_t j = w0 + cand[c];
        const uint32_t ip = pair_p[j], jt = pair_t[j];
        const int n = (int)seq_len[ip], m = (int)seq_len[jt];
        const int L = abs(m - n);
        const unsigned fl = seq_flag[ip] | seq_flag[jt];
        int key = KEY_NONE, pen = 0, clipped = 0;
        if (cf.enabled && !fl && closed_form(packed, seq_off[ip], n, seq_off[jt], m, cf.clip, pen, clipped)) {
            o.out_score[j] = clipped;
            if (o.out_status) o.out_status[j] = TDB_STATUS_ALG_COMPLETED;
            if (o.out_penalty) o.out_penalty[j] = pen;
            atomicAdd(sm.hist + N_HIST, 1u);
        } else if (fl || n > max_len16 || m > max_len16 || L >= WL_MAX) {
            const unsigned q = atomicAdd(o.gcount, 1u);
            o.glist[q] = j;
        } else if (!quad_ok || L >= QL_MAX) {
            key = KEY_WARP0 + (WL_MAX - 1 - L);
        } else if (mini_ok && L < ML_MAX && n >= ML_MINLEN && m >= ML_MINLEN && n <= ML_MAXLEN && m <= ML_MAXLEN) {
            key = KEY_MINI0 + (ML_MAX - 1 - L);
        } else {
            key = KEY_QUAD0 + (QL_MAX - 1 - L);
        }
        if (key != KEY_NONE) {
            o.key[j - p0] = (uint8_t)key;
            atomicAdd(sm.hist + key, 1u);
        }
    }
    __syncthreads();
    if (tid < N_HIST && sm.hist[tid]) atomicAdd(o.hist + tid, sm.hist[tid]);
    if (tid == N_HIST && sm.hist[N_HIST]) atomicAdd(o.n_closed, sm.hist[N_HIST]);
}

// Segment table of the work list: seg[0] = begin, seg[1] = count of the warp-tier segment; seg[2], seg[3] the same for the
// lockstep (quad) tier; seg[4], seg[5] for the thread-per-pair tier.  base[key] = where bucket `key` starts in the list
// (buckets in key order: heaviest first inside every tier).
__global__ void k_plan(const unsigned *__restrict__ hist, unsigned *__restrict__ seg, unsigned *__restrict__ base) {
    if (threadIdx.x == 0) {
        unsigned w = 0, q = 0, t = 0, at = 0;
        for (int i = 0; i < N_HIST; ++i) base[i] = at, at += hist[i];
        for (int i = KEY_WARP0; i < KEY_WARP0 + WL_MAX; ++i) w += hist[i];
        for (int i = KEY_QUAD0; i < KEY_QUAD0 + QL_MAX; ++i) q += hist[i];
        for (int i = KEY_MINI0; i < KEY_MINI0 + ML_MAX; ++i) t += hist[i];
        seg[0] = 0, seg[1] = w;
        seg[2] = w, seg[3] = q;
        seg[4] = w + q, seg[5] = t;
    }
}

// The work list: every pair that needs an alignment goes to its bucket (a counting sort of the few per cent of pairs
// that have a key; the histogram came out of k_pair_classify).  A block ranks its pairs per bucket in shared memory and
// reserves one range per bucket; order inside a bucket follows the pair order up to the block granularity, which keeps
// the pairs of a locus -- alike in length and penalty -- next to each other.
constexpr int BUCKET_PER_THREAD = 4;
__global__ void __launch_bounds__(256) k_bucket(const uint8_t *__restrict__ key, uint32_t p0, uint32_t n,
                                                const unsigned *__restrict__ base, unsigned *__restrict__ cursor,
                                                uint32_t *__restrict__ list) {
    __shared__ unsigned cnt[N_HIST], boff[N_HIST];
    if (threadIdx.x < N_HIST) cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t first = (blockIdx.x * 256u + threadIdx.x) * BUCKET_PER_THREAD;
    int k[BUCKET_PER_THREAD];
    unsigned rank[BUCKET_PER_THREAD];
#pragma unroll
    for (int i = 0; i < BUCKET_PER_THREAD; ++i) {
        const uint32_t j = first + (uint32_t)i;
        k[i] = j < n ? (int)key[j] : KEY_NONE;
        if (k[i] != KEY_NONE) rank[i] = atomicAdd(&cnt[k[i]], 1u);
    }
    __syncthreads();
    if (threadIdx.x < N_HIST && cnt[threadIdx.x]) boff[threadIdx.x] = base[threadIdx.x] + atomicAdd(&cursor[threadIdx.x], cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < BUCKET_PER_THREAD; ++i)
        if (k[i] != KEY_NONE) list[boff[k[i]] + rank[i]] = p0 + first + (uint32_t)i;
}

// ------------------------------------------------------------------------------------------------
// Work order of the thread-per-pair tier.  Its 32 pairs of a warp run to the largest penalty among them, and
// at a given |n - m| the penalties still spread (12 .. 30 and beyond: a second or third sequencing error,
// stutter): sorted by |n - m| alone 58 % of the lane-steps of a warp did useful work (CPU study with the
// CPU checker's cell counts, profiles/r02/NOTES.md).  k_estimate gives every pair of that tier an UPPER BOUND of
// its penalty from a greedy walk -- extend the matches, at a mismatch take the edit (substitution, or a gap of
// up to |n - m| + 2 bases either way) whose next 16 bases match best, repeat -- which is the cost of a real
// alignment, so never below the optimum, and equal to it for 86 % of the pairs.  Pairs bounded by the tier's
// last score are sorted by that bound, heaviest first (90 % useful lane-steps in the same study) and are
// certain to finish there; the others go straight to the lockstep tier instead of running to score 30 first.
// The walks are a kernel of their own over the tier's segment of the work list (k_estimate, 0.65 ms per 937 k loci), followed
// by a second counting sort of that segment; run at the end of k_dedup_window instead (the window's few candidates compacted
// in shared memory) they stretched every window by a third (+1.8 ms): one walk is a long dependent chain.
// Only the ORDER and the tier of a pair depend on the estimate: results are those of the wavefront kernels.
// ------------------------------------------------------------------------------------------------
constexpr int EST_BUCKETS = 16;
#ifndef TDB_EST_SLACK
#define TDB_EST_SLACK 8
#endif
constexpr int EST_SLACK = TDB_EST_SLACK;
__device__ __forceinline__ int est_gap_cost(int g) { return min(4 + 2 * g, 24 + g); } // production penalties 8,4,2,24,1
__device__ inline int greedy_estimate(const uint32_t *__restrict__ pw, int psh, int n, const uint32_t *__restrict__ tw, int tsh,
                                      int m, int cap) {
    int v = 0, h = 0, est = 0;
    const int G = min(abs(m - n), 13) + 2;
    for (int ev = 0; ev < 10; ++ev) {
        const int room = min(n - v, m - h);
        if (room > 0) {
            const int pa = psh + v, ta = tsh + h;
            const int e = (int)Diff16::run(pw + (pa >> 4), (uint32_t)(pa & 15) << 1, tw + (ta >> 4), (uint32_t)(ta & 15) << 1, (uint32_t)room);
            v += e, h += e;
        }
        if (v >= n || h >= m) {
            const int rest = (n - v) + (m - h);
            return rest ? est + est_gap_cost(rest) : est;
        }
        if (est > cap + EST_SLACK) return est; // the bound only grows
        // the next 32 bases of both sequences in one 64-bit register each: every candidate below is a shift of one of them
        // against the other, so an edit is scored without further loads (gaps up to 15 bases: 16 bases of look-ahead)
        uint64_t p64, t64;
        {
            const int pa = psh + v, ta = tsh + h;
            const uint32_t *a = pw + (pa >> 4), *c = tw + (ta >> 4);
            const uint32_t a0 = __ldg(a), a1 = __ldg(a + 1), a2 = __ldg(a + 2), c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2);
            const uint32_t sa = (uint32_t)(pa & 15) << 1, sc = (uint32_t)(ta & 15) << 1;
            p64 = (uint64_t)__funnelshift_r(a0, a1, sa) | ((uint64_t)__funnelshift_r(a1, a2, sa) << 32);
            t64 = (uint64_t)__funnelshift_r(c0, c1, sc) | ((uint64_t)__funnelshift_r(c1, c2, sc) << 32);
        }
        int best_key = INT32_MIN, bc = 0, bv = v, bh = h;
#define TDB_EST_CAND(cost_, dv_, dh_)                                                                             \
    {                                                                                                             \
        const int v2 = v + (dv_), h2 = h + (dh_);                                                                 \
        if (v2 <= n && h2 <= m) {                                                                                 \
            const int room2 = min(n - v2, m - h2);                                                                \
            const uint32_t x = (uint32_t)(p64 >> (2 * (dv_))) ^ (uint32_t)(t64 >> (2 * (dh_)));                    \
            const int r = min(x ? (__ffs(x) - 1) >> 1 : 16, room2);                                               \
            const int key = 4 * r - (cost_) + ((v2 + r >= n && h2 + r >= m) ? 400 : 0);                           \
            if (key > best_key) best_key = key, bc = (cost_), bv = v2, bh = h2;                                   \
        }                                                                                                         \
    }
        TDB_EST_CAND(8, 1, 1)
        for (int g = 1; g <= G; ++g) {
            TDB_EST_CAND(4 + 2 * g, 0, g)
            TDB_EST_CAND(4 + 2 * g, g, 0)
        }
#undef TDB_EST_CAND
        est += bc, v = bv, h = bh;
    }
    return 1 << 20; // too many edits for this tier
}

__global__ void __launch_bounds__(256) k_estimate(const uint32_t *__restrict__ packed, const uint64_t *__restrict__ seq_off,
                                                  const uint32_t *__restrict__ seq_len, const uint32_t *__restrict__ pair_p,
                                                  const uint32_t *__restrict__ pair_t, const uint32_t *__restrict__ list,
                                                  const unsigned *__restrict__ seg, int smax, uint8_t *__restrict__ ekey,
                                                  unsigned *__restrict__ hist2, uint32_t *__restrict__ over_list,
                                                  unsigned *__restrict__ over_count) {
    __shared__ unsigned sh[EST_BUCKETS];
    if (threadIdx.x < EST_BUCKETS) sh[threadIdx.x] = 0;
    __syncthreads();
    const unsigned b = seg[0], c = seg[1];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < c; i += gridDim.x * blockDim.x) {
        const uint32_t idx = list[b + i];
        const uint32_t ip = pair_p[idx], jt = pair_t[idx];
        const uint64_t poff = seq_off[ip], toff = seq_off[jt];
        const int est = greedy_estimate(packed + (poff >> 4), (int)(poff & 15), (int)seq_len[ip], packed + (toff >> 4),
                                        (int)(toff & 15), (int)seq_len[jt], smax);
        int key = KEY_NONE;
        // a bound a little beyond the last score still leaves an even chance that the pair ends below it (the walk pays for
        // an edit it could have merged): such pairs ride with the heaviest bucket, whose warps run to the last score anyway
        if (est <= smax + EST_SLACK) {
            key = min(EST_BUCKETS - 1, max(smax - est, 0) >> 1); // heaviest first
            atomicAdd(&sh[key], 1u);
        } else {
            const unsigned q = atomicAdd(over_count, 1u);
            over_list[q] = idx;
        }
        ekey[i] = (uint8_t)key;
    }
    __syncthreads();
    if (threadIdx.x < EST_BUCKETS && sh[threadIdx.x]) atomicAdd(hist2 + threadIdx.x, sh[threadIdx.x]);
}

__global__ void k_plan2(const unsigned *__restrict__ hist2, unsigned *__restrict__ base2, unsigned *__restrict__ seg2) {
    if (threadIdx.x == 0) {
        unsigned at = 0;
        for (int i = 0; i < EST_BUCKETS; ++i) base2[i] = at, at += hist2[i];
        seg2[0] = 0, seg2[1] = at;
    }
}

// k_bucket for a list: position i of the segment goes to bucket ekey[i]
__global__ void __launch_bounds__(256) k_bucket_list(const uint8_t *__restrict__ ekey, const uint32_t *__restrict__ list,
                                                     const unsigned *__restrict__ seg, const unsigned *__restrict__ base2,
                                                     unsigned *__restrict__ cursor2, uint32_t *__restrict__ out) {
    __shared__ unsigned cnt[EST_BUCKETS], boff[EST_BUCKETS];
    const unsigned b = seg[0], n = seg[1];
    for (uint32_t blk = blockIdx.x; blk * (256u * BUCKET_PER_THREAD) < n; blk += gridDim.x) { // block-uniform trip count
        if (threadIdx.x < EST_BUCKETS) cnt[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t first = (blk * 256u + threadIdx.x) * BUCKET_PER_THREAD;
        int k[BUCKET_PER_THREAD];
        unsigned rank[BUCKET_PER_THREAD];
#pragma unroll
        for (int i = 0; i < BUCKET_PER_THREAD; ++i) {
            const uint32_t j = first + (uint32_t)i;
            k[i] = j < n ? (int)ekey[j] : KEY_NONE;
            if (k[i] != KEY_NONE) rank[i] = atomicAdd(&cnt[k[i]], 1u);
        }
        __syncthreads();
        if (threadIdx.x < EST_BUCKETS && cnt[threadIdx.x])
            boff[threadIdx.x] = base2[threadIdx.x] + atomicAdd(&cursor2[threadIdx.x], cnt[threadIdx.x]);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < BUCKET_PER_THREAD; ++i)
            if (k[i] != KEY_NONE) out[boff[k[i]] + rank[i]] = list[b + first + (uint32_t)i];
        __syncthreads();
    }
}

// The long tier's list, heaviest pairs first.  Its pairs take milliseconds each and arrive in the order the windows and the
// earlier tiers happened to append them; a warp that draws one of the longest pairs at the very end of the launch keeps
// the whole launch waiting.  One CTA sorts the (few tens of thousands of) entries by |n - m| / 64 -- the length of the gap
// is what sets the number of score steps -- with a counting sort in shared memory, descending.
constexpr int LS_BUCKETS = 256;
__global__ void __launch_bounds__(1024) k_sort_long(const uint32_t *__restrict__ in, unsigned n, const uint32_t *__restrict__ pair_p,
                                                    const uint32_t *__restrict__ pair_t, const uint32_t *__restrict__ seq_len,
                                                    uint32_t *__restrict__ out) {
    __shared__ unsigned cnt[LS_BUCKETS], base[LS_BUCKETS];
    for (int i = threadIdx.x; i < LS_BUCKETS; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    auto bucket = [&](uint32_t idx) {
        const int d = abs((int)seq_len[pair_t[idx]] - (int)seq_len[pair_p[idx]]) >> 6;
        return LS_BUCKETS - 1 - min(d, LS_BUCKETS - 1); // heaviest first
    };
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&cnt[bucket(in[i])], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned at = 0;
        for (int b = 0; b < LS_BUCKETS; ++b) base[b] = at, at += cnt[b];
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t idx = in[i];
        out[atomicAdd(&base[bucket(idx)], 1u)] = idx;
    }
}

__global__ void __launch_bounds__(256) k_scatter(const uint32_t *__restrict__ rep, uint32_t n_pairs,
                                                 int32_t *__restrict__ score, int32_t *__restrict__ status,
                                                 int32_t *__restrict__ penalty, const uint32_t *__restrict__ work,
                                                 unsigned long long *__restrict__ algo,
                                                 const unsigned *__restrict__ n_failed_reps) {
    unsigned long long c = 0, e = 0, t = 0, u = 0, f = 0;
    const bool any_failed = *n_failed_reps != 0; // rare: only then are the failed pairs (duplicates included) counted
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_pairs; j += gridDim.x * blockDim.x) {
        const uint32_t r = rep[j];
        if (r != j) {
            score[j] = score[r];
            if (status) status[j] = status[r];
            if (penalty) penalty[j] = penalty[r];
        } else {
            ++u;
        }
        if (any_failed && score[r] == TDB_SCORE_FAILED) ++f;
        if (work) c += work[3 * (size_t)r], e += work[3 * (size_t)r + 1], t += work[3 * (size_t)r + 2];
    }
    for (int o = 16; o; o >>= 1) {
        c += __shfl_xor_sync(TDB_FULL, c, o), e += __shfl_xor_sync(TDB_FULL, e, o);
        t += __shfl_xor_sync(TDB_FULL, t, o), u += __shfl_xor_sync(TDB_FULL, u, o);
        f += __shfl_xor_sync(TDB_FULL, f, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (work) atomicAdd(algo + 0, c), atomicAdd(algo + 1, e), atomicAdd(algo + 2, t);
        atomicAdd(algo + 3, u);
        if (f) atomicAdd(algo + 4, f);
    }
}

// ------------------------------------------------------------------------------------------------
// Alignment tiers.  Persistent warps pull pairs from a device-side counter (work per pair spans three
// orders of magnitude, so static assignment would leave SMs idle).  A tier's items come from up to
// two device-resident lists: a segment of the sorted representative list and the pairs the previous
// tier could not finish.  All counts live on the device, so a batch is enqueued without host syncs.
// ------------------------------------------------------------------------------------------------
struct AlignParams {
    Penalties pn;
    Heur heur;
    int clip;
    const uint32_t *packed;
    const uint8_t *seq_flag;
    const uint8_t *blob;
    const uint64_t *seq_off;
    const uint32_t *seq_len;
    const uint32_t *pair_p;
    const uint32_t *pair_t;
    const uint32_t *src1;      // first item list (nullptr: identity 0..src1_count-1)
    const unsigned *src1_seg;  // device {begin, count} of the segment of src1 (nullptr: 0, src1_count)
    unsigned src1_count;
    const uint32_t *src2;      // second item list (overflow of the previous tier); may be nullptr
    const unsigned *src2_count;
    unsigned *work_counter;
    uint32_t *next_todo; // pairs this tier could not finish (nullptr on the last tier)
    unsigned *next_count;
    int32_t *out_score;
    int32_t *out_status;
    int32_t *out_penalty;
    uint32_t *work; // per-pair algorithmic work (cells, ext16, columns); may be nullptr
    uint8_t *cigar_buf;
    const uint64_t *cigar_off;
    uint32_t *cigar_len;
    unsigned long long *counters; // executed cells, ext16, columns, pairs finished by this launch
    unsigned *fail_count;         // pairs the LAST tier could not finish (score = TDB_SCORE_FAILED, status OOM)
    uint8_t *scratch;             // global tiers: per-warp slices
    size_t scratch_stride;
    int wlog2, s_cap, ops_cap;
    unsigned prov_cap;
};

struct ItemSrc {
    unsigned b1, c2, total;
};
// src2 (what the previous tier could not finish: the heaviest pairs) is served before src1
__device__ __forceinline__ ItemSrc item_src(const AlignParams &p) {
    ItemSrc s;
    s.b1 = p.src1_seg ? p.src1_seg[0] : 0u;
    s.c2 = p.src2 ? *p.src2_count : 0u;
    s.total = s.c2 + (p.src1_seg ? p.src1_seg[1] : p.src1_count);
    return s;
}
__device__ __forceinline__ uint32_t item_at(const AlignParams &p, const ItemSrc &s, unsigned it) {
    if (it < s.c2) return p.src2[it];
    it -= s.c2;
    return p.src1 ? p.src1[s.b1 + it] : it;
}

constexpr int T0_WLOG2 = 6;
constexpr int T0_W = 1 << T0_WLOG2;
constexpr int T0_SCAP = 128;
constexpr int T0_PROV = 2048;
constexpr int T0_WARPS = 4;
constexpr int T0_MAXLEN = 8000; // int16 offsets
constexpr int WORK_CHUNK = 4;

struct __align__(16) T0Smem {
    int16_t M[MS * T0_W];
    int16_t G[4 * GS * T0_W];
    int2 mmeta[MS];
    int2 gmeta[4 * GS];
    int16_t step_lo[T0_SCAP];
    uint16_t step_base[T0_SCAP];
    uint8_t prov[T0_PROV];
};

__device__ __forceinline__ void flush_counts(const AlignParams &p, const WorkCount &wc,
                                             unsigned long long lane_ext16, unsigned long long done, int lane) {
    for (int o = 16; o; o >>= 1) lane_ext16 += __shfl_xor_sync(TDB_FULL, lane_ext16, o);
    if (lane == 0 && p.counters) {
        atomicAdd(p.counters + 0, wc.cells);
        atomicAdd(p.counters + 1, lane_ext16);
        atomicAdd(p.counters + 2, wc.columns);
        atomicAdd(p.counters + 3, done);
    }
}

__device__ __forceinline__ void defer_pair(const AlignParams &p, uint32_t idx, int lane) {
    if (lane == 0) {
        if (p.next_todo) {
            const unsigned j = atomicAdd(p.next_count, 1u);
            p.next_todo[j] = idx;
        } else {
            // no tier left: the pair has NO score.  The sentinel keeps it out of every reduction (k_reduce skips
            // it, so the lists shrink exactly as src/denovo.rs:31-35 drops non-completed alignments)
            p.out_score[idx] = TDB_SCORE_FAILED;
            if (p.out_status) p.out_status[idx] = TDB_STATUS_OOM;
            if (p.out_penalty) p.out_penalty[idx] = INT32_MIN;
            if (p.fail_count) atomicAdd(p.fail_count, 1u);
            if (p.cigar_len) p.cigar_len[idx] = 0;
            if (p.work) p.work[3 * (size_t)idx] = p.work[3 * (size_t)idx + 1] = p.work[3 * (size_t)idx + 2] = 0;
        }
    }
}

// result of one finished pair (whole warp calls; ext1 is a per-lane partial count)
__device__ __forceinline__ void finish_pair(const AlignParams &p, uint32_t idx, int lane, int clipped, int pen,
                                            const WorkCount &wc1, unsigned long long ext1) {
    unsigned e = 0;
    if (p.work) e = __reduce_add_sync(TDB_FULL, (unsigned)ext1);
    if (lane == 0) {
        p.out_score[idx] = clipped;
        if (p.out_status) p.out_status[idx] = TDB_STATUS_ALG_COMPLETED;
        if (p.out_penalty) p.out_penalty[idx] = pen;
        if (p.work) {
            p.work[3 * (size_t)idx] = (uint32_t)wc1.cells, p.work[3 * (size_t)idx + 1] = e;
            p.work[3 * (size_t)idx + 2] = (uint32_t)wc1.columns;
        }
    }
}

// Warp tier: one pair per warp, wavefront history in shared memory.
__global__ void __launch_bounds__(T0_WARPS * 32) k_align_warp(const AlignParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T0Smem &sm = reinterpret_cast<T0Smem *>(smem_raw)[wid];
    Storage<int16_t, int16_t, uint16_t> st;
    st.M = sm.M, st.G = sm.G, st.mmeta = sm.mmeta, st.gmeta = sm.gmeta;
    st.step_lo = sm.step_lo, st.step_base = sm.step_base, st.prov = sm.prov;
    st.ops = reinterpret_cast<uint8_t *>(sm.M);
    st.wlog2 = T0_WLOG2, st.s_cap = T0_SCAP, st.ops_cap = (int)sizeof(sm.M), st.prov_cap = T0_PROV;
    const ItemSrc src = item_src(p);
    WorkCount wc = {0, 0, 0};
    unsigned long long lane_ext16 = 0, done = 0;
    // pairs per work-counter atomic: one when a launch has only a few pairs per warp (they are few and heavy)
    const unsigned n_warps = gridDim.x * (blockDim.x >> 5);
    const unsigned chunk = src.total >= n_warps * 16u ? (unsigned)WORK_CHUNK : 1u;
    for (;;) {
        unsigned b = 0;
        if (lane == 0) b = atomicAdd(p.work_counter, chunk);
        b = __shfl_sync(TDB_FULL, b, 0);
        if (b >= src.total) break;
        const unsigned e = min(b + chunk, src.total);
        for (unsigned it = b; it < e; ++it) {
            const uint32_t idx = item_at(p, src, it);
            const uint32_t ip = p.pair_p[idx], jt = p.pair_t[idx];
            const int n = (int)p.seq_len[ip], m = (int)p.seq_len[jt];
            if ((p.seq_flag[ip] | p.seq_flag[jt]) || n > T0_MAXLEN || m > T0_MAXLEN) {
                defer_pair(p, idx, lane);
                continue;
            }
            const uint64_t poff = p.seq_off[ip], toff = p.seq_off[jt];
            SeqView P = {p.packed + (poff >> 4), nullptr, n, (int)(poff & 15)};
            SeqView T = {p.packed + (toff >> 4), nullptr, m, (int)(toff & 15)};
            int pen = 0, clipped = 0;
            WorkCount wc1 = {0, 0, 0}; // work of this attempt: only counted when the tier finishes the pair
            unsigned long long ext1 = 0;
            const int r = wfa_align_warp<int16_t, int16_t, uint16_t, false>(
                P, T, p.pn, p.heur, p.clip, st, lane, pen, clipped, wc1, ext1, nullptr, nullptr);
            __syncwarp();
            if (r == WFA_OK) {
                ++done;
                wc.cells += wc1.cells, wc.columns += wc1.columns, lane_ext16 += ext1;
                finish_pair(p, idx, lane, clipped, pen, wc1, ext1);
            } else {
                defer_pair(p, idx, lane);
            }
        }
    }
    flush_counts(p, wc, lane_ext16, done, lane);
}

// Global-scratch tier: int32 offsets, runtime ring width, optional CIGAR output.
__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline size_t global_tier_stride(int wlog2, int s_cap, int ops_cap, unsigned prov_cap) {
    const size_t W = (size_t)1 << wlog2;
    return align16(MS * W * 4) + align16(4 * GS * W * 4) + align16(MS * 8) + align16(4 * GS * 8) +
           align16((size_t)s_cap * 4) * 2 + align16((size_t)ops_cap) + align16((size_t)prov_cap);
}

#ifndef TDB_G_MINB
#define TDB_G_MINB 4
#endif
template <bool WANT_CIGAR>
__global__ void __launch_bounds__(128, TDB_G_MINB) k_align_global(const AlignParams p) {
    const int lane = threadIdx.x & 31;
    const size_t gw = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint8_t *base = p.scratch + gw * p.scratch_stride;
    const size_t W = (size_t)1 << p.wlog2;
    Storage<int32_t, int32_t, uint32_t> st;
    st.M = reinterpret_cast<int32_t *>(base), base += align16(MS * W * 4);
    st.G = reinterpret_cast<int32_t *>(base), base += align16(4 * GS * W * 4);
    st.mmeta = reinterpret_cast<int2 *>(base), base += align16(MS * 8);
    st.gmeta = reinterpret_cast<int2 *>(base), base += align16(4 * GS * 8);
    st.step_lo = reinterpret_cast<int32_t *>(base), base += align16((size_t)p.s_cap * 4);
    st.step_base = reinterpret_cast<uint32_t *>(base), base += align16((size_t)p.s_cap * 4);
    st.ops = base, base += align16((size_t)p.ops_cap);
    st.prov = base;
    st.wlog2 = p.wlog2, st.s_cap = p.s_cap, st.ops_cap = p.ops_cap, st.prov_cap = p.prov_cap;
    const ItemSrc src = item_src(p);
    WorkCount wc = {0, 0, 0};
    unsigned long long lane_ext16 = 0, done = 0;
    for (;;) {
        unsigned it = 0;
        if (lane == 0) it = atomicAdd(p.work_counter, 1u);
        it = __shfl_sync(TDB_FULL, it, 0);
        if (it >= src.total) break;
        const uint32_t idx = item_at(p, src, it);
        const uint32_t ip = p.pair_p[idx], jt = p.pair_t[idx];
        const int n = (int)p.seq_len[ip], m = (int)p.seq_len[jt];
        SeqView P, T;
        P.len = n, T.len = m;
        const uint64_t poff = p.seq_off[ip], toff = p.seq_off[jt];
        P.w = p.packed + (poff >> 4), P.sh = (int)(poff & 15);
        T.w = p.packed + (toff >> 4), T.sh = (int)(toff & 15);
        P.b = p.seq_flag[ip] ? p.blob + poff : nullptr; // raw bytes only where 2 bits lose information
        T.b = p.seq_flag[jt] ? p.blob + toff : nullptr;
        int pen = 0, clipped = 0;
        uint8_t *cig = WANT_CIGAR ? p.cigar_buf + p.cigar_off[idx] : nullptr;
        uint32_t *cl = WANT_CIGAR ? p.cigar_len + idx : nullptr;
        WorkCount wc1 = {0, 0, 0};
        unsigned long long ext1 = 0;
        const int r = wfa_align_warp<int32_t, int32_t, uint32_t, WANT_CIGAR>(P, T, p.pn, p.heur, p.clip, st, lane,
                                                                             pen, clipped, wc1, ext1, cig, cl);
        __syncwarp();
        if (r == WFA_OK) {
            ++done;
            wc.cells += wc1.cells, wc.columns += wc1.columns, lane_ext16 += ext1;
            finish_pair(p, idx, lane, clipped, pen, wc1, ext1);
        } else {
            defer_pair(p, idx, lane);
        }
    }
    flush_counts(p, wc, lane_ext16, done, lane);
}

// ------------------------------------------------------------------------------------------------
// Fused per-locus reduction (src/trio/denovo.rs:83-186, src/duo/denovo.rs:46-119, src/denovo.rs:68-129,
// src/math.rs:82-155).  One thread per locus: the lists are a few dozen scores, the arithmetic is
// integer counting plus a handful of IEEE f64 operations that must round exactly as the Rust code
// does (explicit _rn intrinsics: no FMA contraction).
// ------------------------------------------------------------------------------------------------
// Scores equal to TDB_SCORE_FAILED (pairs no tier could finish) are not part of any list: every helper skips
// them, which is what dropping non-completed alignments does in the reference (src/denovo.rs:31-35).
__device__ __forceinline__ bool sc_ok(int x) { return x != TDB_SCORE_FAILED; }
__device__ inline int n_valid(const int32_t *xs, int n) {
    int c = 0;
#pragma unroll 4
    for (int i = 0; i < n; ++i) c += sc_ok(xs[i]);
    return c;
}
// k-th smallest (0-based) of the valid entries.  Short lists (the 30x case: a few dozen scores): count-and-compare,
// n^2 / 2 operations without a single store.  Long lists (deep coverage: hundreds to thousands of reads per allele): a
// radix select on the order-preserving unsigned image of the scores, one pass per bit -- 32 n instead of n^2.
constexpr int KTH_RADIX_MIN = 64;
__device__ __noinline__ int kth_smallest_radix(const int32_t *xs, int n, int k) { // (kept out of line: the 30x path stays lean)
    uint32_t prefix = 0u, decided = 0u; // bits fixed so far
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t b = 1u << bit;
        int zeros = 0; // valid entries that match the prefix and have this bit clear
        for (int j = 0; j < n; ++j) {
            const uint32_t u = (uint32_t)xs[j] ^ 0x80000000u;
            zeros += (int)(sc_ok(xs[j]) & ((u & decided) == prefix) & ((u & b) == 0u));
        }
        if (k >= zeros) k -= zeros, prefix |= b;
        decided |= b;
    }
    return (int)(prefix ^ 0x80000000u);
}
__device__ inline int kth_smallest(const int32_t *xs, int n, int k) {
    if (n >= KTH_RADIX_MIN) return kth_smallest_radix(xs, n, k);
    for (int i = 0; i < n; ++i) {
        const int x = xs[i];
        if (!sc_ok(x)) continue;
        int less = 0, eq = 0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const int y = xs[j];
            less += (y < x) & sc_ok(y);
            eq += y == x;
        }
        if (less <= k && k < less + eq) return x;
    }
    return 0;
}

// math::quantile on one list (src/math.rs:82-98); nv = valid entries > 0.
__device__ inline double list_quantile(const int32_t *xs, int n, int nv, double q) {
    const double index = __dmul_rn(q, __dsub_rn((double)nv, 1.0));
    const double fl = floor(index);
    const long long lower = fl <= 0.0 ? 0 : (fl >= 9.0e18 ? (long long)9.0e18 : (long long)fl);
    const long long upper = lower + 1;
    if (upper >= (long long)nv) {
        int mx = INT32_MIN; // the sentinel is the smallest int: the maximum of the list ignores it by itself
#pragma unroll 4
        for (int i = 0; i < n; ++i) mx = max(mx, xs[i]);
        return (double)mx;
    }
    const double fraction = __dsub_rn(index, (double)lower);
    const double a = (double)kth_smallest(xs, n, (int)lower), b = (double)kth_smallest(xs, n, (int)upper);
    return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), fraction));
}

// get_top_other_score over lists [first, first+n_lists) (src/denovo.rs:93-101)
__device__ inline bool top_other(const int32_t *scores, const uint32_t *off, int first, int n_lists, double q,
                                 double &out) {
    bool have = false;
    double best = 0.0;
    for (int l = 0; l < n_lists; ++l) {
        const int n = (int)(off[first + l + 1] - off[first + l]);
        const int nv = n_valid(scores + off[first + l], n);
        if (nv <= 0) continue;
        const double v = list_quantile(scores + off[first + l], n, nv, q);
        if (!have || v >= best) best = v, have = true;
    }
    out = best;
    return have;
}

// get_score_count_diff (src/denovo.rs:117-129)
__device__ inline void count_diff(double top, const int32_t *a, int n, int &count, float &mean_diff) {
    int c = 0;
    double sum = 0.0;
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
        const double x = (double)a[i];
        if (sc_ok(a[i]) && x > top) ++c, sum = __dadd_rn(sum, x);
    }
    count = c;
    mean_diff = c > 0 ? (float)fabs(__dsub_rn(__ddiv_rn(sum, (double)c), top)) : 0.0f;
}

__device__ inline int overlap_cov(double thr, const int32_t *a, int n) {
    int c = 0;
#pragma unroll 4
    for (int i = 0; i < n; ++i) c += (sc_ok(a[i]) && (double)a[i] >= thr);
    return c;
}

__device__ inline double list_mean(const int32_t *s, uint32_t b, uint32_t e) {
    uint32_t sum = 0, nv = 0;
#pragma unroll 4
    for (uint32_t i = b; i < e; ++i)
        if (sc_ok(s[i])) sum += (uint32_t)s[i], ++nv; // i32 sum (wrapping)
    if (nv == 0) return -INFINITY;
    return __ddiv_rn((double)(int32_t)sum, (double)nv);
}

__device__ inline double list_median(const int32_t *a, int n_raw) {
    const int n = n_valid(a, n_raw);
    if (n == 0) return 1.7976931348623157e308; // unwrap_or(f64::MAX)
    if (n & 1) return (double)kth_smallest(a, n_raw, n / 2);
    const int x = kth_smallest(a, n_raw, n / 2 - 1), y = kth_smallest(a, n_raw, n / 2);
    return __ddiv_rn((double)(int32_t)((uint32_t)x + (uint32_t)y), 2.0);
}

__constant__ int8_t c_combs2[8][2] = {{0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 0}, {3, 0}, {2, 1}, {3, 1}};

__global__ void __launch_bounds__(128) k_reduce(const tdb_trio_locus *__restrict__ loci, uint32_t n_loci,
                                                const int32_t *__restrict__ scores, double q, int is_duo,
                                                tdb_allele_out *__restrict__ out, uint8_t *__restrict__ mask,
                                                unsigned *__restrict__ hit_count, uint32_t *__restrict__ hit_list,
                                                unsigned hit_cap) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_loci) return;
    const tdb_trio_locus loc = loci[li];
    const double F64_MIN = -1.7976931348623157e308;
    double matrix[4][2];
    for (int i = 0; i < 4; ++i) matrix[i][0] = matrix[i][1] = F64_MIN;
    bool any_err = false;
    tdb_allele_out o[2];
    for (int c = 0; c < loc.n_child && c < 2; ++c) {
        const uint32_t *off = loc.list_off[c];
        tdb_allele_out &r = o[c];
        r.denovo_coverage = 0, r.mean_diff_father = 0.f, r.mean_diff_mother = 0.f;
        r.mother_overlap[0] = r.mother_overlap[1] = r.father_overlap[0] = r.father_overlap[1] = 0;
        r.allele_origin = is_duo ? 5 : 4, r.denovo_status = 0, r.error = 0;
        r.top_mother = r.top_father = r.child_median = 0.0;
        r.mean_col[0] = r.mean_col[1] = r.mean_col[2] = r.mean_col[3] = 0.0;
        const int32_t *own = scores + off[4];
        const int nown = (int)(off[5] - off[4]);
        double top_m = 0.0, top_f = 0.0;
        const bool hm = top_other(scores, off, 0, loc.n_mother, q, top_m);
        const bool hf = is_duo ? true : top_other(scores, off, 2, loc.n_father, q, top_f);
        if (!hm || !hf) {
            r.error = 1, any_err = true;
            continue;
        }
        r.top_mother = top_m, r.top_father = top_f;
        const double thr = is_duo ? top_m : fmax(top_m, top_f);
        if (mask)
            for (int i = 0; i < nown; ++i) {
                const bool hit = sc_ok(own[i]) && (double)own[i] > thr;
                mask[off[4] + i] = hit ? 1 : 0;
                // the set bytes are a few in a million: host calls fetch this list instead of the whole mask
                if (hit && hit_count) {
                    const unsigned q_ = atomicAdd(hit_count, 1u);
                    if (q_ < hit_cap) hit_list[q_] = off[4] + (uint32_t)i;
                }
            }
        int cm = 0, cf = 0;
        count_diff(top_m, own, nown, cm, r.mean_diff_mother);
        if (!is_duo) count_diff(top_f, own, nown, cf, r.mean_diff_father);
        r.denovo_coverage = is_duo ? cm : (top_m >= top_f ? cm : cf);
        const double med = list_median(own, nown);
        r.child_median = med;
        for (int a = 0; a < loc.n_mother && a < 2; ++a)
            r.mother_overlap[a] = overlap_cov(med, scores + off[a], (int)(off[a + 1] - off[a]));
        if (!is_duo) {
            for (int a = 0; a < loc.n_father && a < 2; ++a)
                r.father_overlap[a] = overlap_cov(med, scores + off[2 + a], (int)(off[3 + a] - off[2 + a]));
            matrix[0][c] = list_mean(scores, off[0], off[1]);
            matrix[1][c] = loc.n_mother > 1 ? list_mean(scores, off[1], off[2]) : F64_MIN;
            matrix[2][c] = list_mean(scores, off[2], off[3]);
            matrix[3][c] = loc.n_father > 1 ? list_mean(scores, off[3], off[4]) : F64_MIN;
            for (int i = 0; i < 4; ++i) r.mean_col[i] = matrix[i][c];
        } else {
            r.denovo_status = cm == 0 ? 0 : 4;
        }
    }
    if (!is_duo && !any_err) {
        double cs[8];
        int idx[8], ncomb;
        if (loc.n_child > 1) {
            ncomb = 8;
            for (int i = 0; i < 8; ++i) {
                const double s = __dadd_rn(__dadd_rn(0.0, matrix[c_combs2[i][0]][0]), matrix[c_combs2[i][1]][1]);
                cs[i] = s, idx[i] = i;
            }
        } else {
            ncomb = 4;
            for (int i = 0; i < 4; ++i) cs[i] = matrix[i][0], idx[i] = i;
        }
        for (int i = 1; i < ncomb; ++i) { // stable insertion sort, descending (sort_unstable_by on <= 8 items)
            const double v = cs[i];
            const int vi = idx[i];
            int j = i - 1;
            while (j >= 0 && cs[j] < v) cs[j + 1] = cs[j], idx[j + 1] = idx[j], --j;
            cs[j + 1] = v, idx[j + 1] = vi;
        }
        const double comb_diff = fabs(__dsub_rn(cs[0], cs[1]));
        for (int c = 0; c < loc.n_child && c < 2; ++c) {
            tdb_allele_out &r = o[c];
            if (comb_diff < 1.0) r.allele_origin = 4;
            else r.allele_origin = loc.n_child > 1 ? c_combs2[idx[0]][c] : idx[0];
            if (r.denovo_coverage == 0) r.denovo_status = 0;
            else {
                int plen = -1;
                switch (r.allele_origin) {
                case 0: plen = loc.mother_len[0]; break;
                case 1: plen = loc.mother_len[1]; break;
                case 2: plen = loc.father_len[0]; break;
                case 3: plen = loc.father_len[1]; break;
                default: break;
                }
                r.denovo_status = plen < 0 ? 4 : (plen < loc.child_len[c] ? 1 : (plen > loc.child_len[c] ? 2 : 3));
            }
        }
    }
    for (int c = 0; c < 2; ++c) {
        if (c < loc.n_child) out[2 * (size_t)li + c] = o[c];
    }
}

} // namespace tdb
