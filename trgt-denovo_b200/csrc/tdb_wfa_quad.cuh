// tdb_wfa_quad.cuh -- k_align_lock: several pairs per warp in LOCKSTEP (tier 0: four pairs, eight lanes each).
//
// Why: after duplicate elimination and the closed forms the typical pair still has a penalty around 20 and
// wavefronts 3..16 diagonals wide, so a warp per pair leaves most lanes idle and the kernel is bound by the
// instruction issue of its per-score bookkeeping.  Here a warp runs 32/G consecutive pairs of the work-sorted
// list (same |n-m|, usually the same locus: near-identical control flow) under ONE warp-uniform control flow, so
// that every bookkeeping instruction serves all of them.  Further:
//   * existence of the last 26 M / gap wavefronts is tracked in register bit-masks, so the null score steps of
//     the 2-piece-affine recurrence (most steps, penalties 8/6/25) touch no memory;
//   * the M history lives in a FIFO arena (only non-null wavefronts occupy shared memory), which lets a pair
//     with width <= 32 fit in 3 KB and keeps 16 warps resident per SM; a wavefront never wraps around the arena end
//     and a gap row starts at its score's computed lo, so a cell is read at base + k without a mask and the k-1 / k+1
//     neighbours share one address;
//   * penalties are compile-time constants (the reference hard-wires 8,4,2,24,1: src/commands/shared.rs:39-51,
//     src/model.rs:13-23); other penalties use the general tiers (wfa_align_warp, tdb_wfa.cuh).
// The same kernel serves tier 0 (QuadCfg), tier 1 (Duo4Cfg: provenance in global scratch) and the long tier (LongCfg: one
// pair per warp, thousands of score steps, wavefronts up to 192 diagonals).
// Semantics are exactly those of wfa_align_warp: SURVEY.md Appendix A.1-A.4.
#pragma once
#include <type_traits>
#include "tdb_kernels.cuh"

namespace tdb {

// capacities (overridable at build time for tuning sweeps, profiles/run_gpu_sweep.sh)
#ifndef TDB_Q_MAR
#define TDB_Q_MAR 512
#endif
#ifndef TDB_Q_SCAP
#define TDB_Q_SCAP 96
#endif
#ifndef TDB_Q_PROV
#define TDB_Q_PROV 736
#endif
#ifndef TDB_Q_WARPS
#define TDB_Q_WARPS 2
#endif
#ifndef TDB_Q_MINB
#define TDB_Q_MINB 1
#endif
#ifndef TDB_Q_FETCH
#define TDB_Q_FETCH 2 // warp batches per grab of the work counter (at most)
#endif
// Geometry of a lockstep configuration: G lanes per pair (32/G pairs per warp), wavefronts up to W
// diagonals, M arena of MAR entries (power of two), scores < SCAP, PROV provenance bytes.
// GP: the provenance bytes and the per-score step records -- written once, read only by the backtrace -- live in a
// per-pair slice of global scratch (L2) instead of shared memory, which halves a wide pair's shared-memory state.
// WIDE (the long tier): diagonals beyond +-127 and scores beyond 16 bits of provenance -- the metadata of an M wavefront
// takes two words, the provenance index of a score 32 bits, and the op stream of the backtrace (OPS bytes) lives in the
// global slice too; sequences up to 16 000 bases (offsets stay int16).
template <int G_, int W_, int MAR_, int SCAP_, int PROV_, bool GP_ = false, int MINB_ = TDB_Q_MINB, bool WIDE_ = false,
          int OPS_ = 0>
struct LockCfg {
    static constexpr int G = G_, W = W_, MAR = MAR_, SCAP = SCAP_, PROV = PROV_, MINB = MINB_, OPS = OPS_;
    static constexpr bool GP = GP_, WIDE = WIDE_;
    static_assert(!WIDE_ || GP_, "the long tier keeps provenance, step records and ops in global scratch");
    using StepBase = typename std::conditional<WIDE_, uint32_t, uint16_t>::type;
};
#ifndef TDB_Q_GP
#define TDB_Q_GP 0
#endif
#ifndef TDB_Q_GPROV
#define TDB_Q_GPROV 1024
#endif
#ifndef TDB_Q_GMINB
#define TDB_Q_GMINB 10
#endif
#if TDB_Q_GP
// tier 0: four pairs per warp; provenance and step records in global scratch, 1.8 KB of shared memory per pair, and a
// register cap that lets 20 warps stay resident per SM instead of 16 (the kernel waits on dependent-issue latency)
using QuadCfg = LockCfg<8, 32, TDB_Q_MAR, TDB_Q_SCAP, TDB_Q_GPROV, true, TDB_Q_GMINB>;
#else
using QuadCfg = LockCfg<8, 32, TDB_Q_MAR, TDB_Q_SCAP, TDB_Q_PROV>; // tier 0: four pairs per warp
#endif
// tier 0 with EIGHT pairs per warp (4 lanes each, TDB_QUAD8=1): the per-score bookkeeping serves twice as many pairs; the
// provenance bytes and step records in global scratch keep a warp's eight pairs at 14.7 KB of shared memory
#ifndef TDB_Q8_MINB
#define TDB_Q8_MINB 7
#endif
using Quad8Cfg = LockCfg<4, 32, TDB_Q_MAR, TDB_Q_SCAP, TDB_Q_GPROV, true, TDB_Q8_MINB>;
#ifndef TDB_DUO_G
#define TDB_DUO_G 16
#endif
using DuoCfg = LockCfg<TDB_DUO_G, 64, 1024, 128, 2048>;             // tier 1: two pairs per warp
// tier 1 with FOUR pairs per warp (8 lanes each): the per-score bookkeeping of the lockstep loop serves twice as many
// pairs, and 8 lanes fit the ~20 diagonals of these wavefronts better than 16 (3 x 8 instead of 2 x 16 lane slots).  What
// makes it fit: provenance and step records in global scratch (GP), 3.7 KB of shared memory per pair instead of 6.2.
#ifndef TDB_DUO4_PROV
#define TDB_DUO4_PROV 4096
#endif
using Duo4Cfg = LockCfg<8, 64, 1024, 128, TDB_DUO4_PROV, true, 16 / TDB_Q_WARPS>; // 16 warps per SM: 128 registers
// the long tier: ONE pair per warp (32 lanes), wavefronts up to 192 diagonals, scores < 32768, sequences up to 16 000
// bases.  What k_align_global did for the long-expansion stress set from a global-memory ring with run-time penalties
// (1 600 warp instructions and two round trips to L2 per score step), this does from 13 KB of shared memory per pair.
#ifndef TDB_LONG_PROV
#define TDB_LONG_PROV (4 << 20)
#endif
#ifndef TDB_LONG_W
#define TDB_LONG_W 192
#endif
#ifndef TDB_LONG_MAR
#define TDB_LONG_MAR 4736
#endif
using LongCfg = LockCfg<32, TDB_LONG_W, TDB_LONG_MAR, 32768, TDB_LONG_PROV, true, 16 / TDB_Q_WARPS, true, 49152>;
constexpr int Q_MAXLEN = 8000;
constexpr int Q_MAXLEN_WIDE = 16000;
constexpr int Q_WARPS = TDB_Q_WARPS;
constexpr int Q_NULL = -16384;

// Metadata of M[s] in one word: valid range lo..hi (int8 each: |k| < 128 for every geometry below) and the arena index
// of diagonal lo (free-running 16 bits, masked on access) -> one 4-byte shared-memory load per input wavefront.
// The ranges of the gap wavefronts never leave registers: only the last two (I1, D1) / one (I2, D2) scores are read.
__device__ __forceinline__ uint32_t qm_pack(int lo, int hi, unsigned off) {
    return ((uint32_t)lo & 0xffu) | (((uint32_t)hi & 0xffu) << 8) | (off << 16);
}
__device__ __forceinline__ int qm_lo(uint32_t w) { return (int)(int8_t)(w & 0xffu); }
__device__ __forceinline__ int qm_hi(uint32_t w) { return (int)(int8_t)((w >> 8) & 0xffu); }
__device__ __forceinline__ unsigned qm_off(uint32_t w) { return w >> 16; }
// a range in one register: lo in the low half, hi in the high half (int16 each); empty when lo > hi
__device__ __forceinline__ uint32_t qr_pack(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ int qr_lo(uint32_t r) { return (int)(int16_t)(r & 0xffffu); }
__device__ __forceinline__ int qr_hi(uint32_t r) { return (int)r >> 16; }
template <class C>
struct __align__(16) QPairT {
    int16_t marena[C::MAR];       // FIFO arena of M wavefronts; reused as the op stream at backtrace
    int16_t gap[10][C::W];        // I1[3], D1[3] (scores s, s-1, s-2: row s mod 3), I2[2], D2[2] (row s & 1)
    uint32_t mmeta[C::WIDE ? 64 : 32]; // qm_pack(valid range, arena position) of M[s], slot s & 31 (WIDE: qr_pack(range), position)
    int16_t step_lo[C::GP ? 8 : C::SCAP];     // computed lo of score s (provenance row origin)
    typename C::StepBase step_base[C::GP ? 8 : C::SCAP];
    uint8_t prov[C::GP ? 16 : C::PROV];
};
template <class C>
__host__ __device__ constexpr size_t lock_gscratch_bytes() { // per pair: provenance | step_lo | step_base | ops
    return C::GP ? (((size_t)C::PROV + 15) & ~(size_t)15) + (size_t)C::SCAP * (2 + sizeof(typename C::StepBase)) +
                       (((size_t)C::OPS + 15) & ~(size_t)15)
                 : 0;
}
static_assert(sizeof(QPairT<QuadCfg>) % 16 == 0 && sizeof(QPairT<DuoCfg>) % 16 == 0 && sizeof(QPairT<Duo4Cfg>) % 16 == 0 &&
                  sizeof(QPairT<LongCfg>) % 16 == 0 && LongCfg::SCAP % 8 == 0,
              "pair state keeps 16-byte alignment");
template <class C>
constexpr size_t lock_smem_bytes() { return sizeof(QPairT<C>) * (32 / C::G) * Q_WARPS; }

struct QIn { // one input wavefront, group-uniform
    const int16_t *p; // element for diagonal k at p[(base + k) & amask]
    int base;         // arena: off - lo ; ring: 0 with k & (QW-1)
    int lo;
    unsigned span1;
};

// All 32 lanes vote; a group reads its own byte.  Every collective in this kernel is full-warp: the four
// groups run in LOCKSTEP -- one warp-uniform control flow (same score s, same loop trip counts, taken as
// the maximum over the groups) with per-group predicates -- so that every issued instruction serves four
// pairs.  (Group-masked collectives made the groups diverge for good and issue one after the other:
// ncu r01c showed 8 active threads on every vote.)
template <int QG>
__device__ __forceinline__ unsigned qballot(int gshift, bool pred) {
    return (__ballot_sync(TDB_FULL, pred) >> gshift) & ((QG == 32) ? 0xffffffffu : ((1u << QG) - 1u));
}

// The lanes of a group extend one diagonal: lane l compares bases [16l, 16l+16) per round.  `on` is
// group-uniform; groups that are off (or done) idle through the remaining rounds.
template <int QG>
__device__ __forceinline__ int extend_groups(const uint32_t *pw, const uint32_t *tw, int psh, int tsh, int n, int m,
                                             int v, int h, bool on, int gl, int gshift, int lane) {
    const int room = on ? min(n - v, m - h) : 0;
    int res = room;
    bool busy = on && room > 0;
    for (int base = 0; __any_sync(TDB_FULL, busy); base += 16 * QG) {
        const int off = base + gl * 16;
        uint32_t x = 1u; // lanes past the end report a mismatch at their first base: caps the run at `room`
        if (busy && off < room) x = get16(pw, psh + v + off) ^ get16(tw, tsh + h + off);
        const unsigned bal = qballot<QG>(gshift, x != 0);
        const int l = bal ? __ffs(bal) - 1 : 0;
        const uint32_t xl = __shfl_sync(TDB_FULL, x, (lane & ~(QG - 1)) + l);
        if (busy && bal) {
            res = min(base + l * 16 + ((__ffs(xl) - 1) >> 1), room);
            busy = false;
        }
    }
    return res;
}

template <class C, int X, int O1, int E1, int O2, int E2>
__global__ void __launch_bounds__(Q_WARPS * 32, C::MINB) k_align_lock(const AlignParams p) {
    static_assert(X < 32 && O1 + E1 < 32 && O2 + E2 < 32 && E1 == 2 && E2 == 1, "ring sizes; gap ranges are kept for two / one score(s)");
    constexpr int QG = C::G, QW = C::W, Q_MAR = C::MAR, Q_SCAP = C::SCAP, Q_PROV = C::PROV;
    using QPair = QPairT<C>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gl = lane & (QG - 1), gid = lane / QG, gshift = gid * QG;
    QPair &sm = reinterpret_cast<QPair *>(smem_raw)[wid * (32 / QG) + gid];
    using StepBase = typename C::StepBase;
    constexpr bool WIDE = C::WIDE;
    constexpr int MAXLEN = WIDE ? Q_MAXLEN_WIDE : Q_MAXLEN;
    uint8_t *prov = sm.prov;
    int16_t *step_lo = sm.step_lo;
    StepBase *step_base = sm.step_base;
    uint8_t *ops = reinterpret_cast<uint8_t *>(sm.marena); // the arena is dead by the time of the backtrace
    if (C::GP) { // this group's slice of the launch's global scratch
        constexpr size_t GBYTES = lock_gscratch_bytes<C>();
        uint8_t *g = p.scratch + ((size_t)(blockIdx.x * Q_WARPS + wid) * (32 / QG) + gid) * GBYTES;
        prov = g;
        step_lo = reinterpret_cast<int16_t *>(g + (((size_t)Q_PROV + 15) & ~(size_t)15));
        step_base = reinterpret_cast<StepBase *>(step_lo + Q_SCAP);
        if (WIDE) ops = reinterpret_cast<uint8_t *>(step_base + Q_SCAP);
    }
    const Penalties pn = {X, O1, E1, O2, E2};
    const ItemSrc src = item_src(p);
    const unsigned n_items = src.total;
    // pairs per work-counter atomic: 8 batches when there is plenty of work, fewer when a launch (a pipeline
    // chunk) only has a few batches per warp -- coarse grabs then leave warps idle behind the heaviest ones
    constexpr int PER_BATCH = 32 / QG, MAX_FETCH = TDB_Q_FETCH * PER_BATCH;
    const unsigned n_warps = gridDim.x * (blockDim.x >> 5);
    const unsigned per_fetch = WIDE ? 1u : // (the long tier's pairs are few and take milliseconds each)
        (unsigned)PER_BATCH * max(1u, min((unsigned)(MAX_FETCH / PER_BATCH), n_items / (n_warps * 8u * (unsigned)PER_BATCH)));
    unsigned long long acc_cells = 0, acc_cols = 0, acc_ext = 0, acc_done = 0; // acc_* valid on group lane 0

    for (;;) {
        unsigned fb = 0;
        if (lane == 0) fb = atomicAdd(p.work_counter, per_fetch);
        fb = __shfl_sync(TDB_FULL, fb, 0);
        if (fb >= n_items) break;
        for (unsigned bb = fb; bb < fb + per_fetch && bb < n_items; bb += 32 / QG) {
            // ====================== one pair per group, four groups in lockstep ======================
            const unsigned it = bb + gid;
            const bool have = it < n_items;
            uint32_t idx = 0;
            int n = 0, m = 0, psh = 0, tsh = 0;
            const uint32_t *pw = p.packed, *tw = p.packed;
            bool overflow = false;
            if (have) {
                idx = item_at(p, src, it);
                const uint32_t ip = p.pair_p[idx], jt = p.pair_t[idx];
                n = (int)p.seq_len[ip], m = (int)p.seq_len[jt];
                overflow = (p.seq_flag[ip] | p.seq_flag[jt]) || n > MAXLEN || m > MAXLEN;
                const uint64_t poff = p.seq_off[ip], toff = p.seq_off[jt];
                pw = p.packed + (poff >> 4), tw = p.packed + (toff >> 4);
                psh = (int)(poff & 15), tsh = (int)(toff & 15);
            }
            const int k_end = m - n;
            const bool usable = have && !overflow;
            int penalty = 0, clipped = 0;
            unsigned cells = 0, cols = 0, ext = 0; // per pair: far below 2^32
            const int m0 = extend_groups<QG>(pw, tw, psh, tsh, n, m, 0, 0, usable, gl, gshift, lane);
            if (usable && gl == 0) ext += (unsigned)(m0 >> 4) + 1;
            const bool identical = usable && k_end == 0 && m0 >= m;
            if (identical) cols = (unsigned)m; // penalty 0, clipped score 0
            bool running = usable && !identical;
            // ---------------- wavefront loop (score s is warp-uniform) ----------------
            if (running && gl == 0) {
                if (WIDE) sm.mmeta[0] = qr_pack(0, 0), sm.mmeta[1] = 0u;
                else sm.mmeta[0] = qm_pack(0, 0, 0u);
                sm.marena[0] = (int16_t)m0;
                step_lo[0] = 0, step_base[0] = 0;
                prov[0] = 0;
            }
            unsigned Mm = 1u, I1m = 0, I2m = 0, D1m = 0, D2m = 0; // bit j <-> score s-j exists
            // ranges of the gap wavefronts, in registers: I1 / D1 of scores s-1 (a) and s-2 (b), I2 / D2 of score s-1
            uint32_t rI1a = 0, rI1b = 0, rD1a = 0, rD1b = 0, rI2 = 0, rD2 = 0;
            unsigned mhead = 1, prov_used = 1; // mhead: arena position behind the newest M wavefront
            int loA = 0, loB = 0;              // computed lo of scores s-1 / s-2: origin of their gap rows
            int steps_wait = p.heur.steps - (p.heur.kind == 1 ? 1 : 0);
            int s = 0, s3 = 0; // s3 = s mod 3
            __syncwarp();
            while (__any_sync(TDB_FULL, running)) {
                ++s;
                s3 = s3 == 2 ? 0 : s3 + 1;
                if (s >= Q_SCAP) { // warp-uniform
                    if (running) overflow = true, running = false;
                    break;
                }
                Mm <<= 1, I1m <<= 1, I2m <<= 1, D1m <<= 1, D2m <<= 1;
                bool live = running && ((Mm & ((1u << X) | (1u << (O1 + E1)) | (1u << (O2 + E2)))) |
                                        ((I1m | D1m) & (1u << E1)) | ((I2m | D2m) & (1u << E2))) != 0;
                if (!__any_sync(TDB_FULL, live)) { // null step for every group: registers only
                    rI1b = rI1a, rD1b = rD1a, loB = loA;
                    continue;
                }
                // ---- inputs: three M wavefronts (arena, metadata in one word each), four gap wavefronts (rings by score,
                // ranges in registers).  An absent input has span 0: every read of it yields NULL. ----
                int mx_lo = 1, mx_base = 0, mo1_lo = 1, mo1_base = 0, mo2_lo = 1, mo2_base = 0;
                unsigned mx_sp = 0, mo1_sp = 0, mo2_sp = 0;
#define TDB_QM(w, lag)                                                                \
    if (live && (Mm & (1u << (lag)))) {                                               \
        if (WIDE) {                                                                   \
            const uint2 mm_ = reinterpret_cast<const uint2 *>(sm.mmeta)[(s - (lag)) & 31]; \
            w##_lo = qr_lo(mm_.x), w##_sp = (unsigned)(qr_hi(mm_.x) - w##_lo + 1), w##_base = (int)mm_.y - w##_lo; \
        } else {                                                                      \
            const uint32_t mm_ = sm.mmeta[(s - (lag)) & 31];                          \
            w##_lo = qm_lo(mm_), w##_sp = (unsigned)(qm_hi(mm_) - w##_lo + 1), w##_base = (int)qm_off(mm_) - w##_lo; \
        }                                                                             \
    }
                TDB_QM(mx, X)
                TDB_QM(mo1, O1 + E1)
                TDB_QM(mo2, O2 + E2)
#undef TDB_QM
                int i1_lo = 1, d1_lo = 1, i2_lo = 1, d2_lo = 1;
                unsigned i1_sp = 0, d1_sp = 0, i2_sp = 0, d2_sp = 0;
                if (live && (I1m & (1u << E1))) i1_lo = qr_lo(rI1b), i1_sp = (unsigned)(qr_hi(rI1b) - i1_lo + 1);
                if (live && (D1m & (1u << E1))) d1_lo = qr_lo(rD1b), d1_sp = (unsigned)(qr_hi(rD1b) - d1_lo + 1);
                if (live && (I2m & (1u << E2))) i2_lo = qr_lo(rI2), i2_sp = (unsigned)(qr_hi(rI2) - i2_lo + 1);
                if (live && (D2m & (1u << E2))) d2_lo = qr_lo(rD2), d2_sp = (unsigned)(qr_hi(rD2) - d2_lo + 1);
                const int s3r = s3 == 2 ? 0 : s3 + 1; // (s - 2) mod 3
                const int16_t *I1r = sm.gap[0 + s3r], *D1r = sm.gap[3 + s3r]; // warp-uniform rows
                const int16_t *I2r = sm.gap[6 + ((s - E2) & 1)], *D2r = sm.gap[8 + ((s - E2) & 1)];
                int lo = INT32_MAX, hi = INT32_MIN;
#define TDB_LIM(w, dl, dh)                                              \
    if (w##_sp) {                                                       \
        lo = min(lo, w##_lo + (dl));                                    \
        hi = max(hi, w##_lo + (int)w##_sp - 1 + (dh));                  \
    }
                TDB_LIM(mx, 0, 0)
                TDB_LIM(mo1, -1, 1)
                TDB_LIM(mo2, -1, 1)
                TDB_LIM(i1, 1, 1)
                TDB_LIM(i2, 1, 1)
                TDB_LIM(d1, -1, -1)
                TDB_LIM(d2, -1, -1)
#undef TDB_LIM
                int width = 0;
                unsigned mpos = mhead; // arena position of this score's M wavefront (diagonal lo)
                if (live) {
                    lo = max(lo, -n - 1), hi = min(hi, m + 1);
                    width = hi - lo + 1;
                    // Arena room.  The live wavefronts (lags 1..25) occupy the arena from the oldest one's position (tail) up
                    // to mhead, circularly; a wavefront never wraps around the end of the arena (it moves to position 0
                    // instead), so a cell is read at base + k without a mask and k-1 / k+1 share one address.
                    const unsigned livem = Mm & 0x03fffffeu; // lags 1..25
                    bool room = true;
                    if (!livem) {
                        mpos = 0; // nothing to keep: start over
                    } else {
                        const int ts = (s - (31 - __clz(livem))) & 31;
                        const unsigned tail = WIDE ? sm.mmeta[2 * ts + 1] : qm_off(sm.mmeta[ts]);
                        if (mhead > tail) { // live region [tail, mhead): the space behind it, else the space in front
                            if (mhead + (unsigned)width > (unsigned)Q_MAR) mpos = 0, room = (unsigned)width < tail;
                        } else {            // live region wraps: [tail, MAR) and [0, mhead)
                            room = mhead + (unsigned)width < tail;
                        }
                    }
                    if (width > QW || prov_used + (unsigned)width > Q_PROV || !room) {
                        overflow = true, running = false, live = false, width = 0; // this group drops out
                    }
                }
                if (!live) lo = 0, hi = -1;
                int16_t *I1o = sm.gap[0 + s3], *D1o = sm.gap[3 + s3];
                int16_t *I2o = sm.gap[6 + (s & 1)], *D2o = sm.gap[8 + (s & 1)];
                uint8_t *pv = prov + prov_used;
                cells += (unsigned)width;
                // Which cells lie inside the DP matrix decides the trimmed range of each output wavefront (A.2).  M: one vote
                // per iteration, first / last set bit.  Gap components: per lane and component one word of two halfwords,
                // (0xffff - smallest, 1 + largest) diagonal (relative to lo) that holds a valid offset -- both grow, so one
                // VIMNMX.U16x2 per component and iteration, and one halfword butterfly over the group per score.  dpk: the
                // same for the distance to the end over the valid M cells, (max d, 0xffff - min d), for the cut-off below.
                // The second gap piece (I2, D2) does not exist below score O2 + E2: its reads, stores and bookkeeping are
                // skipped (warp-uniform) while no group has a source for it.
                const bool any2 = __any_sync(TDB_FULL, live && ((Mm & (1u << (O2 + E2))) | ((I2m | D2m) & (1u << E2))) != 0);
                int tl_m = INT32_MAX, th_m = INT32_MIN;
                uint32_t wI1 = 0u, wI2 = 0u, wD1 = 0u, wD2 = 0u, dpk = 0u;
                bool fin = false;
                const int n_iter = __reduce_max_sync(TDB_FULL, (width + QG - 1) / QG);
                for (int itk = 0; itk < n_iter; ++itk) {
                    const int kb = lo + itk * QG;
                    const int k = kb + gl;
                    const bool act = k <= hi; // false everywhere in groups that are not live (hi = -1 < lo = 0)
                    int ins1 = Q_NULL, ins2 = Q_NULL, del1 = Q_NULL, del2 = Q_NULL, mv = Q_NULL;
                    bool vm = false;
                    int ev = 0, eroom = 0; // extension state of this lane's diagonal
                    if (act) {
// M wavefronts: arena position of diagonal k is base + k; gap rows: a score's row starts at its computed lo (loA: score
// s-1, loB: score s-2).  dk = -1 / 0 / +1: the neighbours of a cell share one address.
#define TDB_RDM(w, dk) (((unsigned)(k + (dk) - w##_lo) < w##_sp) ? (int)sm.marena[w##_base + k + (dk)] : Q_NULL)
#define TDB_RDG(w, row, org, dk) (((unsigned)(k + (dk) - w##_lo) < w##_sp) ? (int)(row)[k - (org) + (dk)] : Q_NULL)
                        int b = 0, o, e;
                        o = TDB_RDM(mo1, -1), e = TDB_RDG(i1, I1r, loB, -1);
                        if (e >= o) ins1 = e, b |= PB_I1E; else ins1 = o;
                        ins1 += 1;
                        o = TDB_RDM(mo1, +1), e = TDB_RDG(d1, D1r, loB, +1);
                        if (e >= o) del1 = e, b |= PB_D1E; else del1 = o;
                        if (any2) {
                            o = TDB_RDM(mo2, -1), e = TDB_RDG(i2, I2r, loA, -1);
                            if (e >= o) ins2 = e, b |= PB_I2E; else ins2 = o;
                            ins2 += 1;
                            o = TDB_RDM(mo2, +1), e = TDB_RDG(d2, D2r, loA, +1);
                            if (e >= o) del2 = e, b |= PB_D2E; else del2 = o;
                        } else {
                            b |= PB_I2E | PB_D2E; // what two absent sources give (NULL >= NULL): the provenance byte is unchanged
                            ins2 = Q_NULL + 1;
                        }
                        const int mis = TDB_RDM(mx, 0) + 1;
#undef TDB_RDM
#undef TDB_RDG
                        mv = max(max(max(ins1, ins2), max(del1, del2)), mis);
                        int srcm = PM_NONE; // later tests override: priority X > D2 > D1 > I2 > I1
                        if (mv == ins1) srcm = PM_I1;
                        if (mv == ins2) srcm = PM_I2;
                        if (mv == del1) srcm = PM_D1;
                        if (mv == del2) srcm = PM_D2;
                        if (mv == mis) srcm = PM_X;
                        pv[k - lo] = (uint8_t)(b | srcm);
                        if ((unsigned)mv > (unsigned)m || (unsigned)(mv - k) > (unsigned)n) mv = Q_NULL;
                        vm = mv >= 0;
                        if (vm) ev = mv - k, eroom = min(n - ev, m - mv);
                    }
                    // extend at once (trimming never changes offsets); one warp-uniform loop for all lanes
                    // (running word pointers, the upper word of a round is the lower word of the next: two loads per round)
                    int eq = 0;
                    bool ebusy = vm && eroom > 0;
                    const uint32_t *pq = pw + ((psh + ev) >> 4), *tq = tw + ((tsh + mv) >> 4);
                    const int pqs = ((psh + ev) & 15) << 1, tqs = ((tsh + mv) & 15) << 1;
                    uint32_t plo = 0, tlo = 0;
                    if (ebusy) plo = __ldg(pq), tlo = __ldg(tq);
                    while (__any_sync(TDB_FULL, ebusy)) {
                        if (ebusy) {
                            ++pq, ++tq;
                            const uint32_t phi = __ldg(pq), thi = __ldg(tq);
                            const uint32_t x = __funnelshift_r(plo, phi, pqs) ^ __funnelshift_r(tlo, thi, tqs);
                            plo = phi, tlo = thi;
                            if (x) eq += (__ffs(x) - 1) >> 1, ebusy = false;
                            else eq += 16, ebusy = eq < eroom;
                        }
                    }
                    if (act) {
                        if (vm) {
                            eq = min(eq, eroom);
                            ext += (unsigned)(eq >> 4) + 1;
                            mv += eq;
                            const unsigned d = (unsigned)max(n - (mv - k), m - mv);
                            dpk = __vmaxu2(dpk, (d << 16) | (0xffffu - d));
                            fin |= (k == k_end) && (mv >= m);
                        }
                        sm.marena[(int)mpos + (k - lo)] = (int16_t)mv;
                        // a component without a source only holds NULL+1 here: never valid, so its range stays empty
                        I1o[k - lo] = (int16_t)ins1;
                        D1o[k - lo] = (int16_t)del1;
#define TDB_VALID(val) ((unsigned)(val) <= (unsigned)m && (unsigned)((val)-k) <= (unsigned)n)
                        const uint32_t rel = (uint32_t)(k - lo);                        // 0..QW-1
                        const uint32_t cand = ((0xffffu - rel) << 16) | (rel + 1u);    // (0xffff - first, last + 1)
                        wI1 = __vmaxu2(wI1, TDB_VALID(ins1) ? cand : 0u);
                        wD1 = __vmaxu2(wD1, TDB_VALID(del1) ? cand : 0u);
                        if (any2) { // (without a source the rows keep stale values; they are never read: the range stays empty)
                            I2o[k - lo] = (int16_t)ins2;
                            D2o[k - lo] = (int16_t)del2;
                            wI2 = __vmaxu2(wI2, TDB_VALID(ins2) ? cand : 0u);
                            wD2 = __vmaxu2(wD2, TDB_VALID(del2) ? cand : 0u);
                        }
#undef TDB_VALID
                    }
                    const unsigned balm = qballot<QG>(gshift, vm);
                    if (balm) tl_m = min(tl_m, kb + __ffs(balm) - 1), th_m = kb + 31 - __clz(balm);
                }
                // trim_ends (A.2): one halfword butterfly over the group gives the first / last valid diagonal of the gap
                // components; an empty component (word 0) ends up with first = lo + 0xffff > last = lo - 1
#pragma unroll
                for (int o = 1; o < QG; o <<= 1) {
                    wI1 = __vmaxu2(wI1, __shfl_xor_sync(TDB_FULL, wI1, o));
                    wD1 = __vmaxu2(wD1, __shfl_xor_sync(TDB_FULL, wD1, o));
                    dpk = __vmaxu2(dpk, __shfl_xor_sync(TDB_FULL, dpk, o));
                }
                if (any2) {
#pragma unroll
                    for (int o = 1; o < QG; o <<= 1) {
                        wI2 = __vmaxu2(wI2, __shfl_xor_sync(TDB_FULL, wI2, o));
                        wD2 = __vmaxu2(wD2, __shfl_xor_sync(TDB_FULL, wD2, o));
                    }
                }
                int tl_i1 = lo + (int)(0xffffu - (wI1 >> 16)), th_i1 = lo + (int)(wI1 & 0xffffu) - 1;
                int tl_i2 = lo + (int)(0xffffu - (wI2 >> 16)), th_i2 = lo + (int)(wI2 & 0xffffu) - 1;
                int tl_d1 = lo + (int)(0xffffu - (wD1 >> 16)), th_d1 = lo + (int)(wD1 & 0xffffu) - 1;
                int tl_d2 = lo + (int)(0xffffu - (wD2 >> 16)), th_d2 = lo + (int)(wD2 & 0xffffu) - 1;
                if (live && gl == 0) step_lo[s] = (int16_t)lo, step_base[s] = (StepBase)prov_used;
                prov_used += (unsigned)width;
                __syncwarp();
                const bool done_now = qballot<QG>(gshift, fin) != 0; // group-uniform
                if (live && done_now) penalty = s, running = false;
                const bool cont = live && !done_now;
                // ---- wf-adaptive cut-off + equate (A.3) ----
                int mlo = tl_m, mhi = th_m; // empty when mlo > mhi
                const bool heur_on = cont && p.heur.kind == 1 && mlo <= mhi;
                if (heur_on) --steps_wait;
                const bool cut = heur_on && steps_wait <= 0 && (mhi - mlo + 1) >= p.heur.min_len;
                // The scans drop end diagonals whose distance to the end exceeds the smallest one by more than the threshold
                // and stop at the first that does not; both ends of the trimmed range hold valid offsets, so when no valid
                // cell is that far behind nothing is cut and the scans are skipped (almost always).
                const int min_d = min(max(n, m), (int)(0xffffu - (dpk & 0xffffu))), thr = p.heur.max_dist;
                const bool scan_needed = cut && (int)(dpk >> 16) - min_d > thr;
                if (__any_sync(TDB_FULL, scan_needed)) {
                    const int abase = (int)mpos - lo;
#define TDB_DIST_OK(kk, okv)                                                                  \
    {                                                                                         \
        const int off_ = (int)sm.marena[abase + (kk)];                                        \
        const int d_ = off_ >= 0 ? max(n - (off_ - (kk)), m - off_) : (1 << 30);              \
        okv = d_ - min_d <= thr;                                                              \
    }
                    const int top_limit = min(k_end, mhi);
                    int new_lo = mlo;
                    {
                        int first_ok = INT32_MAX;
                        bool scan = scan_needed && top_limit > mlo;
                        for (int kb = mlo; __any_sync(TDB_FULL, scan); kb += QG) {
                            const int k = kb + gl;
                            bool ok = false;
                            if (scan && k < top_limit) TDB_DIST_OK(k, ok)
                            const unsigned bal = qballot<QG>(gshift, ok);
                            if (scan && bal) first_ok = kb + __ffs(bal) - 1;
                            scan = scan && first_ok == INT32_MAX && kb + QG < top_limit;
                        }
                        if (scan_needed && top_limit > mlo) new_lo = min(first_ok, top_limit);
                    }
                    const int bot_limit = max(k_end, new_lo);
                    int new_hi = mhi;
                    {
                        int last_ok = INT32_MIN;
                        bool scan = scan_needed && bot_limit < mhi;
                        for (int kt = mhi; __any_sync(TDB_FULL, scan); kt -= QG) {
                            const int k = kt - gl;
                            bool ok = false;
                            if (scan && k > bot_limit) TDB_DIST_OK(k, ok)
                            const unsigned bal = qballot<QG>(gshift, ok);
                            if (scan && bal) last_ok = kt - (__ffs(bal) - 1);
                            scan = scan && last_ok == INT32_MIN && kt - QG > bot_limit;
                        }
                        if (scan_needed && bot_limit < mhi) new_hi = max(last_ok, bot_limit);
                    }
#undef TDB_DIST_OK
                    if (scan_needed) mlo = new_lo, mhi = new_hi;
                }
                if (cut) steps_wait = p.heur.steps;
                if (heur_on && mlo <= mhi) { // equate
                    tl_i1 = max(tl_i1, mlo), th_i1 = min(th_i1, mhi);
                    tl_i2 = max(tl_i2, mlo), th_i2 = min(th_i2, mhi);
                    tl_d1 = max(tl_d1, mlo), th_d1 = min(th_d1, mhi);
                    tl_d2 = max(tl_d2, mlo), th_d2 = min(th_d2, mhi);
                }
                // ---- publish metadata of score s ----
                rI1b = rI1a, rD1b = rD1a, loB = loA, loA = lo;
                if (cont) {
                    if (mlo <= mhi) Mm |= 1u;
                    if (tl_i1 <= th_i1) I1m |= 1u;
                    if (tl_i2 <= th_i2) I2m |= 1u;
                    if (tl_d1 <= th_d1) D1m |= 1u;
                    if (tl_d2 <= th_d2) D2m |= 1u;
                    rI1a = qr_pack(tl_i1, th_i1), rD1a = qr_pack(tl_d1, th_d1);
                    rI2 = qr_pack(tl_i2, th_i2), rD2 = qr_pack(tl_d2, th_d2);
                    if (mlo <= mhi) {
                        if (gl == 0) {
                            const unsigned at = mpos + (unsigned)(mlo - lo);
                            if (WIDE) reinterpret_cast<uint2 *>(sm.mmeta)[s & 31] = make_uint2(qr_pack(mlo, mhi), at);
                            else sm.mmeta[s & 31] = qm_pack(mlo, mhi, at);
                        }
                        mhead = mpos + (unsigned)width; // null M keeps no arena
                    }
                }
                __syncwarp();
            }
            // ---------------- backtrace + forward replay (groups that found an alignment) ----------------
            bool aligned = usable && !identical && !overflow;
            int nops = 0, n_del = 0;
            {
                // Inside a gap the chain is predictable while it extends (score - e, diagonal -+ 1 per column): lane j of the
                // group reads the byte the chain reaches after j further extensions, one vote finds the first cell that opened
                // the gap instead, and up to G columns are consumed per round of dependent loads.  Lanes past that cell read
                // bytes that are not on the path (the index is clamped into the provenance area); they are not looked at.
                constexpr unsigned GM = (QG == 32) ? 0xffffffffu : ((1u << QG) - 1u);
                int comp = -1, cs = penalty, k = k_end;
                bool walking = aligned && !(comp < 0 && cs == 0);
                while (__any_sync(TDB_FULL, walking)) {
                    const bool gap = walking && comp >= 0;
                    const bool is_ins = comp == CI1 || comp == CI2, first = comp == CI1 || comp == CD1;
                    const int e = first ? E1 : E2, o = first ? O1 : O2, dk = is_ins ? -1 : 1;
                    const int bit = comp == CI1 ? PB_I1E : (comp == CI2 ? PB_I2E : (comp == CD1 ? PB_D1E : PB_D2E));
                    const int j = gap ? gl : 0;
                    const int cj = cs - j * e, kj = k + j * dk;
                    int b = 0;
                    if (walking && cj > 0) {
                        const unsigned at = (unsigned)step_base[cj] + (unsigned)(kj - (int)step_lo[cj]);
                        b = prov[min(at, prov_used - 1u)];
                    }
                    const unsigned opened = ~qballot<QG>(gshift, gap && (b & bit) != 0) & GM;
                    if (walking) {
                        if (WIDE && nops + QG > C::OPS) { // no room for the op stream: the next tier takes the pair
                            overflow = true, aligned = false, walking = false;
                        } else if (!gap) {
                            const int srcm = b & 7;
                            int op;
                            if (srcm == PM_X) op = OP_X, cs -= X;
                            else op = OP_CLOSE, comp = srcm - 1;
                            if (gl == 0) ops[nops] = (uint8_t)op;
                            ++nops;
                        } else {
                            const int cnt = opened ? __ffs(opened) : QG; // cnt - 1 extensions and, if any, the opening column
                            if (gl < cnt) ops[nops + gl] = (uint8_t)(is_ins ? OP_I : OP_D);
                            nops += cnt;
                            if (!is_ins) n_del += cnt;
                            k += cnt * dk;
                            if (opened) cs -= (cnt - 1) * e + o + e, comp = -1;
                            else cs -= QG * e;
                        }
                        walking = walking && !(comp < 0 && cs == 0);
                    }
                }
            }
            __syncwarp();
            {
                constexpr unsigned GM = (QG == 32) ? 0xffffffffu : ((1u << QG) - 1u);
                ClipAcc acc;
                acc.init(m + n_del, p.clip);
                int v = 0, h = 0, i = nops - 1;
                bool in_m = true, replay = aligned;
                while (__any_sync(TDB_FULL, replay)) {
                    const int eq = extend_groups<QG>(pw, tw, psh, tsh, n, m, v, h, replay && in_m, gl, gshift, lane);
                    // a run of equal gap columns is found by one vote over the next G ops and emitted at once
                    int op = OP_CLOSE;
                    bool same = false;
                    if (replay && i >= 0) {
                        op = ops[i];
                        same = i - gl >= 0 && ops[i - gl] == op;
                    }
                    const unsigned differs = ~qballot<QG>(gshift, same) & GM;
                    if (replay) {
                        if (in_m) {
                            acc.emit('M', eq, pn);
                            v += eq, h += eq;
                        }
                        if (i < 0) {
                            replay = false;
                        } else if (op == OP_CLOSE) {
                            in_m = true;
                            --i;
                        } else if (op == OP_X) {
                            acc.emit('X', 1, pn);
                            ++v, ++h, --i;
                        } else {
                            const int run = differs ? __ffs(differs) - 1 : QG; // >= 1: lane 0 holds ops[i] itself
                            acc.emit(op == OP_I ? 'I' : 'D', run, pn);
                            if (op == OP_I) h += run;
                            else v += run;
                            in_m = false;
                            i -= run;
                        }
                    }
                }
                if (aligned) {
                    clipped = acc.finish(pn);
                    cols = (unsigned)acc.col;
                }
            }
            __syncwarp();
            // ---- result ----
            unsigned ext_pair = 0;
            if (p.work) { // per-pair algorithmic work: ext is a per-lane partial count (warp-uniform branch)
                ext_pair = (unsigned)ext;
#pragma unroll
                for (int o = 1; o < QG; o <<= 1) ext_pair += __shfl_xor_sync(TDB_FULL, ext_pair, o);
            }
            if (have) {
                if (gl == 0) {
                    if (!overflow) {
                        p.out_score[idx] = clipped;
                        if (p.out_status) p.out_status[idx] = TDB_STATUS_ALG_COMPLETED;
                        if (p.out_penalty) p.out_penalty[idx] = penalty;
                        if (p.work) {
                            p.work[3 * (size_t)idx] = (uint32_t)cells, p.work[3 * (size_t)idx + 1] = ext_pair;
                            p.work[3 * (size_t)idx + 2] = (uint32_t)cols;
                        }
                        acc_cells += cells, acc_cols += cols, acc_done += 1;
                    } else if (p.next_todo) {
                        const unsigned j = atomicAdd(p.next_count, 1u);
                        p.next_todo[j] = idx;
                    } else {
                        p.out_score[idx] = TDB_SCORE_FAILED;
                        if (p.out_status) p.out_status[idx] = TDB_STATUS_OOM;
                        if (p.out_penalty) p.out_penalty[idx] = INT32_MIN;
                        if (p.fail_count) atomicAdd(p.fail_count, 1u);
                    }
                }
                if (!overflow) acc_ext += ext;
            }
            __syncwarp();
        }
    }
    // per-warp totals: ext is per lane, the others live on each group's lane 0
    unsigned long long c = (gl == 0) ? acc_cells : 0ull, t = (gl == 0) ? acc_cols : 0ull, d = (gl == 0) ? acc_done : 0ull;
    for (int o = 16; o; o >>= 1) {
        acc_ext += __shfl_xor_sync(TDB_FULL, acc_ext, o);
        c += __shfl_xor_sync(TDB_FULL, c, o);
        t += __shfl_xor_sync(TDB_FULL, t, o);
        d += __shfl_xor_sync(TDB_FULL, d, o);
    }
    if (lane == 0 && p.counters) {
        atomicAdd(p.counters + 0, c);
        atomicAdd(p.counters + 1, acc_ext);
        atomicAdd(p.counters + 2, t);
        atomicAdd(p.counters + 3, d);
    }
}

} // namespace tdb
