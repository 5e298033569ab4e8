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
//     with width <= 32 fit in 3 KB and keeps 16 warps resident per SM;
//   * penalties are compile-time constants (the reference hard-wires 8,4,2,24,1: src/commands/shared.rs:39-51,
//     src/model.rs:13-23); other penalties use the general tiers (wfa_align_warp, tdb_wfa.cuh).
// Semantics are exactly those of wfa_align_warp: SURVEY.md Appendix A.1-A.4.
#pragma once
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
// Geometry of a lockstep configuration: G lanes per pair (32/G pairs per warp), wavefronts up to W
// diagonals, M arena of MAR entries (power of two), scores < SCAP, PROV provenance bytes.
template <int G_, int W_, int MAR_, int SCAP_, int PROV_>
struct LockCfg {
    static constexpr int G = G_, W = W_, MAR = MAR_, SCAP = SCAP_, PROV = PROV_;
};
using QuadCfg = LockCfg<8, 32, TDB_Q_MAR, TDB_Q_SCAP, TDB_Q_PROV>; // tier 0: four pairs per warp
#ifndef TDB_DUO_G
#define TDB_DUO_G 16
#endif
using DuoCfg = LockCfg<TDB_DUO_G, 64, 1024, 128, 2048>;             // tier 1: two pairs per warp
constexpr int Q_MAXLEN = 8000;
constexpr int Q_WARPS = TDB_Q_WARPS;
constexpr int Q_NULL = -16384;

struct __align__(8) QMMeta {
    int16_t lo, hi;  // valid range
    uint16_t off;    // arena index of diagonal lo (free-running, masked on access)
    uint16_t pad;
};
struct __align__(4) QGMeta {
    int16_t lo, hi;
};
template <class C>
struct __align__(16) QPairT {
    int16_t marena[C::MAR];       // FIFO arena of M wavefronts; reused as the op stream at backtrace
    int16_t gap[12][C::W];        // I1[4], D1[4], I2[2], D2[2] rings indexed by score
    QMMeta mmeta[32];             // valid range + arena position of M[s], slot s & 31 (one 8-byte load)
    QGMeta gmeta[12];             // valid range of the gap wavefronts (one 4-byte load)
    int16_t step_lo[C::SCAP];     // computed lo of score s (provenance row origin)
    uint16_t step_base[C::SCAP];
    uint8_t prov[C::PROV];
};
static_assert(sizeof(QPairT<QuadCfg>) % 16 == 0 && sizeof(QPairT<DuoCfg>) % 16 == 0, "pair state keeps 16-byte alignment");
template <class C>
constexpr size_t lock_smem_bytes() { return sizeof(QPairT<C>) * (32 / C::G) * Q_WARPS; }

// Bit mask over the (at most W) diagonals of one wavefront; first / last set bit = trimmed range.
template <int W>
struct QMask {
    static_assert(W == 32 || W == 64, "wavefront width");
    unsigned long long bits = 0; // W == 32 only uses the low word (the compiler drops the rest)
    __device__ __forceinline__ void add(unsigned group_bits, int shift) {
        if (W == 32) bits = (unsigned)bits | (group_bits << shift);
        else bits |= (unsigned long long)group_bits << shift;
    }
    __device__ __forceinline__ void range(int lo, int &first, int &last) const {
        if (W == 32) {
            const unsigned b = (unsigned)bits;
            first = b ? lo + __ffs(b) - 1 : INT32_MAX, last = b ? lo + 31 - __clz(b) : INT32_MIN;
        } else {
            first = bits ? lo + __ffsll((long long)bits) - 1 : INT32_MAX, last = bits ? lo + 63 - __clzll((long long)bits) : INT32_MIN;
        }
    }
};

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
__global__ void __launch_bounds__(Q_WARPS * 32, TDB_Q_MINB) k_align_lock(const AlignParams p) {
    static_assert(X < 32 && O1 + E1 < 32 && O2 + E2 < 32 && E1 < 4 && E2 < 2, "ring sizes");
    constexpr int QG = C::G, QW = C::W, Q_MAR = C::MAR, Q_SCAP = C::SCAP, Q_PROV = C::PROV;
    using QPair = QPairT<C>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gl = lane & (QG - 1), gid = lane / QG, gshift = gid * QG;
    QPair &sm = reinterpret_cast<QPair *>(smem_raw)[wid * (32 / QG) + gid];
    const Penalties pn = {X, O1, E1, O2, E2};
    const ItemSrc src = item_src(p);
    const unsigned n_items = src.total;
    // pairs per work-counter atomic: 8 batches when there is plenty of work, fewer when a launch (a pipeline
    // chunk) only has a few batches per warp -- coarse grabs then leave warps idle behind the heaviest ones
    constexpr int PER_BATCH = 32 / QG, MAX_FETCH = 8 * PER_BATCH;
    const unsigned n_warps = gridDim.x * (blockDim.x >> 5);
    const unsigned per_fetch =
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
                overflow = (p.seq_flag[ip] | p.seq_flag[jt]) || n > Q_MAXLEN || m > Q_MAXLEN;
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
                sm.mmeta[0] = QMMeta{0, 0, 0, 0};
                sm.marena[0] = (int16_t)m0;
                sm.step_lo[0] = 0, sm.step_base[0] = 0;
                sm.prov[0] = 0;
            }
            unsigned Mm = 1u, I1m = 0, I2m = 0, D1m = 0, D2m = 0; // bit j <-> score s-j exists
            unsigned mhead = 1, prov_used = 1;
            int steps_wait = p.heur.steps - (p.heur.kind == 1 ? 1 : 0);
            int s = 0;
            __syncwarp();
            while (__any_sync(TDB_FULL, running)) {
                ++s;
                if (s >= Q_SCAP) { // warp-uniform
                    if (running) overflow = true, running = false;
                    break;
                }
                Mm <<= 1, I1m <<= 1, I2m <<= 1, D1m <<= 1, D2m <<= 1;
                bool live = running && ((Mm & ((1u << X) | (1u << (O1 + E1)) | (1u << (O2 + E2)))) |
                                        ((I1m | D1m) & (1u << E1)) | ((I2m | D2m) & (1u << E2))) != 0;
                if (!__any_sync(TDB_FULL, live)) continue; // null step for every group: registers only
                // ---- inputs ----
                QIn mx, mo1, mo2, i1e, i2e, d1e, d2e;
#define TDB_QM(w, lag)                                                              \
    w.p = sm.marena, w.base = 0, w.lo = 1, w.span1 = 0;                             \
    if (live && (Mm & (1u << (lag)))) {                                             \
        const QMMeta mm_ = sm.mmeta[(s - (lag)) & 31];                              \
        w.lo = mm_.lo, w.span1 = (unsigned)(mm_.hi - mm_.lo + 1), w.base = (int)mm_.off - mm_.lo; \
    }
                TDB_QM(mx, X)
                TDB_QM(mo1, O1 + E1)
                TDB_QM(mo2, O2 + E2)
#undef TDB_QM
#define TDB_QG(w, msk, lag, row0, rmask)                                            \
    w.p = sm.gap[0], w.base = 0, w.lo = 1, w.span1 = 0;                             \
    if (live && (msk & (1u << (lag)))) {                                            \
        const int r_ = (row0) + ((s - (lag)) & (rmask));                            \
        const QGMeta gm_ = sm.gmeta[r_];                                            \
        w.p = sm.gap[r_], w.lo = gm_.lo, w.span1 = (unsigned)(gm_.hi - gm_.lo + 1); \
    }
                TDB_QG(i1e, I1m, E1, 0, 3)
                TDB_QG(d1e, D1m, E1, 4, 3)
                TDB_QG(i2e, I2m, E2, 8, 1)
                TDB_QG(d2e, D2m, E2, 10, 1)
#undef TDB_QG
                int lo = INT32_MAX, hi = INT32_MIN;
#define TDB_LIM(w, dl, dh)                                              \
    if ((w).span1) {                                                    \
        lo = min(lo, (w).lo + (dl));                                    \
        hi = max(hi, (w).lo + (int)(w).span1 - 1 + (dh));               \
    }
                TDB_LIM(mx, 0, 0)
                TDB_LIM(mo1, -1, 1)
                TDB_LIM(mo2, -1, 1)
                TDB_LIM(i1e, 1, 1)
                TDB_LIM(i2e, 1, 1)
                TDB_LIM(d1e, -1, -1)
                TDB_LIM(d2e, -1, -1)
#undef TDB_LIM
                int width = 0;
                if (live) {
                    lo = max(lo, -n - 1), hi = min(hi, m + 1);
                    width = hi - lo + 1;
                    // arena room: everything from the oldest live M wavefront up to the new one
                    const unsigned livem = Mm & 0x03fffffeu; // lags 1..25
                    unsigned tail = mhead;
                    if (livem) tail = sm.mmeta[(s - (31 - __clz(livem))) & 31].off;
                    if (width > QW || prov_used + (unsigned)width > Q_PROV ||
                        ((mhead - tail) & 0xffffu) + (unsigned)width > Q_MAR) {
                        overflow = true, running = false, live = false, width = 0; // this group drops out
                    }
                }
                if (!live) lo = 0, hi = -1;
                int16_t *I1o = sm.gap[0 + (s & 3)], *D1o = sm.gap[4 + (s & 3)];
                int16_t *I2o = sm.gap[8 + (s & 1)], *D2o = sm.gap[10 + (s & 1)];
                uint8_t *pv = sm.prov + prov_used;
                cells += (unsigned)width;
                // validity masks of this score's wavefronts, bit = diagonal - lo (a wavefront spans at most QW diagonals)
                QMask<QW> vmask_m, vmask_i1, vmask_i2, vmask_d1, vmask_d2;
                int lane_min_d = 1 << 30;
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
#define TDB_RDM(w, kk) (((unsigned)((kk) - (w).lo) < (w).span1) ? (int)sm.marena[((w).base + (kk)) & (Q_MAR - 1)] : Q_NULL)
#define TDB_RDG(w, kk) (((unsigned)((kk) - (w).lo) < (w).span1) ? (int)(w).p[(kk) & (QW - 1)] : Q_NULL)
                        int b = 0, o, e;
                        o = TDB_RDM(mo1, k - 1), e = TDB_RDG(i1e, k - 1);
                        if (e >= o) ins1 = e, b |= PB_I1E; else ins1 = o;
                        ins1 += 1;
                        o = TDB_RDM(mo2, k - 1), e = TDB_RDG(i2e, k - 1);
                        if (e >= o) ins2 = e, b |= PB_I2E; else ins2 = o;
                        ins2 += 1;
                        o = TDB_RDM(mo1, k + 1), e = TDB_RDG(d1e, k + 1);
                        if (e >= o) del1 = e, b |= PB_D1E; else del1 = o;
                        o = TDB_RDM(mo2, k + 1), e = TDB_RDG(d2e, k + 1);
                        if (e >= o) del2 = e, b |= PB_D2E; else del2 = o;
                        const int mis = TDB_RDM(mx, k) + 1;
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
                    int eq = 0;
                    bool ebusy = vm && eroom > 0;
                    while (__any_sync(TDB_FULL, ebusy)) {
                        if (ebusy) {
                            const uint32_t x = get16(pw, psh + ev + eq) ^ get16(tw, tsh + mv + eq);
                            if (x) eq += (__ffs(x) - 1) >> 1, ebusy = false;
                            else eq += 16, ebusy = eq < eroom;
                        }
                    }
                    if (act) {
                        if (vm) {
                            eq = min(eq, eroom);
                            ext += (unsigned)(eq >> 4) + 1;
                            mv += eq;
                            lane_min_d = min(lane_min_d, max(n - (mv - k), m - mv));
                            fin |= (k == k_end) && (mv >= m);
                        }
                        sm.marena[(mhead + (unsigned)(k - lo)) & (Q_MAR - 1)] = (int16_t)mv;
                        // a component without a source only holds NULL+1 here: never valid, so its range stays empty
                        I1o[k & (QW - 1)] = (int16_t)ins1;
                        I2o[k & (QW - 1)] = (int16_t)ins2;
                        D1o[k & (QW - 1)] = (int16_t)del1;
                        D2o[k & (QW - 1)] = (int16_t)del2;
                    }
                    // which cells lie inside the DP matrix: collected per score, trimmed once below (A.2)
                    const int sh = itk * QG;
                    vmask_m.add(qballot<QG>(gshift, vm), sh);
#define TDB_VALID(val) (act && (unsigned)(val) <= (unsigned)m && (unsigned)((val)-k) <= (unsigned)n)
                    vmask_i1.add(qballot<QG>(gshift, TDB_VALID(ins1)), sh);
                    vmask_i2.add(qballot<QG>(gshift, TDB_VALID(ins2)), sh);
                    vmask_d1.add(qballot<QG>(gshift, TDB_VALID(del1)), sh);
                    vmask_d2.add(qballot<QG>(gshift, TDB_VALID(del2)), sh);
#undef TDB_VALID
                }
                // trim_ends (A.2): first / last diagonal whose offset lies inside the DP matrix
                int tl_m, th_m, tl_i1, th_i1, tl_i2, th_i2, tl_d1, th_d1, tl_d2, th_d2;
                vmask_m.range(lo, tl_m, th_m);
                vmask_i1.range(lo, tl_i1, th_i1);
                vmask_i2.range(lo, tl_i2, th_i2);
                vmask_d1.range(lo, tl_d1, th_d1);
                vmask_d2.range(lo, tl_d2, th_d2);
                if (live && gl == 0) sm.step_lo[s] = (int16_t)lo, sm.step_base[s] = (uint16_t)prov_used;
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
                if (__any_sync(TDB_FULL, cut)) {
                    int md = lane_min_d;
#pragma unroll
                    for (int o = 1; o < QG; o <<= 1) md = min(md, __shfl_xor_sync(TDB_FULL, md, o));
                    const int min_d = min(max(n, m), md), thr = p.heur.max_dist;
                    const int abase = (int)mhead - lo;
#define TDB_DIST_OK(kk, okv)                                                                  \
    {                                                                                         \
        const int off_ = (int)sm.marena[(abase + (kk)) & (Q_MAR - 1)];                        \
        const int d_ = off_ >= 0 ? max(n - (off_ - (kk)), m - off_) : (1 << 30);              \
        okv = d_ - min_d <= thr;                                                              \
    }
                    const int top_limit = min(k_end, mhi);
                    int new_lo = mlo;
                    {
                        int first_ok = INT32_MAX;
                        bool scan = cut && top_limit > mlo;
                        for (int kb = mlo; __any_sync(TDB_FULL, scan); kb += QG) {
                            const int k = kb + gl;
                            bool ok = false;
                            if (scan && k < top_limit) TDB_DIST_OK(k, ok)
                            const unsigned bal = qballot<QG>(gshift, ok);
                            if (scan && bal) first_ok = kb + __ffs(bal) - 1;
                            scan = scan && first_ok == INT32_MAX && kb + QG < top_limit;
                        }
                        if (cut && top_limit > mlo) new_lo = min(first_ok, top_limit);
                    }
                    const int bot_limit = max(k_end, new_lo);
                    int new_hi = mhi;
                    {
                        int last_ok = INT32_MIN;
                        bool scan = cut && bot_limit < mhi;
                        for (int kt = mhi; __any_sync(TDB_FULL, scan); kt -= QG) {
                            const int k = kt - gl;
                            bool ok = false;
                            if (scan && k > bot_limit) TDB_DIST_OK(k, ok)
                            const unsigned bal = qballot<QG>(gshift, ok);
                            if (scan && bal) last_ok = kt - (__ffs(bal) - 1);
                            scan = scan && last_ok == INT32_MIN && kt - QG > bot_limit;
                        }
                        if (cut && bot_limit < mhi) new_hi = max(last_ok, bot_limit);
                    }
#undef TDB_DIST_OK
                    if (cut) mlo = new_lo, mhi = new_hi, steps_wait = p.heur.steps;
                }
                if (heur_on && mlo <= mhi) { // equate
                    tl_i1 = max(tl_i1, mlo), th_i1 = min(th_i1, mhi);
                    tl_i2 = max(tl_i2, mlo), th_i2 = min(th_i2, mhi);
                    tl_d1 = max(tl_d1, mlo), th_d1 = min(th_d1, mhi);
                    tl_d2 = max(tl_d2, mlo), th_d2 = min(th_d2, mhi);
                }
                // ---- publish metadata of score s ----
                if (cont) {
                    if (mlo <= mhi) Mm |= 1u;
                    if (tl_i1 <= th_i1) I1m |= 1u;
                    if (tl_i2 <= th_i2) I2m |= 1u;
                    if (tl_d1 <= th_d1) D1m |= 1u;
                    if (tl_d2 <= th_d2) D2m |= 1u;
                    if (gl == 0) {
                        sm.mmeta[s & 31] = QMMeta{(int16_t)mlo, (int16_t)mhi, (uint16_t)(mhead + (unsigned)(mlo - lo)), 0};
                        sm.gmeta[0 + (s & 3)] = QGMeta{(int16_t)tl_i1, (int16_t)th_i1};
                        sm.gmeta[4 + (s & 3)] = QGMeta{(int16_t)tl_d1, (int16_t)th_d1};
                        sm.gmeta[8 + (s & 1)] = QGMeta{(int16_t)tl_i2, (int16_t)th_i2};
                        sm.gmeta[10 + (s & 1)] = QGMeta{(int16_t)tl_d2, (int16_t)th_d2};
                    }
                    if (mlo <= mhi) mhead = (mhead + (unsigned)width) & 0xffffu; // null M keeps no arena
                }
                __syncwarp();
            }
            // ---------------- backtrace + forward replay (groups that found an alignment) ----------------
            const bool aligned = usable && !identical && !overflow;
            uint8_t *ops = reinterpret_cast<uint8_t *>(sm.marena); // arena is dead now
            int nops = 0, n_del = 0;
            {
                int comp = -1, cs = penalty, k = k_end;
                bool walking = aligned && !(comp < 0 && cs == 0);
                while (__any_sync(TDB_FULL, walking)) {
                    if (walking) {
                        const int b = sm.prov[(unsigned)sm.step_base[cs] + (unsigned)(k - (int)sm.step_lo[cs])];
                        int op;
                        if (comp < 0) {
                            const int srcm = b & 7;
                            if (srcm == PM_X) op = OP_X, cs -= X;
                            else op = OP_CLOSE, comp = srcm - 1;
                        } else if (comp == CI1 || comp == CI2) {
                            op = OP_I;
                            const bool ex = b & (comp == CI1 ? PB_I1E : PB_I2E);
                            const int e = comp == CI1 ? E1 : E2, o = comp == CI1 ? O1 : O2;
                            if (ex) cs -= e; else cs -= o + e, comp = -1;
                            k -= 1;
                        } else {
                            op = OP_D;
                            ++n_del;
                            const bool ex = b & (comp == CD1 ? PB_D1E : PB_D2E);
                            const int e = comp == CD1 ? E1 : E2, o = comp == CD1 ? O1 : O2;
                            if (ex) cs -= e; else cs -= o + e, comp = -1;
                            k += 1;
                        }
                        if (gl == 0) ops[nops] = (uint8_t)op;
                        ++nops;
                        walking = !(comp < 0 && cs == 0);
                    }
                }
            }
            __syncwarp();
            {
                ClipAcc acc;
                acc.init(m + n_del, p.clip);
                int v = 0, h = 0, i = nops - 1;
                bool in_m = true, replay = aligned;
                while (__any_sync(TDB_FULL, replay)) {
                    const int eq = extend_groups<QG>(pw, tw, psh, tsh, n, m, v, h, replay && in_m, gl, gshift, lane);
                    if (replay) {
                        if (in_m) {
                            acc.emit('M', eq, pn);
                            v += eq, h += eq;
                        }
                        if (i < 0) {
                            replay = false;
                        } else {
                            const int op = ops[i];
                            --i;
                            if (op == OP_CLOSE) {
                                in_m = true;
                            } else {
                                acc.emit(op == OP_X ? 'X' : (op == OP_I ? 'I' : 'D'), 1, pn);
                                if (op == OP_X) ++v, ++h, in_m = true;
                                else if (op == OP_I) ++h, in_m = false;
                                else ++v, in_m = false;
                            }
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
