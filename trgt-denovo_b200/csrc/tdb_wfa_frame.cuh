// tdb_wfa_frame.cuh -- k_align_frame: the lockstep tiers on NULL-filled fixed frames (sm_100a).
//
// Same mapping as k_align_lock (tdb_wfa_quad.cuh): 32/G pairs per warp, G lanes per pair, one lane per diagonal,
// one warp-uniform control flow.  What changed is where the per-score bookkeeping went.  An ncu source-level
// count of k_align_lock<DuoCfg> (profiles/r02/NOTES.md) showed ~16 000 warp instructions per pair, spread evenly
// over range-checked reads (18 %), per-component range tracking (finalize + butterfly, 16 %), input metadata and
// limits (14 %), with the recurrence itself at 6 %.  Here:
//   * every wavefront of the window is a full ROW of W offsets indexed by k mod W, NULL outside its trimmed range
//     (the row is NULL-filled with one 8-byte store per lane before its cells are written; an absent input
//     is a pointer to a constant all-NULL row), so the nine reads of a cell are UNCONDITIONAL loads;
//   * the computed range [lo, hi] of a score only has to be a superset of WFA2's (a cell whose sources are all
//     NULL computes NULL: A.2), so the gap components carry ONE conservative range per score in a register
//     (the computed range, clamped to M's by equate) instead of four exact ones; the algorithmic cell count C of
//     SURVEY 8(d) stays exact: WFA2's range is the hull of the cells that have a source inside an input's
//     trimmed range, and the ends of a trimmed range hold valid (>= 0) offsets, so one vote per iteration on
//     "some source >= 0" gives it;
//   * the trimmed ranges of the gap components are only needed when they differ from "NULL where negative":
//     when some computed offset is positive but outside the DP matrix (the wavefront touches the matrix border,
//     the last steps of a pair) or the cut-off dropped diagonals of M.  A warp-uniform slow path then trims
//     (A.2), equates (A.3) and overwrites what fell outside with NULL; otherwise nothing is tracked.
//     (A NULL-ish negative offset stays negative for the whole alignment: NULL = -16384, at most +1 per score,
//     scores < 128.)
//   * M history: 26 rows (max(x, o1+e1, o2+e2) + 1) by score mod 26, their trimmed ranges as packed halfwords
//     (-lo, hi) so that the limits are one VIADDMNMX.S16x2 per input.
// Semantics are exactly those of wfa_align_warp (tdb_wfa.cuh): SURVEY.md Appendix A.1-A.4; reference call sites
// src/aligner.rs:248-260 (align_end_to_end), :359-376 (cigar_score_clipped), config src/commands/shared.rs:39-51.
#pragma once
#include "tdb_wfa_quad.cuh"

namespace tdb {

#ifndef TDB_F_WARPS
#define TDB_F_WARPS 4
#endif
#ifndef TDB_F_MINB
#define TDB_F_MINB 1
#endif
constexpr int F_WARPS = TDB_F_WARPS; // warps per CTA
constexpr int F_MROWS = 26;          // M rows: scores s-25 .. s
template <int G_, int W_, int SCAP_, int PROV_>
struct FrameCfg {
    static constexpr int G = G_, W = W_, SCAP = SCAP_, PROV = PROV_;
};
#ifndef TDB_FQ_PROV
#define TDB_FQ_PROV 736
#endif
#ifndef TDB_FD_PROV
#define TDB_FD_PROV 1792
#endif
using QuadF = FrameCfg<8, 32, 96, TDB_FQ_PROV>;   // tier 0: four pairs per warp, wavefronts up to 31 diagonals
using DuoF = FrameCfg<16, 64, 128, TDB_FD_PROV>;  // tier 1: two pairs per warp, up to 63 diagonals

template <class C>
struct __align__(16) FPairT {
    int16_t mrow[F_MROWS][C::W]; // M[s] at row s mod 26, diagonal k at k & (W-1); NULL outside the trimmed range
    int16_t grow[12][C::W];      // I1[4], D1[4] by s & 3; I2[2], D2[2] by s & 1
    uint32_t mmeta[F_MROWS + 2]; // trimmed range of M[s]: (-lo) | hi << 16 (signed halfwords)
    uint16_t step_base[C::SCAP]; // provenance index of diagonal step_lo[s]
    int8_t step_lo[C::SCAP];     // computed lo of score s
    uint8_t prov[C::PROV];
};
template <class C>
constexpr size_t frame_smem_bytes() { return sizeof(FPairT<C>) * (32 / C::G) * F_WARPS + (size_t)C::W * 2; }

// a range in one register, ready for halfword max: -lo in the low half, hi in the high half
__device__ __forceinline__ uint32_t fr_pack(int lo, int hi) { return ((uint32_t)(-lo) & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ int fr_lo(uint32_t r) { return -(int)(int16_t)(r & 0xffffu); }
__device__ __forceinline__ int fr_hi(uint32_t r) { return (int)r >> 16; }
constexpr uint32_t FR_EMPTY = 0x80008000u; // (-32768, -32768): loses every halfword max
constexpr uint32_t FR_ONE = 0x00010001u;   // widen a range by one diagonal on both sides

template <class C, int X, int O1, int E1, int O2, int E2>
__global__ void __launch_bounds__(F_WARPS * 32, TDB_F_MINB) k_align_frame(const AlignParams p) {
    static_assert(X < F_MROWS && O1 + E1 < F_MROWS && O2 + E2 < F_MROWS && E1 == 2 && E2 == 1,
                  "ring sizes; gap ranges are kept for two / one score(s)");
    static_assert((C::W * 2) % (8 * C::G) == 0 && (C::W * 2) / C::G == 8, "a row is one 8-byte store per lane");
    constexpr int G = C::G, W = C::W, SCAP = C::SCAP, PROV = C::PROV, PPW = 32 / G;
    using FPair = FPairT<C>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gl = lane & (G - 1), gid = lane / G, gshift = gid * G;
    FPair &sm = reinterpret_cast<FPair *>(smem_raw)[wid * PPW + gid];
    int16_t *const nullrow = reinterpret_cast<int16_t *>(smem_raw + sizeof(FPair) * PPW * F_WARPS);
    for (int i = threadIdx.x; i < W; i += blockDim.x) nullrow[i] = (int16_t)Q_NULL;
    __syncthreads();
    const Penalties pn = {X, O1, E1, O2, E2};
    const ItemSrc src = item_src(p);
    const unsigned n_items = src.total;
    constexpr int MAX_FETCH = 8 * PPW;
    const unsigned n_warps = gridDim.x * (blockDim.x >> 5);
    const unsigned per_fetch =
        (unsigned)PPW * max(1u, min((unsigned)(MAX_FETCH / PPW), n_items / (n_warps * 8u * (unsigned)PPW)));
    unsigned long long acc_cells = 0, acc_cols = 0, acc_ext = 0, acc_done = 0; // acc_* valid on group lane 0
    const uint2 null2 = make_uint2(0xc000c000u, 0xc000c000u);                  // four NULL offsets
    static_assert((uint16_t)Q_NULL == 0xc000u, "null2");

    for (;;) {
        unsigned fb = 0;
        if (lane == 0) fb = atomicAdd(p.work_counter, per_fetch);
        fb = __shfl_sync(TDB_FULL, fb, 0);
        if (fb >= n_items) break;
        for (unsigned bb = fb; bb < fb + per_fetch && bb < n_items; bb += PPW) {
            // ====================== one pair per group, the groups of the warp in lockstep ======================
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
            const int m0 = extend_groups<G>(pw, tw, psh, tsh, n, m, 0, 0, usable, gl, gshift, lane);
            if (usable && gl == 0) ext += (unsigned)(m0 >> 4) + 1;
            const bool identical = usable && k_end == 0 && m0 >= m;
            if (identical) cols = (unsigned)m; // penalty 0, clipped score 0
            bool running = usable && !identical;
            // ---------------- score 0: M[0] = {0: m0} ----------------
            if (running) *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(sm.mrow[0]) + gl * 8) = null2;
            __syncwarp();
            if (running && gl == 0) {
                sm.mrow[0][0] = (int16_t)m0;
                sm.mmeta[0] = fr_pack(0, 0);
                sm.step_lo[0] = 0, sm.step_base[0] = 0;
                sm.prov[0] = 0;
            }
            unsigned Mm = 1u, G1m = 0, G2m = 0;  // bit j <-> M[s-j] / some I1 or D1 / some I2 or D2 offset of score s-j exists
            uint32_t R1a = FR_EMPTY, R1b = FR_EMPTY, R2 = FR_EMPTY; // conservative ranges: I1/D1 of s-1, s-2; I2/D2 of s-1
            // what the gap components of those scores add to WFA2's computed range where the vote below cannot see it:
            // after the slow path (trim + equate) a range may end on a negative offset; FR_EMPTY otherwise
            uint32_t X1a = FR_EMPTY, X1b = FR_EMPTY, X2 = FR_EMPTY;
            unsigned prov_used = 1;
            int steps_wait = p.heur.steps - (p.heur.kind == 1 ? 1 : 0);
            int s = 0, slot = 0; // slot = s mod F_MROWS
            __syncwarp();
            // ---------------- wavefront loop (score s is warp-uniform) ----------------
            while (__any_sync(TDB_FULL, running)) {
                ++s;
                slot = slot == F_MROWS - 1 ? 0 : slot + 1;
                if (s >= SCAP) { // warp-uniform
                    if (running) overflow = true, running = false;
                    break;
                }
                Mm <<= 1, G1m <<= 1, G2m <<= 1;
                bool live = running &&
                            ((Mm & ((1u << X) | (1u << (O1 + E1)) | (1u << (O2 + E2)))) | (G1m & (1u << E1)) | (G2m & (1u << E2))) != 0;
                if (!__any_sync(TDB_FULL, live)) { // null step for every group: registers only
                    R1b = R1a, X1b = X1a;
                    continue;
                }
                // ---- inputs: rows of the three M and four gap wavefronts (absent: the all-NULL row), limits ----
                const int sx = slot - X + (slot < X ? F_MROWS : 0), so1 = slot - (O1 + E1) + (slot < O1 + E1 ? F_MROWS : 0);
                const int so2 = slot + 1 == F_MROWS ? 0 : slot + 1; // s - 25 = s + 1 (mod 26)
                static_assert(O2 + E2 == F_MROWS - 1, "so2");
                const bool h_mx = (Mm & (1u << X)) != 0, h_mo1 = (Mm & (1u << (O1 + E1))) != 0, h_mo2 = (Mm & (1u << (O2 + E2))) != 0;
                const bool h_g1 = (G1m & (1u << E1)) != 0, h_g2 = (G2m & (1u << E2)) != 0;
                uint32_t er = FR_EMPTY;
                if (h_mx) er = __vmaxs2(er, sm.mmeta[sx]);
                if (h_mo1) er = __viaddmax_s16x2(sm.mmeta[so1], FR_ONE, er);
                if (h_mo2) er = __viaddmax_s16x2(sm.mmeta[so2], FR_ONE, er);
                // WFA2's range, as far as metadata knows it: the M inputs' trimmed ranges (the cut-off can leave NULL cells at
                // their ends) and the gap ranges the slow path made; the vote on "some source >= 0" adds the rest
                uint32_t xr = er;
                if (h_g1) xr = __vmaxs2(xr, X1b);
                if (h_g2) xr = __vmaxs2(xr, X2);
                if (h_g1) er = __viaddmax_s16x2(R1b, FR_ONE, er);
                if (h_g2) er = __viaddmax_s16x2(R2, FR_ONE, er);
                const int16_t *MXr = h_mx ? sm.mrow[sx] : nullrow, *MO1r = h_mo1 ? sm.mrow[so1] : nullrow;
                const int16_t *MO2r = h_mo2 ? sm.mrow[so2] : nullrow;
                const int16_t *I1r = h_g1 ? sm.grow[0 + ((s - E1) & 3)] : nullrow, *D1r = h_g1 ? sm.grow[4 + ((s - E1) & 3)] : nullrow;
                const int16_t *I2r = h_g2 ? sm.grow[8 + ((s - E2) & 1)] : nullrow, *D2r = h_g2 ? sm.grow[10 + ((s - E2) & 1)] : nullrow;
                int lo = 0, hi = -1, width = 0;
                if (live) {
                    lo = fr_lo(er), hi = fr_hi(er);
                    width = hi - lo + 1;
                    // one slot of the row stays NULL: it is both diagonal lo-1 and diagonal hi+1
                    if (width > W - 1 || prov_used + (unsigned)width > PROV) {
                        overflow = true, running = false, live = false, width = 0, lo = 0, hi = -1; // this group drops out
                    }
                }
                int16_t *Mo = sm.mrow[slot];
                int16_t *I1o = sm.grow[0 + (s & 3)], *D1o = sm.grow[4 + (s & 3)];
                int16_t *I2o = sm.grow[8 + (s & 1)], *D2o = sm.grow[10 + (s & 1)];
                uint8_t *pv = sm.prov + prov_used - lo;
                // the second gap piece does not exist below score O2 + E2: skipped (warp-uniform) while no group has a source
                const bool any2 = __any_sync(TDB_FULL, live && (h_mo2 || h_g2));
                if (live) {
                    *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(Mo) + gl * 8) = null2;
                    *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(I1o) + gl * 8) = null2;
                    *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(D1o) + gl * 8) = null2;
                    if (any2) {
                        *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(I2o) + gl * 8) = null2;
                        *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(D2o) + gl * 8) = null2;
                    }
                }
                __syncwarp();
                int tl_m = INT32_MAX, th_m = INT32_MIN; // first / last diagonal of M inside the DP matrix
                int c_lo = INT32_MAX, c_hi = INT32_MIN; // hull of the cells that have a source: WFA2's computed range
                if (live && xr != FR_EMPTY) c_lo = fr_lo(xr), c_hi = fr_hi(xr);
                uint32_t dpk = 0u;                      // (max d, 0xffff - min d) over the valid M cells: distance to the end
                unsigned fl = 0;                        // 1: some I1/D1 >= 0, 2: some I2/D2 >= 0, 4: positive but outside the matrix, 8: done
                const int n_iter = __reduce_max_sync(TDB_FULL, (width + G - 1) / G);
                for (int itk = 0; itk < n_iter; ++itk) {
                    const int kb = lo + itk * G;
                    const int k = kb + gl;
                    const bool act = k <= hi; // false everywhere in groups that are not live (hi = -1 < lo = 0)
                    const int a0 = k & (W - 1), am = (k - 1) & (W - 1), ap = (k + 1) & (W - 1);
                    int b = 0, o, e, ins1, ins2, del1, del2;
                    o = MO1r[am], e = I1r[am];
                    ins1 = max(o, e) + 1;
                    if (e >= o) b |= PB_I1E;
                    o = MO1r[ap], e = D1r[ap];
                    del1 = max(o, e);
                    if (e >= o) b |= PB_D1E;
                    if (any2) {
                        o = MO2r[am], e = I2r[am];
                        ins2 = max(o, e) + 1;
                        if (e >= o) b |= PB_I2E;
                        o = MO2r[ap], e = D2r[ap];
                        del2 = max(o, e);
                        if (e >= o) b |= PB_D2E;
                    } else {
                        b |= PB_I2E | PB_D2E; // what two absent sources give (NULL >= NULL)
                        ins2 = Q_NULL + 1, del2 = Q_NULL;
                    }
                    const int mis = (int)MXr[a0] + 1;
                    const int gi = max(ins1, ins2), gd = max(del1, del2);
                    int mv = max(max(gi, gd), mis);
                    int srcm = PM_NONE; // later tests override: priority X > D2 > D1 > I2 > I1
                    if (mv == ins1) srcm = PM_I1;
                    if (mv == ins2) srcm = PM_I2;
                    if (mv == del1) srcm = PM_D1;
                    if (mv == del2) srcm = PM_D2;
                    if (mv == mis) srcm = PM_X;
                    const bool hs = act && max(max(gi, mis) - 1, gd) >= 0; // some source offset >= 0
                    const bool vm = act && (unsigned)mv <= (unsigned)m && (unsigned)(mv - k) <= (unsigned)n;
                    if (act) {
                        pv[k] = (uint8_t)(b | srcm);
                        if (max(ins1, del1) >= 0) fl |= 1u;
                        if (mv >= 0 && !vm) fl |= 4u; // mv >= every component: covers each positive offset outside the matrix
                    }
                    if (any2 && act && max(ins2, del2) >= 0) fl |= 2u;
                    // extend at once (trimming never changes offsets); one warp-uniform loop for all lanes
                    const int ev = mv - k, eroom = vm ? min(n - ev, m - mv) : 0;
                    int eq = 0;
                    bool ebusy = eroom > 0;
                    while (__any_sync(TDB_FULL, ebusy)) {
                        if (ebusy) {
                            const uint32_t x = get16(pw, psh + ev + eq) ^ get16(tw, tsh + mv + eq);
                            if (x) eq += (__ffs(x) - 1) >> 1, ebusy = false;
                            else eq += 16, ebusy = eq < eroom;
                        }
                    }
                    if (vm) {
                        eq = min(eq, eroom);
                        ext += (unsigned)(eq >> 4) + 1;
                        mv += eq;
                        const unsigned d = (unsigned)max(n - (mv - k), m - mv);
                        dpk = __vmaxu2(dpk, (d << 16) | (0xffffu - d));
                        if (k == k_end && mv >= m) fl |= 8u;
                    }
                    if (act) {
                        Mo[a0] = (int16_t)(vm ? mv : Q_NULL);
                        I1o[a0] = (int16_t)ins1;
                        D1o[a0] = (int16_t)del1;
                        if (any2) I2o[a0] = (int16_t)ins2, D2o[a0] = (int16_t)del2;
                    }
                    const unsigned balm = qballot<G>(gshift, vm), bals = qballot<G>(gshift, hs);
                    if (balm) tl_m = min(tl_m, kb + __ffs(balm) - 1), th_m = kb + 31 - __clz(balm);
                    if (bals) c_lo = min(c_lo, kb + __ffs(bals) - 1), c_hi = max(c_hi, kb + 31 - __clz(bals));
                }
                // WFA2 keeps its wavefronts inside [-n-1, m+1] (diagonals beyond only hold invalid offsets)
                if (c_lo <= c_hi) {
                    const int w_ = min(c_hi, m + 1) - max(c_lo, -n - 1) + 1;
                    if (w_ > 0) cells += (unsigned)w_;
                }
                if (live && gl == 0) sm.step_lo[s] = (int8_t)lo, sm.step_base[s] = (uint16_t)prov_used;
                prov_used += (unsigned)width;
                const unsigned flg = (__reduce_or_sync(TDB_FULL, fl << (4 * gid)) >> (4 * gid)) & 15u; // group-uniform
                const bool done_now = (flg & 8u) != 0;
                if (live && done_now) penalty = s, running = false;
                const bool cont = live && !done_now;
                // ---- wf-adaptive cut-off (A.3) ----
                int mlo = tl_m, mhi = th_m; // empty when mlo > mhi
                const bool heur_on = cont && p.heur.kind == 1 && mlo <= mhi;
                if (heur_on) --steps_wait;
                const bool cut = heur_on && steps_wait <= 0 && (mhi - mlo + 1) >= p.heur.min_len;
                bool cut_changed = false;
                if (__any_sync(TDB_FULL, cut)) {
#pragma unroll
                    for (int o = 1; o < G; o <<= 1) dpk = __vmaxu2(dpk, __shfl_xor_sync(TDB_FULL, dpk, o));
                    // The scans drop end diagonals whose distance to the end exceeds the smallest one by more than the threshold
                    // and stop at the first that does not; both ends of the trimmed range hold valid offsets, so when no valid
                    // cell is that far behind nothing is cut and the scans are skipped (almost always).
                    const int min_d = min(max(n, m), (int)(0xffffu - (dpk & 0xffffu))), thr = p.heur.max_dist;
                    const bool scan_needed = cut && (int)(dpk >> 16) - min_d > thr;
                    if (__any_sync(TDB_FULL, scan_needed)) {
                        __syncwarp(); // the scans read M[s] as the other lanes stored it
#define TDB_DIST_OK(kk, okv)                                                                  \
    {                                                                                         \
        const int off_ = (int)Mo[(kk) & (W - 1)];                                             \
        const int d_ = off_ >= 0 ? max(n - (off_ - (kk)), m - off_) : (1 << 30);              \
        okv = d_ - min_d <= thr;                                                              \
    }
                        const int top_limit = min(k_end, mhi);
                        int new_lo = mlo;
                        {
                            int first_ok = INT32_MAX;
                            bool scan = scan_needed && top_limit > mlo;
                            for (int kb = mlo; __any_sync(TDB_FULL, scan); kb += G) {
                                const int k = kb + gl;
                                bool ok = false;
                                if (scan && k < top_limit) TDB_DIST_OK(k, ok)
                                const unsigned bal = qballot<G>(gshift, ok);
                                if (scan && bal) first_ok = kb + __ffs(bal) - 1;
                                scan = scan && first_ok == INT32_MAX && kb + G < top_limit;
                            }
                            if (scan_needed && top_limit > mlo) new_lo = min(first_ok, top_limit);
                        }
                        const int bot_limit = max(k_end, new_lo);
                        int new_hi = mhi;
                        {
                            int last_ok = INT32_MIN;
                            bool scan = scan_needed && bot_limit < mhi;
                            for (int kt = mhi; __any_sync(TDB_FULL, scan); kt -= G) {
                                const int k = kt - gl;
                                bool ok = false;
                                if (scan && k > bot_limit) TDB_DIST_OK(k, ok)
                                const unsigned bal = qballot<G>(gshift, ok);
                                if (scan && bal) last_ok = kt - (__ffs(bal) - 1);
                                scan = scan && last_ok == INT32_MIN && kt - G > bot_limit;
                            }
                            if (scan_needed && bot_limit < mhi) new_hi = max(last_ok, bot_limit);
                        }
#undef TDB_DIST_OK
                        if (scan_needed) {
                            cut_changed = new_lo != mlo || new_hi != mhi;
                            mlo = new_lo, mhi = new_hi;
                        }
                    }
                }
                if (cut) steps_wait = p.heur.steps;
                // ---- trim_ends (A.2) + equate (A.3) of the gap components, where they are not the identity ----
                // Valid gap offsets lie on diagonals where M is valid too (M >= every component), unless an offset is positive
                // and outside the matrix, or the cut-off narrowed M: only then do the rows need more than "negative = NULL".
                const bool fix = cont && ((flg & 4u) != 0 || cut_changed);
                uint32_t Rcur = fr_pack(lo, hi), Xcur1 = FR_EMPTY, Xcur2 = FR_EMPTY;
                if (heur_on) Rcur = __vmins2(Rcur, fr_pack(mlo, mhi)); // equate: nothing of a gap component lies outside M
                if (__any_sync(TDB_FULL, fix)) {
                    __syncwarp();
                    const bool eqz = heur_on && mlo <= mhi;
                    int u_lo = INT32_MAX, u_hi = INT32_MIN; // hull of what stays
                    uint32_t x1 = FR_EMPTY, x2 = FR_EMPTY;  // the ranges as the next scores' limits see them: I +1, D -1
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        if (c >= 2 && !any2) break;
                        int16_t *row = c == 0 ? I1o : (c == 1 ? D1o : (c == 2 ? I2o : D2o));
                        int tl = INT32_MAX, th = INT32_MIN;
                        for (int itk = 0; itk < n_iter; ++itk) {
                            const int kb = lo + itk * G, k = kb + gl;
                            bool ok = false;
                            if (fix && k <= hi) {
                                const int val = (int)row[k & (W - 1)];
                                ok = (unsigned)val <= (unsigned)m && (unsigned)(val - k) <= (unsigned)n;
                            }
                            const unsigned bal = qballot<G>(gshift, ok);
                            if (bal) tl = min(tl, kb + __ffs(bal) - 1), th = kb + 31 - __clz(bal);
                        }
                        if (eqz) tl = max(tl, mlo), th = min(th, mhi);
                        for (int itk = 0; itk < n_iter; ++itk) {
                            const int k = lo + itk * G + gl;
                            if (fix && k <= hi && (k < tl || k > th)) row[k & (W - 1)] = (int16_t)Q_NULL;
                        }
                        if (tl <= th) {
                            u_lo = min(u_lo, tl), u_hi = max(u_hi, th);
                            const uint32_t sh = (c & 1) ? fr_pack(tl - 1, th - 1) : fr_pack(tl + 1, th + 1);
                            if (c < 2) x1 = __vmaxs2(x1, sh); else x2 = __vmaxs2(x2, sh);
                        }
                    }
                    if (fix) Xcur1 = x1, Xcur2 = x2;
                    if (cut_changed)
                        for (int itk = 0; itk < n_iter; ++itk) {
                            const int k = lo + itk * G + gl;
                            if (k <= hi && (k < mlo || k > mhi)) Mo[k & (W - 1)] = (int16_t)Q_NULL;
                        }
                    if (fix) Rcur = u_lo <= u_hi ? fr_pack(u_lo, u_hi) : FR_EMPTY;
                }
                // ---- publish metadata of score s ----
                R1b = R1a, X1b = X1a;
                if (cont) {
                    X1a = Xcur1, X2 = Xcur2;
                    if (mlo <= mhi) {
                        Mm |= 1u;
                        if (gl == 0) sm.mmeta[slot] = fr_pack(mlo, mhi);
                    }
                    if (flg & 1u) G1m |= 1u;
                    if (flg & 2u) G2m |= 1u;
                    R1a = Rcur, R2 = Rcur;
                }
                __syncwarp();
            }
            // ---------------- backtrace + forward replay (groups that found an alignment) ----------------
            const bool aligned = usable && !identical && !overflow;
            uint8_t *ops = reinterpret_cast<uint8_t *>(sm.mrow); // the rows are dead now
            constexpr int OPS_CAP = F_MROWS * W * 2;
            int nops = 0, n_del = 0;
            bool ops_over = false;
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
                        if (nops >= OPS_CAP) ops_over = true, walking = false;
                        else {
                            if (gl == 0) ops[nops] = (uint8_t)op;
                            ++nops;
                            walking = !(comp < 0 && cs == 0);
                        }
                    }
                }
            }
            if (ops_over) overflow = true;
            __syncwarp();
            {
                ClipAcc acc;
                acc.init(m + n_del, p.clip);
                int v = 0, h = 0, i = nops - 1;
                bool in_m = true, replay = aligned && !ops_over;
                while (__any_sync(TDB_FULL, replay)) {
                    const int eq = extend_groups<G>(pw, tw, psh, tsh, n, m, v, h, replay && in_m, gl, gshift, lane);
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
                if (aligned && !ops_over) {
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
                for (int o = 1; o < G; o <<= 1) ext_pair += __shfl_xor_sync(TDB_FULL, ext_pair, o);
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
