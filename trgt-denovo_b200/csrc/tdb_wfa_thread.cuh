// tdb_wfa_thread.cuh -- thread-per-pair tier for the cheapest alignments (sm_100a).
//
// Most pairs that reach the wavefront kernels are a read with two or three sequencing errors against
// its allele: final penalty <= 24, wavefronts a handful of diagonals wide.  The lockstep tier spends
// a warp instruction on 8 lanes per pair of which 1..5 hold a live diagonal; here ONE THREAD owns a
// pair and walks its own diagonals, so every lane of a warp instruction is a different alignment.
//
// What makes that fit: with the production penalties (x 8, o1+e1 6, e1 2, o2+e2 25) nothing below
// score 25 can use the second gap piece and only even scores are reachable, so up to score 24 the
// recurrence is single-piece gap-affine over 12 score steps, diagonals |k| <= (s-4)/2 <= 10.  Per
// thread, in shared memory interleaved by thread (element e of thread t at e*96 + t, so a thread's
// column stays in its own bank): a 5-slot M ring + in-place I1/D1 (147 int16) and BOTH SEQUENCES,
// staged once with independent loads (2 x 21 words: sequences up to 304 bases) -- the extension
// compares then never wait for global memory (an ncu capture of the first version, which read the
// packed stream directly, had a third of its stall samples on those loads).  4 provenance bits per
// computed cell (30 words) live in thread-local memory: written once per score, read by the backtrace.
// Scores 25..30: the second gap piece appears (O2 + E2 = 25), but up to score 30 it consists of the two pure chains
// that open at M[0] -- one insertion cell on diagonal +(s - 24), one deletion cell on diagonal -(s - 24) per score --
// and the odd scores hold nothing else (their first-piece sources M[s-6], M[s-8], I1/D1[s-2] would have to be odd and
// older than 25).  Those chains are carried in six registers and computed with the same rules as everything else
// (open / extend, validity, trimming, cut-off, equate, provenance priority X > D2 > D1 > I2 > I1); an odd-score M
// wavefront is never read again below score 31, so it is extended, counted and dropped.
// A pair that is not finished at score TH_SMAX is handed to the lockstep tier, which starts it over.
//
// Semantics are those of wfa_align_warp (tdb_wfa.cuh) step for step: compute + trim (A.2),
// wf-adaptive cut-off + equate (A.3), provenance tie-breaks X > D1 > I1 and matches re-extended only
// in the M component (A.4), flank-clipped score on the fly (src/aligner.rs:320-376).
#pragma once
#include "tdb_kernels.cuh"
#include <type_traits>

namespace tdb {

#ifndef TDB_TH_NT
#define TDB_TH_NT 96
#endif
constexpr int TH_NT = TDB_TH_NT;            // threads (= pairs in flight) per CTA; a multiple of 32
constexpr int TH_KM = 13;                   // diagonals -TH_KM .. TH_KM
constexpr int TH_W = 2 * TH_KM + 1;
constexpr int TH_SMAX = 30;                 // last score this tier computes
constexpr int TH_MSLOTS = 5;                // M ring: scores s, s-2, .., s-8
constexpr int TH_HW = TH_MSLOTS * TH_W + 2 * TH_W; // int16 elements per thread
constexpr int TH_STEPS = TH_SMAX / 2 - 2;   // scores 6, 8, .., 30
constexpr int TH_PW = 4;                    // provenance words per score step (27 nibbles)
static_assert(TH_SMAX <= 30 && (TH_SMAX - 4) / 2 <= TH_KM, "second-piece handling below covers scores 25..30 only");
constexpr int TH_MAXLEN = ML_MAXLEN;        // longest sequence this tier takes
constexpr size_t th_smem_bytes() { return (size_t)TH_NT * (TH_HW * 2 + 2 * TH_SEQW * 4); }
static_assert(15 + TH_MAXLEN + 16 <= 16 * TH_SEQW, "staging window");

// 16 bases from position pos of a sequence staged at words [e0, e0 + TH_SEQW) of this thread's column
__device__ __forceinline__ uint32_t th_get16(const uint32_t *sw, int e0, int pos) {
    const int i = e0 + (pos >> 4);
    return __funnelshift_r(sw[i * TH_NT], sw[(i + 1) * TH_NT], (pos & 15) << 1);
}
// greedy match extension from (v, h) on the staged sequences (psh / tsh: base shift inside the first word)
__device__ __forceinline__ int th_extend(const uint32_t *sw, int psh, int tsh, int n, int m, int v, int h) {
    const int room = min(n - v, m - h);
    int eq = 0;
    while (eq < room) {
        const uint32_t x = th_get16(sw, 0, psh + v + eq) ^ th_get16(sw, TH_SEQW, tsh + h + eq);
        if (x) {
            eq += (__ffs(x) - 1) >> 1;
            break;
        }
        eq += 16;
    }
    return min(eq, room);
}

template <int X, int O1, int E1, int O2, int E2>
__global__ void __launch_bounds__(TH_NT) k_align_thread(const AlignParams p) {
    static_assert(X == 8 && O1 + E1 == 6 && E1 == 2 && O2 + E2 == 25 && E2 == 1, "ring geometry is built for 8/4/2/24/1");
    constexpr int NULLV = OffTraits<int16_t>::NULLV;
    constexpr int HUGE_D = 1 << 30;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    int16_t *hw = reinterpret_cast<int16_t *>(smem_raw) + threadIdx.x;
    uint32_t *sw = reinterpret_cast<uint32_t *>(smem_raw + (size_t)TH_NT * TH_HW * 2) + threadIdx.x;
    uint8_t prov[TH_STEPS * TH_W]; // one nibble per computed cell, one byte each: written once, read by the backtrace (thread-local memory)
#define TH_M(slot, k) hw[((slot) * TH_W + (k) + TH_KM) * TH_NT]
#define TH_I(k) hw[(TH_MSLOTS * TH_W + (k) + TH_KM) * TH_NT]
#define TH_D(k) hw[((TH_MSLOTS + 1) * TH_W + (k) + TH_KM) * TH_NT]
#define TH_P(t, pi) prov[(t) * TH_W + (pi)]
    const int lane = threadIdx.x & 31;
    const ItemSrc src = item_src(p);
    const Penalties pn = {X, O1, E1, O2, E2};
    const Heur heur = p.heur;
    unsigned long long tot_cells = 0, tot_ext = 0, tot_cols = 0, tot_done = 0;
    for (;;) {
        unsigned b = 0;
        if (lane == 0) b = atomicAdd(p.work_counter, 32u);
        b = __shfl_sync(TDB_FULL, b, 0);
        if (b >= src.total) break;
        const unsigned it = b + lane;
        if (it < src.total) {
            const uint32_t idx = item_at(p, src, it);
            const uint32_t ip = p.pair_p[idx], jt = p.pair_t[idx];
            const int n = (int)p.seq_len[ip], m = (int)p.seq_len[jt];
            const uint64_t poff = p.seq_off[ip], toff = p.seq_off[jt];
            const int psh = (int)(poff & 15), tsh = (int)(toff & 15);
            { // stage both sequences: the only global loads of the alignment, all independent
                const uint32_t *gp = p.packed + (poff >> 4), *gt = p.packed + (toff >> 4);
                const int np = (psh + n + 31) >> 4, nt = (tsh + m + 31) >> 4; // words incl. the funnel-shift neighbour
#pragma unroll
                for (int i = 0; i < TH_SEQW; ++i) {
                    if (i < np) sw[i * TH_NT] = __ldg(gp + i);
                    if (i < nt) sw[(TH_SEQW + i) * TH_NT] = __ldg(gt + i);
                }
            }
            const int k_end = m - n;
            unsigned ext = 0, cells = 0;
            int pen = -1, clipped = 0, cols = 0;
            // ---- score 0 ---------------------------------------------------------------------------
            const int m0 = th_extend(sw, psh, tsh, n, m, 0, 0);
            ext += (unsigned)(m0 >> 4) + 1u;
            if (k_end == 0 && m0 >= m) {
                pen = 0, cols = m;
            } else {
                TH_M(0, 0) = (int16_t)m0;
                // wavefront ranges (lo, span = hi - lo + 1; span 0 = null) of M[s-2], M[s-4], M[s-6], M[s-8], I1/D1[s-2]
                int ml2 = 1, ms2 = 0, ml4 = 1, ms4 = 0, ml6 = 0, ms6 = 1, ml8 = 1, ms8 = 0;
                int il = 1, is = 0, dl = 1, ds = 0;
                int steps_wait = heur.steps;
                if (heur.kind == 1) --steps_wait;
                // second gap piece: the cell of I2 / D2 of the previous score (diagonal, un-extended offset), if any
                bool i2ok = false, d2ok = false;
                int i2k = 0, i2v = 0, d2k = 0, d2v = 0;
                unsigned p2 = 0; // bit 2t: M[6 + 2t] on the I2 chain's diagonal came from I2; bit 2t + 1: same for D2
                for (int s = 6; s <= TH_SMAX && pen < 0; s += 2) {
                    if (s >= O2 + E2 + 1) {
                        // ---- the odd score s - 1: only the two second-piece cells exist ----
                        const int t = s - 1;
                        int kI = 0, vI = NULLV, kD = 0, vD = NULLV;
                        bool hI = false, hD = false;
                        if (t == O2 + E2) hI = hD = true, kI = 1, vI = m0 + 1, kD = -1, vD = m0; // open from M[0] = {0: m0}
                        else {
                            if (i2ok) hI = true, kI = i2k + 1, vI = i2v + 1;
                            if (d2ok) hD = true, kD = d2k - 1, vD = d2v;
                        }
                        i2ok = d2ok = false;
                        if (hI || hD) {
                            int lo = t == O2 + E2 ? -1 : (hD ? kD : kI), hi = t == O2 + E2 ? 1 : (hI ? kI : kD);
                            lo = max(lo, -n - 1), hi = min(hi, m + 1);
                            cells += (unsigned)(hi - lo + 1);
                            const bool okI = hI && kI <= hi && (unsigned)vI <= (unsigned)m && (unsigned)(vI - kI) <= (unsigned)n;
                            const bool okD = hD && kD >= lo && (unsigned)vD <= (unsigned)m && (unsigned)(vD - kD) <= (unsigned)n;
                            int dI = HUGE_D, dD = HUGE_D;
                            bool fin_odd = false;
                            if (okI) {
                                const int eq = th_extend(sw, psh, tsh, n, m, vI - kI, vI);
                                ext += (unsigned)(eq >> 4) + 1u;
                                const int mv = vI + eq;
                                dI = max(n - (mv - kI), m - mv);
                                fin_odd |= kI == k_end && mv >= m;
                            }
                            if (okD) {
                                const int eq = th_extend(sw, psh, tsh, n, m, vD - kD, vD);
                                ext += (unsigned)(eq >> 4) + 1u;
                                const int mv = vD + eq;
                                dD = max(n - (mv - kD), m - mv);
                                fin_odd |= kD == k_end && mv >= m;
                            }
                            if (fin_odd) break; // an odd final penalty: not this tier's (never seen: the even family gets there first)
                            int mlo = okD ? kD : kI, mhi = okI ? kI : kD; // trimmed M range (empty when neither cell is valid)
                            bool m_any = okI || okD;
                            if (heur.kind == 1 && m_any) {
                                --steps_wait;
                                if (steps_wait <= 0 && mhi - mlo + 1 >= heur.min_len) {
                                    const int md = min(max(n, m), min(dI, dD));
#define TH_ODD_D(kk) ((kk) == kI && okI ? dI : ((kk) == kD && okD ? dD : HUGE_D))
                                    const int top_limit = min(k_end, mhi);
                                    int new_lo = mlo;
                                    if (top_limit > mlo) {
                                        int k = mlo;
                                        for (; k < top_limit; ++k)
                                            if (TH_ODD_D(k) - md <= heur.max_dist) break;
                                        new_lo = k;
                                    }
                                    const int bot_limit = max(k_end, new_lo);
                                    int new_hi = mhi;
                                    if (bot_limit < mhi) {
                                        int k = mhi;
                                        for (; k > bot_limit; --k)
                                            if (TH_ODD_D(k) - md <= heur.max_dist) break;
                                        new_hi = k;
                                    }
#undef TH_ODD_D
                                    mlo = new_lo, mhi = new_hi;
                                    steps_wait = heur.steps;
                                }
                                m_any = mhi >= mlo;
                                // equate: the gap wavefronts never reach beyond M
                                if (m_any) {
                                    i2ok = okI && kI >= mlo && kI <= mhi;
                                    d2ok = okD && kD >= mlo && kD <= mhi;
                                } else {
                                    i2ok = okI, d2ok = okD; // (the general code leaves the gap ranges alone when M ends up empty)
                                }
                            } else {
                                i2ok = okI, d2ok = okD;
                            }
                            i2k = kI, i2v = vI, d2k = kD, d2v = vD;
                        }
                    }
                    const int hh = s >> 1;
                    const int cur = hh % TH_MSLOTS, s6 = (hh + 2) % TH_MSLOTS, s8 = (hh + 1) % TH_MSLOTS;
                    int nml = 1, nms = 0, nil = 1, nis = 0, ndl = 1, nds = 0; // ranges of score s
                    if (ms8 | ms6 | is | ds | (int)i2ok | (int)d2ok) {
                        int lo = INT32_MAX, hi = INT32_MIN;
                        if (ms8) lo = min(lo, ml8), hi = max(hi, ml8 + ms8 - 1);
                        if (ms6) lo = min(lo, ml6 - 1), hi = max(hi, ml6 + ms6);
                        if (is) lo = min(lo, il + 1), hi = max(hi, il + is);
                        if (ds) lo = min(lo, dl - 1), hi = max(hi, dl + ds - 2);
                        if (i2ok) lo = min(lo, i2k + 1), hi = max(hi, i2k + 1);
                        if (d2ok) lo = min(lo, d2k - 1), hi = max(hi, d2k - 1);
                        lo = max(lo, -n - 1);
                        hi = min(hi, m + 1);
                        cells += (unsigned)(hi - lo + 1);
                        const bool has_i1 = ms6 || is, has_d1 = ms6 || ds;
                        int tl_m = INT32_MAX, th_m = INT32_MIN, tl_i = INT32_MAX, th_i = INT32_MIN, tl_d = INT32_MAX,
                            th_d = INT32_MIN;
                        int min_d = HUGE_D;
                        bool fin = false;
                        uint32_t vmask = 0; // vmask: diagonals (bit k + TH_KM) whose M offset lies inside the matrix
                        // the second piece's cells of this score: extensions of the previous score's (none before score 26)
                        const int kI = i2k + 1, vI = i2v + 1, kD = d2k - 1, vD = d2v;
                        const bool hI = i2ok && kI <= hi, hD = d2ok && kD >= lo;
                        uint8_t *const pv = &TH_P(hh - 3, TH_KM); // provenance row of this score, indexed by diagonal
                        // One body, compiled twice: the second piece's two cells only exist from score 26 on (TWO), and the
                        // compares that look for them are a sixth of the loop.  Running pointers into this thread's column:
                        // q6 / q8 / qc at diagonal k of M[s-6] / M[s-8] / M[s], qg at diagonal k of I1 (D1 is one row further).
                        auto cell_loop = [&](auto two_tag) {
                            constexpr bool TWO = decltype(two_tag)::value;
                            const int16_t *q6 = &TH_M(s6, lo), *q8 = &TH_M(s8, lo);
                            int16_t *qc = &TH_M(cur, lo), *qg = &TH_I(lo);
                            constexpr int DROW = TH_W * TH_NT; // TH_D(k) == (&TH_I(k))[DROW]
                            uint32_t kbit = 1u << (lo + TH_KM);
                            // I1 is updated in place: the old I1[k-1] travels in a register
                            int carry_i = (unsigned)(lo - 1 - il) < (unsigned)is ? (int)qg[-TH_NT] : NULLV;
                            for (int k = lo; k <= hi; ++k, q6 += TH_NT, q8 += TH_NT, qc += TH_NT, qg += TH_NT, kbit <<= 1) {
                                const int old_i = (unsigned)(k - il) < (unsigned)is ? (int)qg[0] : NULLV;
                                int nib = 0, o, e;
                                o = (unsigned)(k - 1 - ml6) < (unsigned)ms6 ? (int)q6[-TH_NT] : NULLV;
                                e = carry_i;
                                carry_i = old_i;
                                int ins1;
                                if (e >= o) ins1 = e, nib |= 4; else ins1 = o;
                                ins1 += 1;
                                o = (unsigned)(k + 1 - ml6) < (unsigned)ms6 ? (int)q6[TH_NT] : NULLV;
                                e = (unsigned)(k + 1 - dl) < (unsigned)ds ? (int)qg[DROW + TH_NT] : NULLV;
                                int del1;
                                if (e >= o) del1 = e, nib |= 8; else del1 = o;
                                const int mis = ((unsigned)(k - ml8) < (unsigned)ms8 ? (int)q8[0] : NULLV) + 1;
                                int mv = max(max(ins1, del1), mis);
                                int sc = 0; // later tests override: priority X > D2 > D1 > I2 > I1
                                if (TWO) {
                                    const int ins2 = hI && k == kI ? vI : NULLV + 1, del2 = hD && k == kD ? vD : NULLV;
                                    mv = max(mv, max(ins2, del2));
                                    bool f_i2 = false, f_d2 = false;
                                    if (mv == ins1) sc = 1;
                                    if (mv == ins2) f_i2 = true;
                                    if (mv == del1) sc = 2, f_i2 = false;
                                    if (mv == del2) f_d2 = true, f_i2 = false;
                                    if (mv == mis) sc = 3, f_i2 = f_d2 = false;
                                    if (hI && k == kI && f_i2) p2 |= 1u << (2 * (hh - 3));
                                    if (hD && k == kD && f_d2) p2 |= 2u << (2 * (hh - 3));
                                } else {
                                    if (mv == ins1) sc = 1;
                                    if (mv == del1) sc = 2;
                                    if (mv == mis) sc = 3;
                                }
                                nib |= sc;
                                pv[k] = (uint8_t)nib;
                                if ((unsigned)mv > (unsigned)m || (unsigned)(mv - k) > (unsigned)n) mv = NULLV;
                                if (mv >= 0) vmask |= kbit, tl_m = min(tl_m, k), th_m = k;
                                qc[0] = (int16_t)mv;
                                if (has_i1) {
                                    qg[0] = (int16_t)ins1;
                                    if ((unsigned)ins1 <= (unsigned)m && (unsigned)(ins1 - k) <= (unsigned)n) tl_i = min(tl_i, k), th_i = k;
                                }
                                if (has_d1) {
                                    qg[DROW] = (int16_t)del1;
                                    if ((unsigned)del1 <= (unsigned)m && (unsigned)(del1 - k) <= (unsigned)n) tl_d = min(tl_d, k), th_d = k;
                                }
                            }
                        };
                        if (s >= O2 + E2 + 1) cell_loop(std::true_type{}); // (the score is the same in every lane still at work)
                        else cell_loop(std::false_type{});
                        // Greedy extension of the live diagonals (trimming never changes offsets, A.1 step 1), as ONE
                        // loop over 16-base compares: a lane moves on to its next diagonal as soon as the current one
                        // stops matching, so a warp pays the longest per-lane TOTAL, not the sum of per-diagonal maxima.
                        // (running word pointers into the staged sequences; the upper word of a round is the lower word of the
                        // next: two shared-memory loads per round)
                        {
                            int k = 0, h = 0, room = 0, eq = 0, pqs = 0, tqs = 0;
                            const uint32_t *pq = sw, *tq = sw;
                            uint32_t plo = 0, tlo = 0;
                            bool active = false;
                            for (;;) {
                                if (!active) {
                                    if (!vmask) break;
                                    const int pi = __ffs(vmask) - 1;
                                    vmask &= vmask - 1;
                                    k = pi - TH_KM;
                                    h = (int)TH_M(cur, k);
                                    const int v = h - k;
                                    room = min(n - v, m - h), eq = 0;
                                    pq = sw + ((psh + v) >> 4) * TH_NT, tq = sw + (TH_SEQW + ((tsh + h) >> 4)) * TH_NT;
                                    pqs = ((psh + v) & 15) << 1, tqs = ((tsh + h) & 15) << 1;
                                    plo = *pq, tlo = *tq;
                                    active = true;
                                }
                                bool stop = eq >= room;
                                if (!stop) {
                                    pq += TH_NT, tq += TH_NT;
                                    const uint32_t phi = *pq, thi = *tq;
                                    const uint32_t x = __funnelshift_r(plo, phi, pqs) ^ __funnelshift_r(tlo, thi, tqs);
                                    plo = phi, tlo = thi;
                                    if (x) eq += (__ffs(x) - 1) >> 1, stop = true;
                                    else eq += 16, stop = eq >= room;
                                }
                                if (stop) {
                                    eq = min(eq, room);
                                    ext += (unsigned)(eq >> 4) + 1u;
                                    const int mv = h + eq;
                                    TH_M(cur, k) = (int16_t)mv;
                                    min_d = min(min_d, max(n - (mv - k), m - mv));
                                    fin |= (k == k_end) && (mv >= m);
                                    active = false;
                                }
                            }
                        }
                        // this score's second-piece cells: inside the matrix? (their trimmed ranges, A.2)
                        bool nI = hI && (unsigned)vI <= (unsigned)m && (unsigned)(vI - kI) <= (unsigned)n;
                        bool nD = hD && (unsigned)vD <= (unsigned)m && (unsigned)(vD - kD) <= (unsigned)n;
                        if (fin) {
                            pen = s;
                            break;
                        }
                        if (tl_m <= th_m) nml = tl_m, nms = th_m - tl_m + 1;
                        if (tl_i <= th_i) nil = tl_i, nis = th_i - tl_i + 1;
                        if (tl_d <= th_d) ndl = tl_d, nds = th_d - tl_d + 1;
                        // ---- wf-adaptive cut-off + equate (A.3) ------------------------------------------
                        if (heur.kind == 1 && nms) {
                            int mlo = nml, mhi = nml + nms - 1;
                            --steps_wait;
                            if (steps_wait <= 0 && nms >= heur.min_len) {
                                const int md = min(max(n, m), min_d);
                                const int top_limit = min(k_end, mhi);
                                int new_lo = mlo;
                                if (top_limit > mlo) {
                                    int k = mlo;
                                    for (; k < top_limit; ++k) {
                                        const int off = (int)TH_M(cur, k);
                                        const int d = off >= 0 ? max(n - (off - k), m - off) : HUGE_D;
                                        if (d - md <= heur.max_dist) break;
                                    }
                                    new_lo = k; // first diagonal close enough, or top_limit
                                }
                                const int bot_limit = max(k_end, new_lo);
                                int new_hi = mhi;
                                if (bot_limit < mhi) {
                                    int k = mhi;
                                    for (; k > bot_limit; --k) {
                                        const int off = (int)TH_M(cur, k);
                                        const int d = off >= 0 ? max(n - (off - k), m - off) : HUGE_D;
                                        if (d - md <= heur.max_dist) break;
                                    }
                                    new_hi = k;
                                }
                                mlo = new_lo, mhi = new_hi;
                                steps_wait = heur.steps;
                            }
                            nml = mlo, nms = mhi - mlo + 1;
                            if (nms > 0) { // equate: the gap wavefronts never reach beyond M
                                int a = max(nil, mlo), z = min(nil + nis - 1, mhi);
                                if (nis) nil = a, nis = max(0, z - a + 1);
                                a = max(ndl, mlo), z = min(ndl + nds - 1, mhi);
                                if (nds) ndl = a, nds = max(0, z - a + 1);
                                nI = nI && kI >= mlo && kI <= mhi;
                                nD = nD && kD >= mlo && kD <= mhi;
                            } else {
                                nml = 1, nms = 0;
                            }
                        }
                        i2ok = nI, i2k = kI, i2v = vI, d2ok = nD, d2k = kD, d2v = vD;
                    }
                    ml8 = ml6, ms8 = ms6, ml6 = ml4, ms6 = ms4, ml4 = ml2, ms4 = ms2, ml2 = nml, ms2 = nms;
                    il = nil, is = nis, dl = ndl, ds = nds;
                }
                // ---- backtrace (A.4) + forward replay with the clipped score --------------------------
                if (pen > 0) {
                    unsigned long long ops = 0;
                    int nops = 0, n_del = 0, comp = 0 /* 0 M, 1 I1, 2 D1 */, cs = pen, k = k_end;
                    while (!(comp == 0 && cs == 0)) {
                        const int t = (cs >> 1) - 3, pi = k + TH_KM;
                        if (t < 0 || (unsigned)pi >= (unsigned)TH_W || nops >= 32) {
                            pen = -1; // never expected: leave the pair to the next tier
                            break;
                        }
                        const int nib = (int)TH_P(t, pi);
                        int op;
                        const int j2 = cs - (O2 + E2 - 1); // length of the second-piece chain that ends at score cs
                        if (comp == 0 && j2 >= 2 && ((k == j2 && ((p2 >> (2 * t)) & 1u)) || (k == -j2 && ((p2 >> (2 * t)) & 2u)))) {
                            // M came from I2 / D2: the whole chain back to its opening at M[0] (A.4), one gap of j2 bases
                            const bool is_ins = k > 0;
                            if (nops + j2 + 1 > 32) {
                                pen = -1;
                                break;
                            }
                            ops |= (unsigned long long)OP_CLOSE << (2 * nops), ++nops;
                            for (int q = 0; q < j2; ++q) ops |= (unsigned long long)(is_ins ? OP_I : OP_D) << (2 * nops), ++nops;
                            if (!is_ins) n_del += j2;
                            cs = 0, k = 0;
                            continue;
                        }
                        if (comp == 0) {
                            const int sc = nib & 3;
                            if (sc == 3) op = OP_X, cs -= X;
                            else if (sc == 0) {
                                pen = -1;
                                break;
                            } else op = OP_CLOSE, comp = sc;
                        } else if (comp == 1) {
                            op = OP_I;
                            if (nib & 4) cs -= E1; else cs -= O1 + E1, comp = 0;
                            k -= 1;
                        } else {
                            op = OP_D;
                            ++n_del;
                            if (nib & 8) cs -= E1; else cs -= O1 + E1, comp = 0;
                            k += 1;
                        }
                        ops |= (unsigned long long)op << (2 * nops);
                        ++nops;
                    }
                    if (pen > 0) {
                        ClipAcc acc;
                        acc.init(m + n_del, p.clip);
                        int v = 0, h = 0;
                        bool in_m = true;
                        for (int i = nops - 1; i >= -1; --i) {
                            if (in_m) {
                                const int eq = th_extend(sw, psh, tsh, n, m, v, h);
                                acc.emit('M', eq, pn);
                                v += eq, h += eq;
                            }
                            if (i < 0) break;
                            const int op = (int)(ops >> (2 * i)) & 3;
                            if (op == OP_CLOSE) {
                                in_m = true;
                                continue;
                            }
                            acc.emit(op == OP_X ? 'X' : (op == OP_I ? 'I' : 'D'), 1, pn);
                            if (op == OP_X) ++v, ++h;
                            else if (op == OP_I) ++h, in_m = false;
                            else ++v, in_m = false;
                        }
                        clipped = acc.finish(pn);
                        cols = acc.col;
                    }
                }
            }
            if (pen >= 0) {
                p.out_score[idx] = clipped;
                if (p.out_status) p.out_status[idx] = TDB_STATUS_ALG_COMPLETED;
                if (p.out_penalty) p.out_penalty[idx] = pen;
                if (p.work) {
                    p.work[3 * (size_t)idx] = cells, p.work[3 * (size_t)idx + 1] = ext;
                    p.work[3 * (size_t)idx + 2] = (uint32_t)cols;
                }
                tot_cells += cells, tot_ext += ext, tot_cols += (unsigned)cols, ++tot_done;
            } else {
                const unsigned j = atomicAdd(p.next_count, 1u);
                p.next_todo[j] = idx;
            }
        }
        __syncwarp();
    }
#undef TH_M
#undef TH_I
#undef TH_D
#undef TH_P
    for (int o = 16; o; o >>= 1) {
        tot_cells += __shfl_xor_sync(TDB_FULL, tot_cells, o), tot_ext += __shfl_xor_sync(TDB_FULL, tot_ext, o);
        tot_cols += __shfl_xor_sync(TDB_FULL, tot_cols, o), tot_done += __shfl_xor_sync(TDB_FULL, tot_done, o);
    }
    if (lane == 0 && p.counters) {
        atomicAdd(p.counters + 0, tot_cells), atomicAdd(p.counters + 1, tot_ext);
        atomicAdd(p.counters + 2, tot_cols), atomicAdd(p.counters + 3, tot_done);
    }
}

} // namespace tdb
