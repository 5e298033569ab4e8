// tdb_wfa.cuh -- warp-cooperative gap-affine-2-piece wavefront alignment (sm_100a).
//
// One warp aligns one (read, target) pair: lanes own diagonals of the current wavefront, the
// wavefront history (26-score window for M, e+1 for the gap components) lives in a ring indexed by
// (score, k mod W) -- in shared memory for pairs of moderate width (k_align_warp) or in a per-warp slice of
// global scratch for long / wide pairs (tiers 1,2).  One provenance byte per computed cell is kept so
// that the alignment path can be rebuilt on the device and the flank-clipped score
// (reference: src/aligner.rs:320-357) produced without ever materialising a CIGAR in HBM.
//
// Semantics replaced (reference call sites): WFAligner::align_end_to_end (src/aligner.rs:248-260) ->
// wfa2::wavefront_align in the configuration of src/commands/shared.rs:41-51, followed by
// WFAligner::cigar_score_clipped (src/aligner.rs:359-376).  WFA2-lib behaviour reproduced:
// SURVEY.md Appendix A.1 (loop), A.2 (compute + trim), A.3 (wf-adaptive + equate), A.4 (tie-breaks,
// matches re-extended only in the M component).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tdb {

#define TDB_FULL 0xffffffffu

struct Penalties {
    int x, o1, e1, o2, e2;
};
struct Heur {
    int kind, min_len, max_dist, steps;
};

// A sequence as the kernels see it: a window of the 2-bit packed blob (the byte at blob position x is
// packed into word x/16 at bits 2*(x%16); base i of the sequence is at blob position seq_off + i, so
// w points at word seq_off/16 and sh = seq_off%16 is added to every base index),
// plus the raw bytes (b != nullptr) when the sequence holds a byte outside ACGT.  A pair is compared
// 16 bases per instruction on the packed words when both sides are pure ACGT; otherwise byte by byte
// (raw byte equality, 'N'=='N', case-sensitive), the pure-ACGT side decoded from its packed form.
struct SeqView {
    const uint32_t *w;
    const uint8_t *b;
    int len;
    int sh;
};
__device__ __forceinline__ uint32_t base_at(const SeqView &S, int i) {
    if (S.b) return S.b[i];
    const int x = S.sh + i;
    const uint32_t code = (__ldg(S.w + (x >> 4)) >> ((x & 15) << 1)) & 3u;
    return (0x47544341u >> (code << 3)) & 0xffu; // (byte >> 1) & 3: A 0, C 1, T 2, G 3
}

// provenance byte: bits 0-2 source of M, bits 3..6: I1,I2,D1,D2 came from "extend" (else "open")
enum { PM_NONE = 0, PM_I1 = 1, PM_I2 = 2, PM_D1 = 3, PM_D2 = 4, PM_X = 5 };
enum { PB_I1E = 8, PB_I2E = 16, PB_D1E = 32, PB_D2E = 64 };
enum { OP_X = 0, OP_I = 1, OP_D = 2, OP_CLOSE = 3 };
enum { CI1 = 0, CI2 = 1, CD1 = 2, CD2 = 3 };

constexpr int MS = 32; // M ring slots (>= max(x, o1+e1, o2+e2) + 1, power of two)
constexpr int GS = 4;  // gap ring slots (>= e + 1)

// result of one alignment attempt
enum { WFA_OK = 0, WFA_OVERFLOW = 1 };

template <typename OffT>
struct OffTraits;
template <>
struct OffTraits<int16_t> {
    static constexpr int NULLV = -16384;
};
template <>
struct OffTraits<int32_t> {
    static constexpr int NULLV = INT32_MIN / 2;
};

// Storage handed to the aligner (shared memory for k_align_warp, global scratch for k_align_global).
template <typename OffT, typename StepLoT, typename StepBaseT>
struct Storage {
    OffT *M;           // [MS][W]
    OffT *G;           // [4][GS][W]
    int2 *mmeta;       // [MS]   (lo, hi); lo > hi = null
    int2 *gmeta;       // [4*GS]
    StepLoT *step_lo;  // [s_cap] computed lo of score s
    StepBaseT *step_base; // [s_cap] provenance index of diagonal lo
    uint8_t *prov;     // [prov_cap]
    uint8_t *ops;      // [ops_cap] backtrace op stream (may alias M: dead by then)
    int wlog2, s_cap, ops_cap;
    unsigned prov_cap;
};

struct WorkCount {
    unsigned long long cells, ext16, columns;
};

__device__ __forceinline__ uint32_t get16(const uint32_t *w, int pos) {
    const int i = pos >> 4;
    const uint32_t lo = __ldg(w + i), hi = __ldg(w + i + 1);
    return __funnelshift_r(lo, hi, (pos & 15) << 1);
}

// Per-lane greedy match extension from (v,h): while v<n && h<m && pattern[v]==text[h].
__device__ __forceinline__ int extend_lane(const SeqView &P, const SeqView &T, int v, int h,
                                           unsigned long long &ncmp) {
    const int room = min(P.len - v, T.len - h);
    int eq = 0;
    if (!P.b && !T.b) {
        while (eq < room) {
            const uint32_t x = get16(P.w, P.sh + v + eq) ^ get16(T.w, T.sh + h + eq);
            if (x) {
                eq += (__ffs(x) - 1) >> 1;
                break;
            }
            eq += 16;
        }
        eq = min(eq, room);
        ncmp += (eq >> 4) + 1; // E of SURVEY.md 8(d): 16-base compares, at least one per live diagonal
        return eq;
    }
    while (eq < room && base_at(P, v + eq) == base_at(T, h + eq)) ++eq;
    ncmp += (eq >> 4) + 1;
    return eq;
}

// Whole-warp extension of ONE diagonal (v,h uniform): lane l compares bases [16l, 16l+16) per round.
__device__ __forceinline__ int extend_coop(const SeqView &P, const SeqView &T, int v, int h, int lane) {
    const int room = min(P.len - v, T.len - h);
    if (!P.b && !T.b) {
        for (int base = 0; base < room; base += 512) {
            const int off = base + lane * 16;
            uint32_t x = 1u;
            if (off < room) x = get16(P.w, P.sh + v + off) ^ get16(T.w, T.sh + h + off);
            const unsigned bal = __ballot_sync(TDB_FULL, x != 0);
            if (bal) {
                const int l = __ffs(bal) - 1;
                const uint32_t xl = __shfl_sync(TDB_FULL, x, l);
                return min(base + l * 16 + ((__ffs(xl) - 1) >> 1), room);
            }
        }
        return room;
    }
    for (int base = 0; base < room; base += 32) {
        const int off = base + lane;
        const bool ne = off >= room || base_at(P, v + off) != base_at(T, h + off);
        const unsigned bal = __ballot_sync(TDB_FULL, ne);
        if (bal) return min(base + __ffs(bal) - 1, room);
    }
    return room;
}

// cost of a run inside the clip window, src/aligner.rs:303-318
__device__ __forceinline__ int run_cost(int op, int len, const Penalties &pn) {
    if (op == 'M') return 0;
    if (op == 'X') return len * pn.x;
    return min(pn.o1 + pn.e1 * len, pn.o2 + pn.e2 * len);
}

struct ClipAcc {
    int wb, we, col, run_op, run_len, score;
    __device__ __forceinline__ void init(int total_cols, int clip) {
        wb = clip, we = total_cols - clip, col = 0, run_op = 0, run_len = 0, score = 0;
    }
    __device__ __forceinline__ void emit(int op, int cnt, const Penalties &pn) {
        const int a = max(col, wb), b = min(col + cnt, we);
        if (b > a) {
            if (op == run_op) run_len += b - a;
            else {
                if (run_len) score -= run_cost(run_op, run_len, pn);
                run_op = op, run_len = b - a;
            }
        }
        col += cnt;
    }
    __device__ __forceinline__ int finish(const Penalties &pn) {
        if (run_len) score -= run_cost(run_op, run_len, pn);
        run_len = 0;
        return score;
    }
};

// One input wavefront of a compute step.
template <typename OffT>
struct InWf {
    const OffT *p;
    int lo;
    unsigned span1; // hi - lo + 1, 0 when null
    __device__ __forceinline__ bool null() const { return span1 == 0; }
    __device__ __forceinline__ int rd(int k, int mask) const {
        return ((unsigned)(k - lo) < span1) ? (int)p[k & mask] : OffTraits<OffT>::NULLV;
    }
};

template <typename OffT>
__device__ __forceinline__ InWf<OffT> make_in(const OffT *ring, const int2 *meta, int s, int slots_mask, int wlog2) {
    InWf<OffT> w;
    w.p = ring, w.lo = 1, w.span1 = 0;
    if (s >= 0) {
        const int slot = s & slots_mask;
        const int2 mm = meta[slot];
        if (mm.x <= mm.y) {
            w.p = ring + ((size_t)slot << wlog2);
            w.lo = mm.x;
            w.span1 = (unsigned)(mm.y - mm.x + 1);
        }
    }
    return w;
}

// Aligns one pair with the whole warp.  Returns WFA_OK and (penalty, clipped) or WFA_OVERFLOW when the
// storage is too small (caller re-queues the pair on the next tier).
// WANT_CIGAR: additionally write the CIGAR columns to cigar_out (forward order) and *cigar_cols.
template <typename OffT, typename StepLoT, typename StepBaseT, bool WANT_CIGAR>
__device__ int wfa_align_warp(const SeqView &P, const SeqView &T, const Penalties &pn, const Heur &heur,
                              const int clip, const Storage<OffT, StepLoT, StepBaseT> &st, const int lane,
                              int &out_penalty, int &out_clipped, WorkCount &wc,
                              unsigned long long &lane_ext16, uint8_t *cigar_out, uint32_t *cigar_cols) {
    constexpr int NULLV = OffTraits<OffT>::NULLV;
    constexpr int HUGE_D = 1 << 30;
    const int n = P.len, m = T.len;
    const int k_end = m - n;
    const int W = 1 << st.wlog2, mask = W - 1, wlog2 = st.wlog2;

    // ---- score 0 -------------------------------------------------------------------------------
    for (int i = lane; i < MS; i += 32) st.mmeta[i] = make_int2(1, 0);
    for (int i = lane; i < 4 * GS; i += 32) st.gmeta[i] = make_int2(1, 0);
    __syncwarp();
    int m0 = extend_coop(P, T, 0, 0, lane);
    lane_ext16 += (lane == 0) ? (unsigned long long)(m0 / 16 + 1) : 0ull;
    if (k_end == 0 && m0 >= m) { // identical sequences: nothing to trace
        out_penalty = 0;
        out_clipped = 0;
        wc.columns += (unsigned)m;
        if (WANT_CIGAR) {
            for (int i = lane; i < m; i += 32) cigar_out[i] = 'M';
            if (lane == 0) *cigar_cols = (uint32_t)m;
        }
        return WFA_OK;
    }
    if (lane == 0) {
        st.mmeta[0] = make_int2(0, 0);
        st.M[0] = (OffT)m0;
        st.step_lo[0] = (StepLoT)0;
        st.step_base[0] = (StepBaseT)0;
        st.prov[0] = 0;
    }
    unsigned prov_used = 1;
    int steps_wait = heur.steps;
    // wf-adaptive bookkeeping on M[0] (width 1 < min length: only the wait counter moves)
    if (heur.kind == 1) --steps_wait;
    __syncwarp();

    int s = 0;
    for (;;) {
        ++s;
        if (s >= st.s_cap) return WFA_OVERFLOW;
        // ---- inputs (A.2) -----------------------------------------------------------------------
        const InWf<OffT> mx = make_in<OffT>(st.M, st.mmeta, s - pn.x, MS - 1, wlog2);
        const InWf<OffT> mo1 = make_in<OffT>(st.M, st.mmeta, s - pn.o1 - pn.e1, MS - 1, wlog2);
        const InWf<OffT> mo2 = make_in<OffT>(st.M, st.mmeta, s - pn.o2 - pn.e2, MS - 1, wlog2);
        const size_t gstride = (size_t)GS << wlog2;
        const InWf<OffT> i1e = make_in<OffT>(st.G + CI1 * gstride, st.gmeta + CI1 * GS, s - pn.e1, GS - 1, wlog2);
        const InWf<OffT> i2e = make_in<OffT>(st.G + CI2 * gstride, st.gmeta + CI2 * GS, s - pn.e2, GS - 1, wlog2);
        const InWf<OffT> d1e = make_in<OffT>(st.G + CD1 * gstride, st.gmeta + CD1 * GS, s - pn.e1, GS - 1, wlog2);
        const InWf<OffT> d2e = make_in<OffT>(st.G + CD2 * gstride, st.gmeta + CD2 * GS, s - pn.e2, GS - 1, wlog2);
        const int mslot = s & (MS - 1), gslot = s & (GS - 1);
        if (mx.null() && mo1.null() && mo2.null() && i1e.null() && i2e.null() && d1e.null() && d2e.null()) {
            // null step: every output wavefront of score s is null
            __syncwarp();
            if (lane == 0) {
                st.mmeta[mslot] = make_int2(1, 0);
                st.gmeta[CI1 * GS + gslot] = make_int2(1, 0);
                st.gmeta[CI2 * GS + gslot] = make_int2(1, 0);
                st.gmeta[CD1 * GS + gslot] = make_int2(1, 0);
                st.gmeta[CD2 * GS + gslot] = make_int2(1, 0);
            }
            __syncwarp();
            continue;
        }
        int lo = INT32_MAX, hi = INT32_MIN;
#define TDB_LIM(w, dl, dh)                                      \
    if (!(w).null()) {                                          \
        lo = min(lo, (w).lo + (dl));                            \
        hi = max(hi, (w).lo + (int)(w).span1 - 1 + (dh));       \
    }
        TDB_LIM(mx, 0, 0)
        TDB_LIM(mo1, -1, 1)
        TDB_LIM(mo2, -1, 1)
        TDB_LIM(i1e, 1, 1)
        TDB_LIM(i2e, 1, 1)
        TDB_LIM(d1e, -1, -1)
        TDB_LIM(d2e, -1, -1)
#undef TDB_LIM
        lo = max(lo, -n - 1);
        hi = min(hi, m + 1);
        const int width = hi - lo + 1;
        if (width > W || prov_used + (unsigned)width > st.prov_cap) return WFA_OVERFLOW;
        const bool has_i1 = !mo1.null() || !i1e.null(), has_d1 = !mo1.null() || !d1e.null();
        const bool has_i2 = !mo2.null() || !i2e.null(), has_d2 = !mo2.null() || !d2e.null();
        OffT *Mo = st.M + ((size_t)mslot << wlog2);
        OffT *I1o = st.G + CI1 * gstride + ((size_t)gslot << wlog2);
        OffT *I2o = st.G + CI2 * gstride + ((size_t)gslot << wlog2);
        OffT *D1o = st.G + CD1 * gstride + ((size_t)gslot << wlog2);
        OffT *D2o = st.G + CD2 * gstride + ((size_t)gslot << wlog2);
        uint8_t *pv = st.prov + prov_used;
        wc.cells += (unsigned)width;

        int tl_m = INT32_MAX, th_m = INT32_MIN, tl_i1 = INT32_MAX, th_i1 = INT32_MIN, tl_i2 = INT32_MAX,
            th_i2 = INT32_MIN, tl_d1 = INT32_MAX, th_d1 = INT32_MIN, tl_d2 = INT32_MAX, th_d2 = INT32_MIN;
        int lane_min_d = HUGE_D;
        bool fin = false;
        for (int kb = lo; kb <= hi; kb += 32) {
            const int k = kb + lane;
            const bool act = k <= hi;
            int ins1 = NULLV, ins2 = NULLV, del1 = NULLV, del2 = NULLV, mv = NULLV;
            bool vm = false;
            if (act) {
                int b = 0, o, e;
                o = mo1.rd(k - 1, mask), e = i1e.rd(k - 1, mask);
                if (e >= o) ins1 = e, b |= PB_I1E; else ins1 = o;
                ins1 += 1;
                o = mo2.rd(k - 1, mask), e = i2e.rd(k - 1, mask);
                if (e >= o) ins2 = e, b |= PB_I2E; else ins2 = o;
                ins2 += 1;
                o = mo1.rd(k + 1, mask), e = d1e.rd(k + 1, mask);
                if (e >= o) del1 = e, b |= PB_D1E; else del1 = o;
                o = mo2.rd(k + 1, mask), e = d2e.rd(k + 1, mask);
                if (e >= o) del2 = e, b |= PB_D2E; else del2 = o;
                const int mis = mx.rd(k, mask) + 1;
                mv = max(max(max(ins1, ins2), max(del1, del2)), mis);
                int src = PM_NONE; // later tests override: priority X > D2 > D1 > I2 > I1
                if (mv == ins1) src = PM_I1;
                if (mv == ins2) src = PM_I2;
                if (mv == del1) src = PM_D1;
                if (mv == del2) src = PM_D2;
                if (mv == mis) src = PM_X;
                pv[k - lo] = (uint8_t)(b | src);
                if ((unsigned)mv > (unsigned)m || (unsigned)(mv - k) > (unsigned)n) mv = NULLV;
                vm = mv >= 0;
                if (vm) { // extend right away: trimming never changes offsets (A.1 step 1)
                    mv += extend_lane(P, T, mv - k, mv, lane_ext16);
                    lane_min_d = min(lane_min_d, max(n - (mv - k), m - mv));
                    fin |= (k == k_end) && (mv >= m);
                }
                Mo[k & mask] = (OffT)mv;
                if (has_i1) I1o[k & mask] = (OffT)ins1;
                if (has_i2) I2o[k & mask] = (OffT)ins2;
                if (has_d1) D1o[k & mask] = (OffT)del1;
                if (has_d2) D2o[k & mask] = (OffT)del2;
            }
            // trim_ends (A.2): first / last diagonal whose offset lies inside the DP matrix
            unsigned bal = __ballot_sync(TDB_FULL, vm);
            if (bal) tl_m = min(tl_m, kb + __ffs(bal) - 1), th_m = kb + 31 - __clz(bal);
#define TDB_TRIM(has, val, tl, th)                                                                        \
    if (has) {                                                                                            \
        bal = __ballot_sync(TDB_FULL, act && (unsigned)(val) <= (unsigned)m && (unsigned)((val)-k) <= (unsigned)n); \
        if (bal) tl = min(tl, kb + __ffs(bal) - 1), th = kb + 31 - __clz(bal);                            \
    }
            TDB_TRIM(has_i1, ins1, tl_i1, th_i1)
            TDB_TRIM(has_i2, ins2, tl_i2, th_i2)
            TDB_TRIM(has_d1, del1, tl_d1, th_d1)
            TDB_TRIM(has_d2, del2, tl_d2, th_d2)
#undef TDB_TRIM
        }
        if (lane == 0) {
            st.step_lo[s] = (StepLoT)lo;
            st.step_base[s] = (StepBaseT)prov_used;
        }
        prov_used += (unsigned)width;
        int2 mm = tl_m <= th_m ? make_int2(tl_m, th_m) : make_int2(1, 0);
        int2 g_i1 = tl_i1 <= th_i1 ? make_int2(tl_i1, th_i1) : make_int2(1, 0);
        int2 g_i2 = tl_i2 <= th_i2 ? make_int2(tl_i2, th_i2) : make_int2(1, 0);
        int2 g_d1 = tl_d1 <= th_d1 ? make_int2(tl_d1, th_d1) : make_int2(1, 0);
        int2 g_d2 = tl_d2 <= th_d2 ? make_int2(tl_d2, th_d2) : make_int2(1, 0);
        __syncwarp();
        // ---- termination (A.1 step 2) -----------------------------------------------------------
        if (__any_sync(TDB_FULL, fin)) {
            break;
        }
        // ---- wf-adaptive cut-off + equate (A.3) --------------------------------------------------
        if (heur.kind == 1 && mm.x <= mm.y) {
            --steps_wait;
            if (steps_wait <= 0 && (mm.y - mm.x + 1) >= heur.min_len) {
                const int min_d = min(max(n, m), __reduce_min_sync(TDB_FULL, lane_min_d));
                const int thr = heur.max_dist;
                const int top_limit = min(k_end, mm.y);
                int new_lo = mm.x;
                if (top_limit > mm.x) {
                    int first_ok = INT32_MAX;
                    for (int kb = mm.x; kb < top_limit && first_ok == INT32_MAX; kb += 32) {
                        const int k = kb + lane;
                        bool ok = false;
                        if (k < top_limit) {
                            const int off = (int)Mo[k & mask];
                            const int d = off >= 0 ? max(n - (off - k), m - off) : HUGE_D;
                            ok = d - min_d <= thr;
                        }
                        const unsigned bal = __ballot_sync(TDB_FULL, ok);
                        if (bal) first_ok = kb + __ffs(bal) - 1;
                    }
                    new_lo = min(first_ok, top_limit);
                }
                const int bot_limit = max(k_end, new_lo);
                int new_hi = mm.y;
                if (bot_limit < mm.y) {
                    int last_ok = INT32_MIN;
                    for (int kt = mm.y; kt > bot_limit && last_ok == INT32_MIN; kt -= 32) {
                        const int k = kt - lane;
                        bool ok = false;
                        if (k > bot_limit) {
                            const int off = (int)Mo[k & mask];
                            const int d = off >= 0 ? max(n - (off - k), m - off) : HUGE_D;
                            ok = d - min_d <= thr;
                        }
                        const unsigned bal = __ballot_sync(TDB_FULL, ok);
                        if (bal) last_ok = kt - (__ffs(bal) - 1);
                    }
                    new_hi = max(last_ok, bot_limit);
                }
                mm = make_int2(new_lo, new_hi);
                steps_wait = heur.steps;
            }
            // equate: gap wavefronts never extend beyond M (runs on every cut-off call)
            if (mm.x <= mm.y) {
                g_i1.x = max(g_i1.x, mm.x), g_i1.y = min(g_i1.y, mm.y);
                g_i2.x = max(g_i2.x, mm.x), g_i2.y = min(g_i2.y, mm.y);
                g_d1.x = max(g_d1.x, mm.x), g_d1.y = min(g_d1.y, mm.y);
                g_d2.x = max(g_d2.x, mm.x), g_d2.y = min(g_d2.y, mm.y);
            }
        }
        if (lane == 0) {
            st.mmeta[mslot] = mm;
            st.gmeta[CI1 * GS + gslot] = g_i1;
            st.gmeta[CI2 * GS + gslot] = g_i2;
            st.gmeta[CD1 * GS + gslot] = g_d1;
            st.gmeta[CD2 * GS + gslot] = g_d2;
        }
        __syncwarp();
    }

    // ---- backtrace (A.4): provenance chain from (M, s, k_end) back to (M, 0, 0) --------------------
    // Inside a gap the chain is predictable while it extends (score - e, diagonal -+ 1 per column), and every step of the
    // chain is a dependent load from the scratch (HBM / L2 for the global tiers: microseconds per column of a long
    // expansion).  So lane j reads the byte the chain reaches after j further extensions, one vote finds the first cell
    // that OPENED the gap instead, and up to 32 columns are consumed per round.  Lanes past that cell have read bytes that
    // are not on the path (any byte inside the provenance area: the index is clamped); they are not looked at.
    out_penalty = s;
    int nops = 0, n_del = 0;
    {
        int comp = -1 /* M */, cs = s, k = k_end;
        while (!(comp < 0 && cs == 0)) {
            if (comp < 0) {
                const int b = st.prov[(unsigned)st.step_base[cs] + (unsigned)(k - (int)st.step_lo[cs])];
                const int src = b & 7;
                int op;
                if (src == PM_X) op = OP_X, cs -= pn.x;
                else op = OP_CLOSE, comp = src - 1; // PM_I1..PM_D2 -> CI1..CD2
                if (nops >= st.ops_cap) return WFA_OVERFLOW;
                if (lane == 0) st.ops[nops] = (uint8_t)op;
                ++nops;
                continue;
            }
            const bool is_ins = comp == CI1 || comp == CI2, first = comp == CI1 || comp == CD1;
            const int e = first ? pn.e1 : pn.e2, o = first ? pn.o1 : pn.o2, dk = is_ins ? -1 : 1;
            const int bit = comp == CI1 ? PB_I1E : (comp == CI2 ? PB_I2E : (comp == CD1 ? PB_D1E : PB_D2E));
            const int cj = cs - lane * e, kj = k + lane * dk;
            bool ext = false;
            if (cj > 0) {
                const unsigned at = (unsigned)st.step_base[cj] + (unsigned)(kj - (int)st.step_lo[cj]);
                ext = (st.prov[min(at, prov_used - 1u)] & bit) != 0;
            }
            const unsigned opened = ~__ballot_sync(TDB_FULL, ext);
            const int cnt = opened ? __ffs(opened) : 32; // columns consumed: cnt - 1 extensions and, if any, the opening one
            if (nops + cnt > st.ops_cap) return WFA_OVERFLOW;
            if (lane < cnt) st.ops[nops + lane] = (uint8_t)(is_ins ? OP_I : OP_D);
            nops += cnt;
            if (!is_ins) n_del += cnt;
            k += cnt * dk;
            if (opened) cs -= (cnt - 1) * e + o + e, comp = -1;
            else cs -= 32 * e;
        }
    }
    __syncwarp();
    // ---- forward replay: matches re-extended only in the M component; clipped score on the fly ------
    // (a run of equal gap columns is found by one vote over the next 32 ops and emitted at once)
    ClipAcc acc;
    const int total_cols = m + n_del;
    acc.init(total_cols, clip);
    int v = 0, h = 0;
    bool in_m = true;
    for (int i = nops - 1;;) {
        if (in_m) {
            const int eq = extend_coop(P, T, v, h, lane);
            if (WANT_CIGAR)
                for (int j = lane; j < eq; j += 32) cigar_out[acc.col + j] = 'M';
            acc.emit('M', eq, pn);
            v += eq, h += eq;
        }
        if (i < 0) break;
        const int op = st.ops[i];
        if (op == OP_CLOSE) {
            in_m = true;
            --i;
            continue;
        }
        if (op == OP_X) {
            if (WANT_CIGAR && lane == 0) cigar_out[acc.col] = (uint8_t)'X';
            acc.emit('X', 1, pn);
            ++v, ++h, --i;
            continue;
        }
        const int j = i - lane;
        const unsigned differs = ~__ballot_sync(TDB_FULL, j >= 0 && st.ops[j] == op);
        const int run = differs ? __ffs(differs) - 1 : 32; // >= 1: lane 0 holds ops[i] itself
        const int ch = op == OP_I ? 'I' : 'D';
        if (WANT_CIGAR && lane < run) cigar_out[acc.col + lane] = (uint8_t)ch;
        acc.emit(ch, run, pn);
        if (op == OP_I) h += run;
        else v += run;
        in_m = false;
        i -= run;
    }
    out_clipped = acc.finish(pn);
    wc.columns += (unsigned)acc.col;
    if (WANT_CIGAR && lane == 0) *cigar_cols = (uint32_t)acc.col;
    __syncwarp();
    return WFA_OK;
}

} // namespace tdb
