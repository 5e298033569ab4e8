/*
 * tdo_wfa.c -- ORACLE (test infrastructure only, see tdo_oracle.h).
 *
 * Scalar restatement of the WFA2-lib gap-affine-2-piece end-to-end aligner in the configuration the
 * reference builds in src/commands/shared.rs:41-51 (x=8,o1=4,e1=2,o2=24,e2=1, scope=alignment,
 * memory=low, default heuristic wf-adaptive(10,50,1)) and calls in src/aligner.rs:248-260.
 * WFA2-lib is an un-vendored dependency (wfa2-sys 0.1.0, rev 4342b3b0, Cargo.toml:39); what follows
 * restates its published algorithm per SURVEY.md Appendix A:
 *   A.1 main loop      (wavefront_unialign / wavefront_extend_end2end)
 *   A.2 compute step   (wavefront_compute_affine2p_idm_piggyback + trim_ends)
 *   A.3 wf-adaptive    (wavefront_heuristic_wfadaptive + equate)
 *   A.4 backtrace      (piggy-back provenance; matches re-extended forward only in M state,
 *                       as WFA2's pcigar_unpack_affine does)
 * Parity pinned on src/aligner.rs tests (B4..B7, B9); heuristic-altered paths are unpinned.
 */
#include "tdo_oracle.h"

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OFF_NULL (INT32_MIN / 2)

/* provenance byte: bits0-2 = source of M, bit3..6 = I1,I2,D1,D2 came from "extend" (else "open") */
enum { PM_NONE = 0, PM_I1 = 1, PM_I2 = 2, PM_D1 = 3, PM_D2 = 4, PM_X = 5 };
#define PB_I1E 8
#define PB_I2E 16
#define PB_D1E 32
#define PB_D2E 64

/* backtrace op stream symbols */
enum { OP_X = 0, OP_I = 1, OP_D = 2, OP_CLOSE = 3 };

typedef struct {
    int lo, hi;   /* lo > hi or !exists => null wavefront */
    int exists;
    int32_t *off; /* off[k + koff] */
} wf_slot;

typedef struct {
    int lo, hi;   /* computed range (before trimming); lo > hi for a null step */
    size_t base;  /* provenance arena index of diagonal lo */
} step_rec;

enum { CM = 0, CI1 = 1, CI2 = 2, CD1 = 3, CD2 = 4, NCOMP = 5 };

struct tdo_aligner {
    tdo_penalties pen;
    tdo_heuristic heur;
    int scope;
    /* per-alignment geometry */
    int plen, tlen, koff, cap;
    uint8_t *pbuf, *tbuf; /* padded private copies */
    size_t seqcap;
    wf_slot *slots[NCOMP]; /* scope slots each */
    int32_t *offpool;
    size_t offpool_cap;
    uint8_t *prov;
    size_t prov_cap, prov_used;
    step_rec *steps;
    size_t steps_cap;
    uint8_t *ops;
    size_t ops_cap;
    char *cigar;
    size_t cigar_cap;
    int cigar_len;
    int score; /* -penalty */
    int status;
    tdo_stats st;
};

static void *xrealloc(void *p, size_t n) {
    void *q = realloc(p, n ? n : 1);
    if (!q) {
        fprintf(stderr, "tdo_oracle: out of memory (%zu bytes)\n", n);
        abort();
    }
    return q;
}

tdo_aligner *tdo_aligner_new(const tdo_penalties *pen, const tdo_heuristic *heur) {
    tdo_aligner *a = (tdo_aligner *)calloc(1, sizeof(*a));
    a->pen = *pen;
    if (heur)
        a->heur = *heur;
    else {
        a->heur.kind = 1, a->heur.min_wf_len = 10, a->heur.max_dist = 50, a->heur.steps = 1;
    }
    int sc = pen->x;
    if (pen->o1 + pen->e1 > sc) sc = pen->o1 + pen->e1;
    if (pen->o2 >= 0 && pen->o2 + pen->e2 > sc) sc = pen->o2 + pen->e2;
    a->scope = sc + 1;
    for (int c = 0; c < NCOMP; ++c) a->slots[c] = (wf_slot *)calloc((size_t)a->scope, sizeof(wf_slot));
    a->status = TDO_STATUS_UNATTAINABLE;
    return a;
}

void tdo_aligner_set_heuristic(tdo_aligner *a, const tdo_heuristic *heur) { a->heur = *heur; }

void tdo_aligner_delete(tdo_aligner *a) {
    if (!a) return;
    for (int c = 0; c < NCOMP; ++c) free(a->slots[c]);
    free(a->offpool);
    free(a->pbuf);
    free(a->tbuf);
    free(a->prov);
    free(a->steps);
    free(a->ops);
    free(a->cigar);
    free(a);
}

static void prepare(tdo_aligner *a, const uint8_t *p, int n, const uint8_t *t, int m) {
    a->plen = n;
    a->tlen = m;
    a->koff = n + 1;
    a->cap = n + m + 3;
    size_t need = (size_t)(n > m ? n : m) + 32;
    if (need > a->seqcap) {
        a->pbuf = (uint8_t *)xrealloc(a->pbuf, need);
        a->tbuf = (uint8_t *)xrealloc(a->tbuf, need);
        a->seqcap = need;
    }
    memcpy(a->pbuf, p, (size_t)n);
    memset(a->pbuf + n, '!', 16); /* WFA2 pads pattern/text with distinct end-of-sequence bytes */
    memcpy(a->tbuf, t, (size_t)m);
    memset(a->tbuf + m, '?', 16);
    size_t pool = (size_t)NCOMP * (size_t)a->scope * (size_t)a->cap;
    if (pool > a->offpool_cap) {
        a->offpool = (int32_t *)xrealloc(a->offpool, pool * sizeof(int32_t));
        a->offpool_cap = pool;
    }
    for (int c = 0; c < NCOMP; ++c)
        for (int s = 0; s < a->scope; ++s) {
            wf_slot *w = &a->slots[c][s];
            w->exists = 0;
            w->lo = 1;
            w->hi = 0;
            w->off = a->offpool + ((size_t)c * a->scope + s) * (size_t)a->cap;
        }
    a->prov_used = 0;
    memset(&a->st, 0, sizeof(a->st));
}

static inline int wf_is_null(const wf_slot *w) { return !w->exists || w->lo > w->hi; }
/* Reads outside a wavefront's [lo,hi] yield NULL. */
static inline int32_t wf_get(const wf_slot *w, int k, int koff) {
    return (k < w->lo || k > w->hi) ? OFF_NULL : w->off[k + koff];
}
static inline const wf_slot *in_slot(const tdo_aligner *a, int comp, int s) {
    static const wf_slot null_slot = {1, 0, 0, NULL};
    if (s < 0) return &null_slot;
    const wf_slot *w = &a->slots[comp][s % a->scope];
    return wf_is_null(w) ? &null_slot : w;
}

static void ensure_step(tdo_aligner *a, int s) {
    if ((size_t)s + 1 > a->steps_cap) {
        size_t nc = a->steps_cap ? a->steps_cap * 2 : 256;
        while (nc < (size_t)s + 1) nc *= 2;
        a->steps = (step_rec *)xrealloc(a->steps, nc * sizeof(step_rec));
        a->steps_cap = nc;
    }
}
static uint8_t *prov_alloc(tdo_aligner *a, size_t n) {
    if (a->prov_used + n > a->prov_cap) {
        size_t nc = a->prov_cap ? a->prov_cap * 2 : 4096;
        while (nc < a->prov_used + n) nc *= 2;
        a->prov = (uint8_t *)xrealloc(a->prov, nc);
        a->prov_cap = nc;
    }
    uint8_t *r = a->prov + a->prov_used;
    a->prov_used += n;
    return r;
}

/* wavefront_compute_trim_ends: drop leading/trailing diagonals whose offset is out of the DP matrix. */
static void trim_ends(const tdo_aligner *a, wf_slot *w) {
    const uint32_t n = (uint32_t)a->plen, m = (uint32_t)a->tlen;
    int k;
    for (k = w->hi; k >= w->lo; --k) {
        int32_t off = w->off[k + a->koff];
        uint32_t h = (uint32_t)off, v = (uint32_t)(off - k);
        if (h <= m && v <= n) break;
    }
    w->hi = k;
    for (k = w->lo; k <= w->hi; ++k) {
        int32_t off = w->off[k + a->koff];
        uint32_t h = (uint32_t)off, v = (uint32_t)(off - k);
        if (h <= m && v <= n) break;
    }
    w->lo = k;
}

/* A.2: compute wavefronts of score s.  Returns 0 for a null step. */
static int compute_step(tdo_aligner *a, int s) {
    const tdo_penalties *pn = &a->pen;
    const int two = pn->o2 >= 0;
    const wf_slot *mx = in_slot(a, CM, s - pn->x);
    const wf_slot *mo1 = in_slot(a, CM, s - pn->o1 - pn->e1);
    const wf_slot *mo2 = two ? in_slot(a, CM, s - pn->o2 - pn->e2) : in_slot(a, CM, -1);
    const wf_slot *i1e = in_slot(a, CI1, s - pn->e1);
    const wf_slot *d1e = in_slot(a, CD1, s - pn->e1);
    const wf_slot *i2e = two ? in_slot(a, CI2, s - pn->e2) : in_slot(a, CM, -1);
    const wf_slot *d2e = two ? in_slot(a, CD2, s - pn->e2) : in_slot(a, CM, -1);
    const int sm = s % a->scope;
    wf_slot *om = &a->slots[CM][sm], *oi1 = &a->slots[CI1][sm], *oi2 = &a->slots[CI2][sm];
    wf_slot *od1 = &a->slots[CD1][sm], *od2 = &a->slots[CD2][sm];
    ensure_step(a, s);
    if (!mx->exists && !mo1->exists && !mo2->exists && !i1e->exists && !d1e->exists && !i2e->exists &&
        !d2e->exists) {
        om->exists = oi1->exists = oi2->exists = od1->exists = od2->exists = 0;
        a->steps[s].lo = 1;
        a->steps[s].hi = 0;
        a->steps[s].base = 0;
        return 0;
    }
    /* wavefront_compute_limits_input, restricted to non-null inputs (null ones only add all-NULL
     * diagonals that trim_ends removes again). */
    int lo = INT_MAX, hi = INT_MIN;
#define LIM(w, dl, dh)                           \
    if ((w)->exists) {                           \
        if ((w)->lo + (dl) < lo) lo = (w)->lo + (dl); \
        if ((w)->hi + (dh) > hi) hi = (w)->hi + (dh); \
    }
    LIM(mx, 0, 0)
    LIM(mo1, -1, +1)
    LIM(mo2, -1, +1)
    LIM(i1e, +1, +1)
    LIM(i2e, +1, +1)
    LIM(d1e, -1, -1)
    LIM(d2e, -1, -1)
#undef LIM
    /* keep inside the allocation (diagonals outside [-n-1, m+1] can only hold invalid offsets) */
    if (lo < -a->plen - 1) lo = -a->plen - 1;
    if (hi > a->tlen + 1) hi = a->tlen + 1;
    if (lo > hi) { /* cannot happen for finite inputs; treat as null step */
        om->exists = oi1->exists = oi2->exists = od1->exists = od2->exists = 0;
        a->steps[s].lo = 1, a->steps[s].hi = 0, a->steps[s].base = 0;
        return 0;
    }
    const int koff = a->koff;
    const uint32_t n = (uint32_t)a->plen, m = (uint32_t)a->tlen;
    uint8_t *pv = prov_alloc(a, (size_t)(hi - lo + 1));
    a->steps[s].lo = lo;
    a->steps[s].hi = hi;
    a->steps[s].base = (size_t)(pv - a->prov);
    a->st.cells += hi - lo + 1;
    const int has_i1 = mo1->exists || i1e->exists, has_d1 = mo1->exists || d1e->exists;
    const int has_i2 = mo2->exists || i2e->exists, has_d2 = mo2->exists || d2e->exists;
    for (int k = lo; k <= hi; ++k) {
        uint8_t b = 0;
        /* gap components: extend preferred over open on ties */
        int32_t o, e, ins1, ins2, del1, del2;
        o = wf_get(mo1, k - 1, koff), e = wf_get(i1e, k - 1, koff);
        if (e >= o) { ins1 = e; b |= PB_I1E; } else ins1 = o;
        ins1 += 1;
        o = wf_get(mo2, k - 1, koff), e = wf_get(i2e, k - 1, koff);
        if (e >= o) { ins2 = e; b |= PB_I2E; } else ins2 = o;
        ins2 += 1;
        o = wf_get(mo1, k + 1, koff), e = wf_get(d1e, k + 1, koff);
        if (e >= o) { del1 = e; b |= PB_D1E; } else del1 = o;
        o = wf_get(mo2, k + 1, koff), e = wf_get(d2e, k + 1, koff);
        if (e >= o) { del2 = e; b |= PB_D2E; } else del2 = o;
        int32_t mis = wf_get(mx, k, koff) + 1;
        int32_t mxv = ins1;
        if (ins2 > mxv) mxv = ins2;
        if (del1 > mxv) mxv = del1;
        if (del2 > mxv) mxv = del2;
        if (mis > mxv) mxv = mis;
        /* later tests override earlier ones: priority X > D2 > D1 > I2 > I1 */
        int src = PM_NONE;
        if (mxv == ins1) src = PM_I1;
        if (mxv == ins2) src = PM_I2;
        if (mxv == del1) src = PM_D1;
        if (mxv == del2) src = PM_D2;
        if (mxv == mis) src = PM_X;
        uint32_t h = (uint32_t)mxv, v = (uint32_t)(mxv - k);
        if (h > m || v > n) mxv = OFF_NULL;
        pv[k - lo] = (uint8_t)(b | src);
        om->off[k + koff] = mxv;
        oi1->off[k + koff] = ins1;
        oi2->off[k + koff] = ins2;
        od1->off[k + koff] = del1;
        od2->off[k + koff] = del2;
    }
    om->exists = 1, om->lo = lo, om->hi = hi;
    trim_ends(a, om);
#define OUT(w, has)                    \
    (w)->exists = (has);               \
    (w)->lo = lo, (w)->hi = hi;        \
    if (has) trim_ends(a, (w));
    OUT(oi1, has_i1)
    OUT(od1, has_d1)
    OUT(oi2, has_i2)
    OUT(od2, has_d2)
#undef OUT
    return 1;
}

/* wavefront_extend_matches_packed_kernel: 8-byte blocks, then clamp to the sequence ends. */
static inline int extend_one(const tdo_aligner *a, int v, int h) {
    const uint8_t *p = a->pbuf + v, *t = a->tbuf + h;
    int room = a->plen - v < a->tlen - h ? a->plen - v : a->tlen - h;
    int eq = 0;
    while (eq < room) {
        uint64_t x, y;
        memcpy(&x, p + eq, 8);
        memcpy(&y, t + eq, 8);
        uint64_t c = x ^ y;
        if (c) {
            eq += __builtin_ctzll(c) >> 3;
            break;
        }
        eq += 8;
    }
    return eq < room ? eq : room;
}

static void extend_wavefront(tdo_aligner *a, wf_slot *w) {
    for (int k = w->lo; k <= w->hi; ++k) {
        int32_t off = w->off[k + a->koff];
        if (off < 0) continue; /* NULL */
        int eq = extend_one(a, off - k, off);
        w->off[k + a->koff] = off + eq;
        a->st.ext16 += eq / 16 + 1;
    }
}

/* A.3 wf-adaptive cut-off + equate of the gap wavefronts. */
static void heuristic_cutoff(tdo_aligner *a, int s, int *steps_wait) {
    const int sm = s % a->scope;
    wf_slot *w = &a->slots[CM][sm];
    if (wf_is_null(w)) return;
    --(*steps_wait);
    if (*steps_wait <= 0 && (w->hi - w->lo + 1) >= a->heur.min_wf_len) {
        const int n = a->plen, m = a->tlen, koff = a->koff;
        int min_d = n > m ? n : m;
        for (int k = w->lo; k <= w->hi; ++k) {
            int32_t off = w->off[k + koff];
            int d = off >= 0 ? ((n - (off - k)) > (m - off) ? (n - (off - k)) : (m - off)) : -OFF_NULL;
            if (d < min_d) min_d = d;
        }
#define DIST(k) \
    (w->off[(k) + koff] >= 0 ? ((n - (w->off[(k) + koff] - (k))) > (m - w->off[(k) + koff]) ? (n - (w->off[(k) + koff] - (k))) : (m - w->off[(k) + koff])) : -OFF_NULL)
        const int k_end = m - n, thr = a->heur.max_dist;
        int top_limit = k_end < w->hi ? k_end : w->hi;
        int lo_red = w->lo;
        for (int k = w->lo; k < top_limit; ++k) {
            if (DIST(k) - min_d <= thr) break;
            ++lo_red;
        }
        w->lo = lo_red;
        int bot_limit = k_end > w->lo ? k_end : w->lo;
        int hi_red = w->hi;
        for (int k = w->hi; k > bot_limit; --k) {
            if (DIST(k) - min_d <= thr) break;
            --hi_red;
        }
        w->hi = hi_red;
#undef DIST
        *steps_wait = a->heur.steps;
    }
    if (w->lo > w->hi) return;
    for (int c = CI1; c <= CD2; ++c) {
        wf_slot *g = &a->slots[c][sm];
        if (!g->exists) continue;
        if (w->lo > g->lo) g->lo = w->lo;
        if (w->hi < g->hi) g->hi = w->hi;
    }
}

static void ops_push(tdo_aligner *a, size_t *n, uint8_t op) {
    if (*n + 1 > a->ops_cap) {
        a->ops_cap = a->ops_cap ? a->ops_cap * 2 : 1024;
        a->ops = (uint8_t *)xrealloc(a->ops, a->ops_cap);
    }
    a->ops[(*n)++] = op;
}

/* A.4: follow provenance back from (M, S, k_end); then replay forward, re-extending matches only
 * while in the M component (WFA2 pcigar_unpack_affine). */
static int backtrace(tdo_aligner *a, int S) {
    const tdo_penalties *pn = &a->pen;
    size_t nops = 0;
    int comp = CM, s = S, k = a->tlen - a->plen;
    while (!(comp == CM && s == 0)) {
        if (s < 0 || a->steps[s].lo > a->steps[s].hi || k < a->steps[s].lo || k > a->steps[s].hi) return -1;
        uint8_t b = a->prov[a->steps[s].base + (size_t)(k - a->steps[s].lo)];
        switch (comp) {
        case CM:
            switch (b & 7) {
            case PM_X: ops_push(a, &nops, OP_X); s -= pn->x; break;
            case PM_I1: ops_push(a, &nops, OP_CLOSE); comp = CI1; break;
            case PM_I2: ops_push(a, &nops, OP_CLOSE); comp = CI2; break;
            case PM_D1: ops_push(a, &nops, OP_CLOSE); comp = CD1; break;
            case PM_D2: ops_push(a, &nops, OP_CLOSE); comp = CD2; break;
            default: return -1;
            }
            break;
        case CI1:
            ops_push(a, &nops, OP_I);
            if (b & PB_I1E) s -= pn->e1; else { s -= pn->o1 + pn->e1; comp = CM; }
            k -= 1;
            break;
        case CI2:
            ops_push(a, &nops, OP_I);
            if (b & PB_I2E) s -= pn->e2; else { s -= pn->o2 + pn->e2; comp = CM; }
            k -= 1;
            break;
        case CD1:
            ops_push(a, &nops, OP_D);
            if (b & PB_D1E) s -= pn->e1; else { s -= pn->o1 + pn->e1; comp = CM; }
            k += 1;
            break;
        case CD2:
            ops_push(a, &nops, OP_D);
            if (b & PB_D2E) s -= pn->e2; else { s -= pn->o2 + pn->e2; comp = CM; }
            k += 1;
            break;
        }
    }
    if (k != 0) return -1;
    size_t need = (size_t)a->plen + (size_t)a->tlen + 1;
    if (need > a->cigar_cap) {
        a->cigar = (char *)xrealloc(a->cigar, need);
        a->cigar_cap = need;
    }
    int v = 0, h = 0, len = 0, in_m = 1;
    for (size_t i = nops; i-- > 0;) {
        if (in_m) {
            int eq = extend_one(a, v, h);
            memset(a->cigar + len, 'M', (size_t)eq);
            len += eq, v += eq, h += eq;
        }
        switch (a->ops[i]) {
        case OP_X: a->cigar[len++] = 'X'; ++v, ++h; break;
        case OP_I: a->cigar[len++] = 'I'; ++h; in_m = 0; break;
        case OP_D: a->cigar[len++] = 'D'; ++v; in_m = 0; break;
        case OP_CLOSE: in_m = 1; break;
        }
        if (v > a->plen || h > a->tlen) return -1;
    }
    if (!in_m) return -1;
    int eq = extend_one(a, v, h);
    memset(a->cigar + len, 'M', (size_t)eq);
    len += eq, v += eq, h += eq;
    if (v != a->plen || h != a->tlen) return -1;
    a->cigar_len = len;
    a->st.columns = len;
    return 0;
}

int tdo_align_end_to_end(tdo_aligner *a, const uint8_t *pattern, int plen, const uint8_t *text, int tlen) {
    prepare(a, pattern, plen, text, tlen);
    a->cigar_len = 0;
    a->score = INT32_MIN;
    const int k_end = tlen - plen;
    wf_slot *w0 = &a->slots[CM][0];
    w0->exists = 1, w0->lo = w0->hi = 0, w0->off[a->koff] = 0;
    ensure_step(a, 0);
    uint8_t *p0 = prov_alloc(a, 1);
    p0[0] = 0;
    a->steps[0].lo = a->steps[0].hi = 0, a->steps[0].base = 0;
    int steps_wait = a->heur.steps, null_steps = 0, s = 0;
    for (;;) {
        wf_slot *w = &a->slots[CM][s % a->scope];
        if (!w->exists) {
            if (null_steps > a->scope) return a->status = TDO_STATUS_UNATTAINABLE;
        } else if (w->lo <= w->hi) {
            extend_wavefront(a, w);
            if (w->hi - w->lo + 1 > a->st.max_width) a->st.max_width = w->hi - w->lo + 1;
            if (w->lo <= k_end && k_end <= w->hi && w->off[k_end + a->koff] >= tlen) break;
            if (a->heur.kind == 1) heuristic_cutoff(a, s, &steps_wait);
        }
        ++s;
        if (compute_step(a, s)) null_steps = 0; else ++null_steps;
        if (s > 8 * (plen + tlen) * (a->pen.x + 1) + 4096) return a->status = TDO_STATUS_MAX_STEPS_REACHED;
    }
    a->st.steps = s;
    a->score = -s;
    if (backtrace(a, s) != 0) return a->status = TDO_STATUS_OOM; /* internal inconsistency */
    return a->status = TDO_STATUS_ALG_COMPLETED;
}

int tdo_score(const tdo_aligner *a) { return a->score; }

const char *tdo_cigar_ops(const tdo_aligner *a, int *len) {
    *len = a->cigar_len;
    return a->cigar;
}

void tdo_get_stats(const tdo_aligner *a, tdo_stats *st) { *st = a->st; }

/* src/aligner.rs:303-318 */
static int op_score(char op, int len, const tdo_penalties *pn) {
    switch (op) {
    case 'M': return 0;
    case 'X': return len * pn->x;
    default: {
        int s1 = pn->o1 + pn->e1 * len;
        if (pn->o2 < 0) return s1;
        int s2 = pn->o2 + pn->e2 * len;
        return s1 < s2 ? s1 : s2;
    }
    }
}

/* src/aligner.rs:320-357 */
int tdo_cigar_score_clipped(const tdo_aligner *a, int clip) {
    long begin = (long)clip, end = (long)a->cigar_len - (long)clip;
    int score = 0, oplen = 0;
    char last = 0;
    for (long i = begin; i < end; ++i) {
        char cur = a->cigar[i];
        if (last && cur != last) {
            score -= op_score(last, oplen, &a->pen);
            oplen = 0;
        }
        last = cur;
        ++oplen;
    }
    if (last) score -= op_score(last, oplen, &a->pen);
    return score;
}

/* src/aligner.rs:404-436 */
int tdo_cigar_string(const tdo_aligner *a, int clip, char *buf, int buflen) {
    int begin = clip, end = a->cigar_len - clip, w = 0;
    if (buflen > 0) buf[0] = 0;
    if (begin >= end) return 0;
    char last = a->cigar[begin];
    int run = 1;
    for (int i = begin + 1; i <= end; ++i) {
        if (i < end && a->cigar[i] == last) {
            ++run;
            continue;
        }
        int r = snprintf(buf + w, (size_t)(buflen - w), "%d%c", run, last);
        if (r < 0 || r >= buflen - w) return -1;
        w += r;
        if (i < end) last = a->cigar[i], run = 1;
    }
    return w;
}

/* Independent optimal-score check: Gotoh with two affine gap pieces, O(n*m) time, O(m) space. */
int64_t tdo_gotoh2p_penalty(const uint8_t *p, int n, const uint8_t *t, int m, const tdo_penalties *pn) {
    const int64_t INF = (int64_t)1 << 50;
    const int two = pn->o2 >= 0;
    size_t W = (size_t)m + 1;
    int64_t *M = (int64_t *)malloc(5 * W * sizeof(int64_t));
    int64_t *D1 = M + W, *D2 = D1 + W, *I1 = D2 + W, *I2 = I1 + W; /* D*: gap consuming pattern (vertical) */
    /* row 0: only insertions (consume text) */
    M[0] = 0;
    D1[0] = D2[0] = INF;
    for (int j = 1; j <= m; ++j) {
        int64_t g1 = pn->o1 + (int64_t)pn->e1 * j, g2 = two ? pn->o2 + (int64_t)pn->e2 * j : INF;
        I1[j] = g1, I2[j] = g2;
        M[j] = g1 < g2 ? g1 : g2;
        D1[j] = D2[j] = INF;
    }
    I1[0] = I2[0] = INF;
    for (int i = 1; i <= n; ++i) {
        int64_t diag = M[0]; /* M[i-1][0] */
        int64_t g1 = pn->o1 + (int64_t)pn->e1 * i, g2 = two ? pn->o2 + (int64_t)pn->e2 * i : INF;
        D1[0] = g1, D2[0] = g2;
        M[0] = g1 < g2 ? g1 : g2;
        int64_t i1 = INF, i2 = INF; /* horizontal state along this row */
        for (int j = 1; j <= m; ++j) {
            int64_t up = M[j]; /* M[i-1][j] */
            int64_t d1 = D1[j] + pn->e1, od1 = up + pn->o1 + pn->e1;
            if (od1 < d1) d1 = od1;
            int64_t d2 = INF;
            if (two) {
                d2 = D2[j] + pn->e2;
                int64_t od2 = up + pn->o2 + pn->e2;
                if (od2 < d2) d2 = od2;
            }
            int64_t left = M[j - 1];
            int64_t ni1 = i1 + pn->e1, oi1 = left + pn->o1 + pn->e1;
            if (oi1 < ni1) ni1 = oi1;
            int64_t ni2 = INF;
            if (two) {
                ni2 = i2 + pn->e2;
                int64_t oi2 = left + pn->o2 + pn->e2;
                if (oi2 < ni2) ni2 = oi2;
            }
            i1 = ni1, i2 = ni2;
            int64_t best = diag + (p[i - 1] == t[j - 1] ? 0 : pn->x);
            if (d1 < best) best = d1;
            if (d2 < best) best = d2;
            if (i1 < best) best = i1;
            if (i2 < best) best = i2;
            diag = up;
            M[j] = best, D1[j] = d1, D2[j] = d2;
        }
        (void)I1, (void)I2;
    }
    int64_t r = M[m];
    free(M);
    return r;
}
