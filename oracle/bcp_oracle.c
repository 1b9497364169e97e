/*
 * bcp_oracle.c -- TEST INFRASTRUCTURE (see bcp_oracle.h). Plain-C restatement of the reference's
 * BCP path. One global mutable solver state, two watched literals per long clause, binary clauses
 * held only in watch rows, LIFO propagation stack -- as the reference does it; written from the
 * reference's behaviour, not from its text. Each function cites the reference lines it follows.
 *
 * Deliberate simplifications, none of which can change a fixpoint, a conflict flag or the level-0 set:
 *  - initial watch choice is "first two literals" (the reference ranks candidates with WLIsLitBetter,
 *    TopiWL.cc:606-649; layout/order only, SURVEY.md section 2 "Params");
 *  - the fresh-variable-unit aliasing trick (Topi.cc:559-572) is restated as a plain level-0
 *    assignment of a new variable (same external observable: the variable is true at level 0);
 *  - ProcessDelayedImplication (TopiBcp.cc:413-744) is unreachable when BCP runs to fixpoint after
 *    every decision from a level-0 fixpoint and is omitted;
 *  - on a conflict the propagation of that BCP call stops (NewContradiction's same-level branch,
 *    TopiBcp.cc:141-158, erases every current-level literal from the stack, so the reference stops too).
 */
#include "bcp_oracle.h"
#include "bcp_io.h"

#include <stdlib.h>
#include <string.h>

typedef uint32_t ulit; /* internal literal = 2*var + neg, vars from 1 (Topi.hpp:3057-3076) */

typedef struct { uint32_t *p; uint32_t n, cap; } u32vec;

static void u32vec_push(u32vec *v, uint32_t x)
{
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 4; v->p = (uint32_t *)realloc(v->p, (size_t)v->cap * 4); }
    v->p[v->n++] = x;
}

typedef struct {
    u32vec bins;   /* other literal of each binary clause containing this literal (Topi.hpp:1952-1962) */
    u32vec longs;  /* pairs (cached literal, clause index) */
    uint8_t ever;  /* row was ever allocated: the reference's !wi.IsEmpty() (Topi.hpp:1939) */
} wrow;

struct orc {
    /* ext <-> int (Topi.hpp:413-425, Topi.cc:285-355) */
    uint32_t *e2i; uint64_t e2i_cap;
    int32_t *i2e; uint64_t i2e_cap;
    uint32_t nvars;
    int32_t max_user_var;
    /* assignment (Topi.hpp:2140-2200): 0 unassigned, 1 var true, 2 var false */
    uint8_t *val; uint32_t *level; uint64_t var_cap;
    wrow *rows; uint64_t rows_cap;
    u32vec arena;           /* [size][lits...]; index 0 is the bad clause index (Topi.hpp:3092) */
    u32vec trail, toprop;
    uint32_t dec_level;
    uint32_t trail_l0;      /* number of trail entries at level 0 */
    int64_t *stamp; uint64_t stamp_cap; int64_t counter; /* CHandleNewCls (Topi.hpp:3110-3185) */
    u32vec cur;
    int contradictory;
    uint64_t n_impl, n_asg, n_bcp;
};

static inline ulit neg(ulit l) { return l ^ 1u; }
static inline uint32_t var_of(ulit l) { return l >> 1; }
static inline int is_true(const orc *o, ulit l) { uint8_t v = o->val[var_of(l)]; return v && ((v == 2) == (int)(l & 1)); }
static inline int is_false(const orc *o, ulit l) { uint8_t v = o->val[var_of(l)]; return v && ((v == 2) != (int)(l & 1)); }
static inline int is_assigned(const orc *o, ulit l) { return o->val[var_of(l)] != 0; }

orc *orc_create(void)
{
    orc *o = (orc *)calloc(1, sizeof(orc));
    u32vec_push(&o->arena, 0);
    return o;
}

void orc_release(orc *o)
{
    if (!o) return;
    for (uint64_t i = 0; i < o->rows_cap; ++i) { free(o->rows[i].bins.p); free(o->rows[i].longs.p); }
    free(o->rows); free(o->e2i); free(o->i2e); free(o->val); free(o->level); free(o->arena.p);
    free(o->trail.p); free(o->toprop.p); free(o->stamp); free(o->cur.p);
    free(o);
}

static void grow_vars(orc *o, uint64_t nv)
{
    if (nv + 1 > o->var_cap) {
        uint64_t nc = o->var_cap ? o->var_cap : 64;
        while (nc < nv + 1) nc *= 2;
        o->val = (uint8_t *)realloc(o->val, nc); memset(o->val + o->var_cap, 0, nc - o->var_cap);
        o->level = (uint32_t *)realloc(o->level, nc * 4); memset(o->level + o->var_cap, 0, (nc - o->var_cap) * 4);
        o->i2e = (int32_t *)realloc(o->i2e, nc * 4); memset(o->i2e + o->var_cap, 0, (nc - o->var_cap) * 4);
        o->stamp = (int64_t *)realloc(o->stamp, nc * 8); memset(o->stamp + o->var_cap, 0, (nc - o->var_cap) * 8);
        o->var_cap = nc;
    }
    if (2 * nv + 2 > o->rows_cap) {
        uint64_t nc = o->rows_cap ? o->rows_cap : 128;
        while (nc < 2 * nv + 2) nc *= 2;
        o->rows = (wrow *)realloc(o->rows, nc * sizeof(wrow)); memset(o->rows + o->rows_cap, 0, (nc - o->rows_cap) * sizeof(wrow));
        o->rows_cap = nc;
    }
}

/* HandleIncomingUserVar (Topi.cc:285-355): vars are numbered in order of first appearance */
static void incoming_var(orc *o, int32_t v)
{
    if ((uint64_t)v >= o->e2i_cap) {
        uint64_t nc = o->e2i_cap ? o->e2i_cap : 64;
        while (nc <= (uint64_t)v) nc *= 2;
        o->e2i = (uint32_t *)realloc(o->e2i, nc * 4); memset(o->e2i + o->e2i_cap, 0, (nc - o->e2i_cap) * 4);
        o->e2i_cap = nc;
    }
    if (o->e2i[v] == 0) {
        ++o->nvars;
        grow_vars(o, o->nvars);
        o->e2i[v] = 2u * o->nvars;
        o->i2e[o->nvars] = v;
    }
    if (v > o->max_user_var) o->max_user_var = v;
}

/* Assign (TopiAsg.cc:46-139). Returns 1 if l is already false (contradiction), else 0 */
static int assign(orc *o, ulit l, uint32_t level)
{
    ++o->n_asg;
    if (is_assigned(o, l)) return is_false(o, l);
    o->val[var_of(l)] = (l & 1) ? 2 : 1;
    o->level[var_of(l)] = level;
    u32vec_push(&o->trail, l);
    if (level == 0) o->trail_l0 = o->trail.n;
    u32vec_push(&o->toprop, l);
    return 0;
}

/* Backtrack (TopiBacktrack.cc:18-53) to level 0: the only target the probing scenario needs */
static void backtrack0(orc *o)
{
    while (o->trail.n > o->trail_l0) o->val[var_of(o->trail.p[--o->trail.n])] = 0;
    o->toprop.n = 0;
    o->dec_level = 0;
}

/* WLBinaryWatchExists (TopiWL.cc:15-26) */
static int bin_exists(const orc *o, ulit a, ulit b)
{
    const u32vec *r = &o->rows[a].bins;
    for (uint32_t i = 0; i < r->n; ++i) if (r->p[i] == b) return 1;
    return 0;
}

static void add_long_watch(orc *o, ulit l, ulit cached, uint32_t cls)
{
    o->rows[l].ever = 1;
    u32vec_push(&o->rows[l].longs, cached);
    u32vec_push(&o->rows[l].longs, cls);
}

/* AddClsToBufferAndWatch (TopiConflictAnalysis.cc:14-113) for an initial clause of size >= 2 */
static uint32_t add_to_buffer_and_watch(orc *o, const ulit *c, uint32_t n)
{
    if (n == 2) {
        /* default /bcp/existing_bin_wl_start = 1: a duplicate binary is not added again (25-41) */
        if (!bin_exists(o, c[0], c[1])) {
            o->rows[c[0]].ever = o->rows[c[1]].ever = 1;
            u32vec_push(&o->rows[c[0]].bins, c[1]);
            u32vec_push(&o->rows[c[1]].bins, c[0]);
        }
        return 0;
    }
    uint32_t start = o->arena.n;
    u32vec_push(&o->arena, n);
    for (uint32_t i = 0; i < n; ++i) u32vec_push(&o->arena, c[i]);
    add_long_watch(o, c[0], c[1], start);
    add_long_watch(o, c[1], c[0], start);
    return start;
}

/* AddUserClause (Topi.cc:377-653) */
void orc_add_clause(orc *o, const int32_t *lits, size_t n)
{
    if (o->contradictory) return;
    if (n == 0 || lits[0] == 0) { o->contradictory = 1; return; } /* 451-456 */
    for (size_t i = 0; i < n && lits[i] != 0; ++i) { /* the probe universe spans every variable named, kept or not */
        const int32_t a = lits[i] < 0 ? -lits[i] : lits[i];
        if (a > o->max_user_var) o->max_user_var = a;
    }
    const uint32_t nvars_start = o->nvars;
    ++o->counter;
    o->cur.n = 0;
    int tautology = 0;
    for (size_t i = 0; i < n && lits[i] != 0; ++i) {
        const int32_t l = lits[i];
        const int32_t v = l < 0 ? -l : l;
        incoming_var(o, v);
        const ulit il = l > 0 ? o->e2i[v] : neg(o->e2i[v]);
        if (is_false(o, il) && o->level[var_of(il)] == 0) continue; /* 487-491 */
        const int64_t st = o->stamp[var_of(il)];
        const int64_t me = (il & 1) ? -o->counter : o->counter;
        if (st == me) continue;                      /* duplicate literal, 506-510 */
        if (st == -me) { tautology = 1; break; }     /* tautology, 501-504 */
        o->stamp[var_of(il)] = me;
        u32vec_push(&o->cur, il);
    }
    if (tautology) {
        /* roll back variables first seen in this clause (434-443) */
        for (uint32_t iv = nvars_start + 1; iv <= o->nvars; ++iv) { o->e2i[o->i2e[iv]] = 0; o->i2e[iv] = 0; o->stamp[iv] = 0; }
        o->nvars = nvars_start;
        return;
    }
    if (o->cur.n == 0) { o->contradictory = 1; return; } /* 552-557 */
    if (o->cur.n == 1) {                                  /* 590-619 */
        if (!is_true(o, o->cur.p[0])) assign(o, o->cur.p[0], 0);
        return;
    }
    add_to_buffer_and_watch(o, o->cur.p, o->cur.n);
    /* 630-645 cannot fire: level-0-false literals were dropped above and nothing else is assigned */
}

void orc_add_words(orc *o, const int32_t *w, uint64_t nwords)
{
    uint64_t s = 0;
    for (uint64_t i = 0; i < nwords; ++i) {
        if (w[i] == 0) { orc_add_clause(o, w + s, (size_t)(i - s)); s = i + 1; }
        else { int32_t a = w[i] < 0 ? -w[i] : w[i]; if (a > o->max_user_var) o->max_user_var = a; }
    }
}

/* BCP (TopiBcp.cc:103-410). Returns 1 on contradiction */
static int bcp(orc *o)
{
    ++o->n_bcp;
    int conflict = 0;
    while (o->toprop.n && !conflict) {
        const ulit p = o->toprop.p[--o->toprop.n];        /* LIFO, 183 */
        const ulit np = neg(p);
        wrow *wi = &o->rows[np];
        if (!wi->ever) continue;                         /* 193-196 */
        ++o->n_impl;                                     /* 198 */
        const uint32_t lvl = o->level[var_of(p)];
        /* binary watches first, 207-246 */
        for (uint32_t i = 0; i < wi->bins.n; ++i) {
            const ulit other = wi->bins.p[i];
            if (!is_assigned(o, other)) assign(o, other, lvl);
            else if (is_false(o, other)) { conflict = 1; break; }
        }
        /* long watches, 249-375 */
        for (uint32_t i = 0; !conflict && i < wi->longs.n; i += 2) {
            if (is_true(o, wi->longs.p[i])) continue;    /* cached literal true, 255-259 */
            const uint32_t ci = wi->longs.p[i + 1];
            uint32_t *c = o->arena.p + ci + 1;
            const uint32_t sz = o->arena.p[ci];
            if (c[1] == np) { c[1] = c[0]; c[0] = np; }   /* 270 */
            const ulit other = c[1];
            wi->longs.p[i] = other;                      /* 279 */
            if (is_true(o, other)) continue;             /* 282-291 */
            uint32_t k = 2;
            while (k < sz && is_false(o, c[k])) ++k;     /* FindBestWLCand, 43-53 */
            if (k < sz) {
                /* SwapCurrWatch (27-40): swap-remove this watch, re-watch on c[k] */
                const ulit nw = c[k];
                c[k] = c[0]; c[0] = nw;
                wi->longs.p[i] = wi->longs.p[wi->longs.n - 2];
                wi->longs.p[i + 1] = wi->longs.p[wi->longs.n - 1];
                wi->longs.n -= 2;
                i -= 2;
                add_long_watch(o, nw, c[1], ci);
                wi = &o->rows[np];
                continue;
            }
            if (is_false(o, other)) conflict = 1;        /* 313-343 */
            else assign(o, other, lvl);                  /* unit, 361-367 */
        }
    }
    o->toprop.n = 0;                                     /* 166-174 */
    return conflict;
}

int orc_level0(orc *o)
{
    if (o->contradictory) return 1;
    if (bcp(o)) o->contradictory = 1;
    return o->contradictory;
}

int32_t orc_max_user_var(const orc *o) { return o->max_user_var; }

static inline int mapped(const orc *o, int32_t v) { return (uint64_t)v < o->e2i_cap && o->e2i[v] != 0; }

uint64_t orc_level0_lits(const orc *o, int32_t *out, uint64_t cap)
{
    uint64_t n = 0;
    for (int32_t v = 1; v <= o->max_user_var; ++v) {
        if (!mapped(o, v)) continue;
        const uint32_t iv = var_of(o->e2i[v]);
        if (o->val[iv] && o->level[iv] == 0) { if (out && n < cap) out[n] = o->val[iv] == 1 ? v : -v; ++n; }
    }
    return n;
}

typedef struct { int32_t *p; uint64_t n, cap; } i32big;
static void i32big_push(i32big *v, int32_t x)
{
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->p = (int32_t *)realloc(v->p, v->cap * 4); }
    v->p[v->n++] = x;
}

static int cmp_absvar(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    x = x < 0 ? -x : x; y = y < 0 ? -y : y;
    return (x > y) - (x < y);
}

/* HandleAssumptions + AssignAssumptions (Topi.cc:797-938, 715-769) for one cube */
static uint8_t run_cube(orc *o, const int32_t *lits, uint64_t n, i32big *out)
{
    for (uint64_t i = 0; i < n && lits[i]; ++i) {
        const int32_t v = lits[i] < 0 ? -lits[i] : lits[i];
        if (!mapped(o, v)) return BCP_FLAG_UNMAPPED;
    }
    int any_live = 0;
    for (uint64_t i = 0; i < n && lits[i]; ++i) {
        const int32_t l = lits[i];
        const ulit il = l > 0 ? o->e2i[l] : neg(o->e2i[-l]);
        if (is_assigned(o, il)) { if (is_false(o, il)) return BCP_FLAG_FALSE_L0; } /* 851-860 */
        else any_live = 1;
    }
    if (!any_live) return BCP_FLAG_TRIVIAL_L0;                                     /* 861-864 */
    uint8_t flag = BCP_FLAG_OK;
    for (uint64_t i = 0; i < n && lits[i]; ++i) {
        const int32_t l = lits[i];
        const ulit il = l > 0 ? o->e2i[l] : neg(o->e2i[-l]);
        if (is_assigned(o, il)) {
            if (is_false(o, il)) { flag = BCP_FLAG_CONFLICT; break; }  /* 866-874, 740-746, 918-934 */
            continue;                                                 /* duplicate / implied / true at level 0 */
        }
        ++o->dec_level;                                               /* NewDecLevel, Topi.cc:1529 */
        assign(o, il, o->dec_level);
        if (bcp(o)) { flag = BCP_FLAG_CONFLICT; break; }
    }
    if (flag == BCP_FLAG_OK && out) {
        const uint64_t start = out->n;
        for (uint32_t t = o->trail_l0; t < o->trail.n; ++t) {
            const ulit il = o->trail.p[t];
            const int32_t e = o->i2e[var_of(il)];
            i32big_push(out, (il & 1) ? -e : e);
        }
        qsort(out->p + start, out->n - start, 4, cmp_absvar);
    }
    backtrack0(o);
    return flag;
}

int orc_run_cubes(orc *o, const int32_t *lits, const uint64_t *off, uint64_t n, uint64_t lo, uint64_t hi,
                  uint64_t stride, uint8_t *flag, uint32_t *props, uint64_t *impl_off, int32_t **impl_lits)
{
    if (o->contradictory) return 1;
    i32big out = {0, 0, 0};
    if (stride == 0) stride = 1;
    for (uint64_t i = 0; i < n; ++i) {
        impl_off[i] = out.n;
        flag[i] = 255; if (props) props[i] = 0;
        if (i < lo || i >= hi || (i - lo) % stride) continue;
        const uint64_t before = o->n_impl;
        if (lits) flag[i] = run_cube(o, lits + off[i], off[i + 1] - off[i], impl_lits ? &out : NULL);
        else {
            const int32_t v = (int32_t)(i / 2 + 1);
            const int32_t l = (i & 1) ? -v : v;
            flag[i] = run_cube(o, &l, 1, impl_lits ? &out : NULL);
        }
        if (props) props[i] = (uint32_t)(o->n_impl - before);
    }
    impl_off[n] = out.n;
    if (impl_lits) *impl_lits = out.p; else free(out.p);
    return 0;
}

int orc_probe_all(orc *o, uint64_t lo, uint64_t hi, uint64_t stride, uint8_t *flag, uint32_t *props,
                  uint64_t *impl_off, int32_t **impl_lits)
{
    return orc_run_cubes(o, NULL, NULL, 2ull * (uint64_t)o->max_user_var, lo, hi, stride, flag, props, impl_off, impl_lits);
}

void orc_free(void *p) { free(p); }
uint64_t orc_implications(const orc *o) { return o->n_impl; }
uint64_t orc_assignments(const orc *o) { return o->n_asg; }
uint64_t orc_bcps(const orc *o) { return o->n_bcp; }
