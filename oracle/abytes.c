/*
 * abytes.c -- TEST / MEASUREMENT INFRASTRUCTURE (oracle side). SURVEY.md 8d's ABYTES: the order-independent lower
 * bound on the bytes any two-watched-literal BCP must move for a batch of probes / cubes, computed on the CPU from
 * the clause database and the CHECKER's fixpoints (the .bres files written by oracle/_ref/ref_bcp or oracle_bcp), so
 * that the roofline figure of bench.py's workload blocks does not rest on engine counters. The C form of
 * tests/abytes.py (same definition, checked against it in tests/test_abytes.py), fast enough for the strided samples
 * of the BASELINE-size instances.
 *
 *   ABYTES(item) = sum over p in T of [ 8 + 4*nbin(not p) + 8*nlong(not p)
 *                                       + sum over clauses C (>= 3 literals) watched by not-p whose other watch is
 *                                         not true in T of 4*(1+|C|) ]
 *                  + 4*|T| + 4*|cube|            (conflicting items: 4*|cube| only)
 *   on the level-0-simplified database (satisfied clauses dropped, false literals stripped, duplicate binaries once),
 *   canonical watches = the first two remaining literals of each clause in input order.
 *
 * usage: abytes <instance(.cnf|.bcnf)> [--cubes f.cubes] <results.bres>...
 *        prints {"items": ran, "ok": n, "abytes_total": bytes, "lits": sum |T|}
 */
#include "bcp_io.h"

typedef struct { uint32_t other, len; } watch_t;

static int cmp_u64(const void *a, const void *b)
{
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

static inline uint32_t ilit(int32_t l) { return l < 0 ? 2u * (uint32_t)(-l) + 1u : 2u * (uint32_t)l; }

typedef struct {
    uint64_t n, nlits, status, nlevel0;
    uint8_t *flag;
    uint64_t *off;
    int32_t *lits, *level0;
} bres_t;

static int read_bres(const char *path, bres_t *r)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    uint64_t hdr[8];
    if (fread(hdr, 8, 8, f) != 8 || hdr[0] != BRES_MAGIC) { fclose(f); return -2; }
    r->n = hdr[1]; r->nlits = hdr[2]; r->status = hdr[3]; r->nlevel0 = hdr[4];
    const uint64_t n = r->n;
    r->flag = (uint8_t *)malloc(((n + 7) & ~7ull) + 8);
    r->off = (uint64_t *)malloc((n + 1) * 8);
    r->lits = (int32_t *)malloc(((r->nlits * 4 + 7) & ~7ull) + 8);
    r->level0 = (int32_t *)malloc(r->nlevel0 * 4 + 8);
    int ok = fread(r->flag, 1, (n + 7) & ~7ull, f) == ((n + 7) & ~7ull);
    ok = ok && fseek(f, (long)((n * 4 + 7) & ~7ull), SEEK_CUR) == 0;           /* props */
    ok = ok && fread(r->off, 8, n + 1, f) == n + 1;
    ok = ok && fread(r->lits, 1, (r->nlits * 4 + 7) & ~7ull, f) == ((r->nlits * 4 + 7) & ~7ull);
    ok = ok && fread(r->level0, 4, r->nlevel0, f) == r->nlevel0;
    fclose(f);
    return ok ? 0 : -3;
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <instance> [--cubes f] <results.bres>...\n", argv[0]); return 2; }
    const char *cubes_path = NULL;
    int first_res = 2;
    if (!strcmp(argv[2], "--cubes") && argc >= 5) { cubes_path = argv[3]; first_res = 4; }
    bcp_instance ins;
    if (bcp_read_instance(argv[1], &ins)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 3; }
    bcp_cubes cubes; memset(&cubes, 0, sizeof(cubes));
    if (cubes_path && bcp_read_cubes(cubes_path, &cubes)) { fprintf(stderr, "cannot read %s\n", cubes_path); return 3; }
    bres_t r0;
    if (read_bres(argv[first_res], &r0)) { fprintf(stderr, "cannot read %s\n", argv[first_res]); return 3; }

    uint32_t maxv = 0;
    for (uint64_t i = 0; i < ins.nwords; ++i) { const uint32_t a = (uint32_t)abs(ins.words[i]); if (a > maxv) maxv = a; }
    const size_t nl = 2 * ((size_t)maxv + 1);
    /* level-0 values: 1 = literal true, 2 = literal false */
    uint8_t *val0 = (uint8_t *)calloc(nl, 1);
    for (uint64_t i = 0; i < r0.nlevel0; ++i) { const uint32_t l = ilit(r0.level0[i]); if (l < nl) { val0[l] = 1; val0[l ^ 1] = 2; } }

    /* pass 1 over the clauses: simplified form; collect binaries (deduplicated by sorting), count canonical watches */
    uint32_t *nbin = (uint32_t *)calloc(nl, 4), *nw = (uint32_t *)calloc(nl + 1, 4);
    uint64_t *bins = (uint64_t *)malloc((ins.nclauses + 1) * 8), nbins = 0;
    uint64_t *stamp = (uint64_t *)calloc(nl, 8), cnt = 0;
    uint32_t *cur = (uint32_t *)malloc(((size_t)maxv + 2) * 2 * 4);
    /* simplified clauses of >= 3 literals are re-derived in pass 2 (same code), so only counts are kept here */
    for (int pass = 1; pass <= 2; ++pass) {
        uint64_t *wpos = NULL;
        watch_t *wl = NULL;
        static watch_t *g_wl; static uint64_t *g_wstart;
        if (pass == 2) {
            g_wstart = (uint64_t *)malloc((nl + 1) * 8);
            g_wstart[0] = 0;
            for (size_t l = 0; l < nl; ++l) g_wstart[l + 1] = g_wstart[l] + nw[l];
            g_wl = (watch_t *)malloc((g_wstart[nl] + 1) * sizeof(watch_t));
            wpos = (uint64_t *)malloc(nl * 8);
            memcpy(wpos, g_wstart, nl * 8);
            wl = g_wl;
        }
        for (uint64_t s = 0; s < ins.nwords;) {
            uint64_t e = s;
            while (e < ins.nwords && ins.words[e] != 0) ++e;
            ++cnt;
            uint32_t k = 0;
            int drop = 0;
            for (uint64_t i = s; i < e && !drop; ++i) {
                const uint32_t l = ilit(ins.words[i]);
                if (stamp[l ^ 1] == cnt) { drop = 1; break; }          /* tautology */
                if (stamp[l] == cnt) continue;                         /* duplicate literal */
                stamp[l] = cnt;
                cur[k++] = l;
            }
            s = e + 1;
            if (drop) continue;
            uint32_t m = 0;
            for (uint32_t i = 0; i < k; ++i) { if (val0[cur[i]] == 1) { drop = 1; break; } if (val0[cur[i]] == 0) cur[m++] = cur[i]; }
            if (drop || m < 2) continue;
            if (m == 2) {
                if (pass == 1) { const uint32_t a = cur[0] < cur[1] ? cur[0] : cur[1], b = cur[0] < cur[1] ? cur[1] : cur[0]; bins[nbins++] = ((uint64_t)a << 32) | b; }
                continue;
            }
            if (pass == 1) { ++nw[cur[0]]; ++nw[cur[1]]; }
            else {
                wl[wpos[cur[0]]++] = (watch_t){cur[1], m};
                wl[wpos[cur[1]]++] = (watch_t){cur[0], m};
            }
        }
        if (pass == 1) {
            qsort(bins, nbins, 8, cmp_u64);
            for (uint64_t i = 0; i < nbins; ++i) {
                if (i && bins[i] == bins[i - 1]) continue;
                ++nbin[(uint32_t)(bins[i] >> 32)];
                ++nbin[(uint32_t)bins[i]];
            }
            free(bins);
        } else {
            free(wpos);
            /* the per-item sums */
            const watch_t *W = g_wl;
            const uint64_t *WS = g_wstart;
            uint64_t *mark = stamp;   /* reuse: mark[l] == token <=> l in T of the current item */
            memset(mark, 0, nl * 8);
            uint64_t token = 0, items = 0, ok = 0, total = 0, nlits = 0;
            for (int a = first_res; a < argc; ++a) {
                bres_t r;
                if (a == first_res) r = r0; else if (read_bres(argv[a], &r)) { fprintf(stderr, "cannot read %s\n", argv[a]); return 3; }
                for (uint64_t i = 0; i < r.n; ++i) {
                    if (r.flag[i] == 255) continue;
                    ++items;
                    const uint64_t cube = cubes_path ? cubes.off[i + 1] - cubes.off[i] : 1;
                    total += 4 * cube;
                    if (r.flag[i] != 0) continue;
                    ++ok;
                    ++token;
                    const int32_t *T = r.lits + r.off[i];
                    const uint64_t nT = r.off[i + 1] - r.off[i];
                    for (uint64_t j = 0; j < nT; ++j) mark[ilit(T[j])] = token;
                    for (uint64_t j = 0; j < nT; ++j) {
                        const uint32_t q = ilit(T[j]) ^ 1u;            /* the row of not-p is walked when p becomes true */
                        const uint64_t b = WS[q], e2 = WS[q + 1];
                        total += 8 + 4ull * nbin[q] + 8ull * (e2 - b);
                        for (uint64_t w = b; w < e2; ++w) if (mark[W[w].other] != token) total += 4ull * (1 + W[w].len);
                    }
                    total += 4 * nT;
                    nlits += nT;
                }
                free(r.flag); free(r.off); free(r.lits); free(r.level0);
            }
            printf("{\"items\": %llu, \"ok\": %llu, \"abytes_total\": %llu, \"lits\": %llu}\n", (unsigned long long)items, (unsigned long long)ok,
                   (unsigned long long)total, (unsigned long long)nlits);
        }
    }
    return 0;
}
