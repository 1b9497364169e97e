/*
 * cnfgen.c -- deterministic synthetic CNF generators for the bench / parity configurations of
 * BASELINE.json. One tool, used by the engine's bench, the GPU parity tests and the
 * CPU baseline alike, so every arm sees byte-identical instances.
 *
 * PRNG: xoshiro256** seeded through splitmix64.
 *
 *   cnfgen 3sat   n=<vars> m=<clauses> seed=<s> out=<f.bcnf>
 *        C2: each clause 3 distinct variables uniform on [1,n], each sign a fair coin
 *   cnfgen aig    frames=<F> gates=<G per frame> inputs=<P per frame> latches=<L> window=<W> seed=<s> out=<f.bcnf>
 *        C3: BMC-style unrolling of a random AND-inverter graph, Tseitin-encoded:
 *        gate g = a & b  ->  (-g a) (-g b) (g -a -b). Frame k holds L latch signals (frame 0: free
 *        variables; later: outputs of uniformly chosen gates of frame k-1, i.e. the same variables),
 *        P fresh primary inputs and G gates; the fan-ins of a gate are drawn uniformly from the L latch
 *        signals and the W signals (inputs or gates) preceding it in its frame, each inverted with prob 1/2
 *   cnfgen comm   n=<vars> m=<clauses> comms=<K> seed=<s> out=<f.bcnf> [cubes=<f.cubes> ncubes=<C> split=<S> extra=<E> cseed=<s2>]
 *        C4: clause length 2/3/4/5 w.p. .15/.55/.20/.10; w.p. 0.9 all variables of a clause come from one
 *        community (K equal, contiguous communities), else uniform; distinct variables; fair signs.
 *        Cubes: the S highest-occurrence variables are the split variables; cube i takes the polarity of
 *        split variable j from bit j of i (a complete lookahead tree of depth S when C = 2^S), plus E
 *        pseudo-random literals per cube.
 *   cnfgen plaw   n=<vars> m=<clauses> alpha=<0.8> seed=<s> out=<f.bcnf>
 *        C5: variable drawn with probability ~ rank^(-alpha) (rank spread over ids by a fixed
 *        multiplicative permutation), clause length 2 / 3 / uniform 4..8 w.p. .4/.4/.2; distinct vars
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BCNF_MAGIC 0x31303030464e4342ull
#define CUBE_MAGIC 0x3130303045425543ull

static uint64_t s[4];
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix(uint64_t *x)
{
    uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static void seed_rng(uint64_t seed) { for (int i = 0; i < 4; ++i) s[i] = splitmix(&seed); }
static inline uint64_t next64(void)
{
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
}
/* uniform on [0,n) by 128-bit multiply (bias < 2^-32 for n < 2^32: irrelevant here, and deterministic) */
static inline uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next64() * n) >> 64); }
static inline double unit(void) { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }

typedef struct { int32_t *p; uint64_t n, cap, ncls; } words_t;
static void push(words_t *w, int32_t x)
{
    if (w->n == w->cap) { w->cap = w->cap ? w->cap * 2 : (1u << 20); w->p = (int32_t *)realloc(w->p, w->cap * 4); if (!w->p) { fprintf(stderr, "oom\n"); exit(1); } }
    w->p[w->n++] = x;
}
static void end_clause(words_t *w) { push(w, 0); w->ncls++; }

static const char *arg(int argc, char **argv, const char *key, const char *def)
{
    size_t k = strlen(key);
    for (int i = 2; i < argc; ++i) if (!strncmp(argv[i], key, k) && argv[i][k] == '=') return argv[i] + k + 1;
    return def;
}
static uint64_t argu(int argc, char **argv, const char *key, uint64_t def)
{
    const char *v = arg(argc, argv, key, NULL);
    return v ? strtoull(v, 0, 10) : def;
}

static int write_bcnf(const char *path, const words_t *w, uint64_t nvars)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    uint64_t hdr[4] = {BCNF_MAGIC, nvars, w->ncls, w->n};
    fwrite(hdr, 8, 4, f);
    fwrite(w->p, 4, w->n, f);
    fclose(f);
    return 0;
}

static int distinct(const int32_t *v, int k, int32_t x) { for (int i = 0; i < k; ++i) if (v[i] == x) return 0; return 1; }

static void gen_3sat(words_t *w, uint64_t n, uint64_t m)
{
    for (uint64_t c = 0; c < m; ++c) {
        int32_t v[3];
        for (int k = 0; k < 3;) { int32_t x = (int32_t)(1 + below(n)); if (distinct(v, k, x)) v[k++] = x; }
        for (int k = 0; k < 3; ++k) push(w, (next64() & 1) ? -v[k] : v[k]);
        end_clause(w);
    }
}

static uint64_t gen_aig(words_t *w, uint64_t F, uint64_t G, uint64_t P, uint64_t L, uint64_t W)
{
    int32_t next_var = 1;
    int32_t *latch = (int32_t *)malloc(L * 4 + 4), *nlatch = (int32_t *)malloc(L * 4 + 4);
    int32_t *sig = (int32_t *)malloc((P + G) * 4 + 4);
    for (uint64_t i = 0; i < L; ++i) latch[i] = next_var++;
    for (uint64_t f = 0; f < F; ++f) {
        uint64_t ns = 0;
        for (uint64_t i = 0; i < P; ++i) sig[ns++] = next_var++;
        for (uint64_t g = 0; g < G; ++g) {
            const uint64_t win = ns < W ? ns : W;
            int32_t in[2];
            for (int k = 0; k < 2;) {
                const uint64_t r = below(L + win);
                const int32_t x = r < L ? latch[r] : sig[ns - win + (r - L)];
                if (k == 1 && x == in[0]) continue;
                in[k++] = x;
            }
            const int32_t a = (next64() & 1) ? -in[0] : in[0], b = (next64() & 1) ? -in[1] : in[1];
            const int32_t gv = next_var++;
            push(w, -gv); push(w, a); end_clause(w);
            push(w, -gv); push(w, b); end_clause(w);
            push(w, gv); push(w, -a); push(w, -b); end_clause(w);
            sig[ns++] = gv;
        }
        for (uint64_t i = 0; i < L; ++i) nlatch[i] = sig[P + below(G)];
        int32_t *t = latch; latch = nlatch; nlatch = t;
    }
    free(latch); free(nlatch); free(sig);
    return (uint64_t)next_var - 1;
}

static int pick_len_comm(void) { const double u = unit(); return u < .15 ? 2 : u < .70 ? 3 : u < .90 ? 4 : 5; }

static void gen_comm(words_t *w, uint64_t n, uint64_t m, uint64_t K)
{
    const uint64_t csz = n / K;
    for (uint64_t c = 0; c < m; ++c) {
        const int len = pick_len_comm();
        const int inside = unit() < 0.9;
        const uint64_t com = below(K);
        int32_t v[8];
        for (int k = 0; k < len;) {
            const int32_t x = (int32_t)(inside ? 1 + com * csz + below(csz) : 1 + below(n));
            if (distinct(v, k, x)) v[k++] = x;
        }
        for (int k = 0; k < len; ++k) push(w, (next64() & 1) ? -v[k] : v[k]);
        end_clause(w);
    }
}

static void gen_plaw(words_t *w, uint64_t n, uint64_t m, double alpha)
{
    const double e = 1.0 - alpha, top = pow((double)n + 1.0, e) - 1.0;
    uint64_t A = (uint64_t)(0.6180339887 * (double)n) | 1;
    for (;;) { uint64_t a = A, b = n; while (b) { uint64_t t = a % b; a = b; b = t; } if (a == 1) break; A += 2; }
    for (uint64_t c = 0; c < m; ++c) {
        const double u = unit();
        const int len = u < .4 ? 2 : u < .8 ? 3 : 4 + (int)below(5);
        int32_t v[8];
        for (int k = 0; k < len;) {
            uint64_t rank = (uint64_t)pow(1.0 + unit() * top, 1.0 / e);
            if (rank < 1) rank = 1;
            if (rank > n) rank = n;
            const int32_t x = (int32_t)(1 + (uint64_t)(((__uint128_t)rank * A) % n));
            if (distinct(v, k, x)) v[k++] = x;
        }
        for (int k = 0; k < len; ++k) push(w, (next64() & 1) ? -v[k] : v[k]);
        end_clause(w);
    }
}

static int write_comm_cubes(const char *path, const words_t *w, uint64_t n, uint64_t C, uint64_t S, uint64_t E, uint64_t cseed)
{
    uint32_t *occ = (uint32_t *)calloc(n + 1, 4);
    for (uint64_t i = 0; i < w->n; ++i) if (w->p[i]) occ[w->p[i] < 0 ? -w->p[i] : w->p[i]]++;
    int32_t *split = (int32_t *)malloc(S * 4 + 4);
    for (uint64_t j = 0; j < S; ++j) { /* S is tiny: selection by repeated max, ties to the lower id */
        uint32_t best = 0; int32_t bv = 0;
        for (uint64_t v = 1; v <= n; ++v) if (occ[v] > best) { best = occ[v]; bv = (int32_t)v; }
        split[j] = bv; occ[bv] = 0;
    }
    seed_rng(cseed);
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    const uint64_t depth = S + E;
    uint64_t hdr[3] = {CUBE_MAGIC, C, C * depth};
    fwrite(hdr, 8, 3, f);
    for (uint64_t i = 0; i <= C; ++i) { uint64_t o = i * depth; fwrite(&o, 8, 1, f); }
    for (uint64_t i = 0; i < C; ++i) {
        for (uint64_t j = 0; j < S; ++j) { int32_t l = ((i >> j) & 1) ? -split[j] : split[j]; fwrite(&l, 4, 1, f); }
        for (uint64_t j = 0; j < E; ++j) { int32_t v = (int32_t)(1 + below(n)); int32_t l = (next64() & 1) ? -v : v; fwrite(&l, 4, 1, f); }
    }
    fclose(f);
    free(occ); free(split);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: cnfgen <3sat|aig|comm|plaw> key=value... out=<f.bcnf>\n"); return 2; }
    const char *kind = argv[1];
    const char *out = arg(argc, argv, "out", NULL);
    if (!out) { fprintf(stderr, "out= missing\n"); return 2; }
    words_t w = {0, 0, 0, 0};
    uint64_t nvars = 0;
    seed_rng(argu(argc, argv, "seed", 1));
    if (!strcmp(kind, "3sat")) {
        nvars = argu(argc, argv, "n", 1000000);
        gen_3sat(&w, nvars, argu(argc, argv, "m", 4200000));
    } else if (!strcmp(kind, "aig")) {
        nvars = gen_aig(&w, argu(argc, argv, "frames", 50), argu(argc, argv, "gates", 66667), argu(argc, argv, "inputs", 2000),
                        argu(argc, argv, "latches", 2000), argu(argc, argv, "window", 5000));
    } else if (!strcmp(kind, "comm")) {
        nvars = argu(argc, argv, "n", 1250000);
        gen_comm(&w, nvars, argu(argc, argv, "m", 5000000), argu(argc, argv, "comms", 1000));
        const char *cubes = arg(argc, argv, "cubes", NULL);
        if (cubes && write_comm_cubes(cubes, &w, nvars, argu(argc, argv, "ncubes", 65536), argu(argc, argv, "split", 16),
                                      argu(argc, argv, "extra", 4), argu(argc, argv, "cseed", 4))) { fprintf(stderr, "cannot write %s\n", cubes); return 4; }
    } else if (!strcmp(kind, "plaw")) {
        nvars = argu(argc, argv, "n", 12500000);
        gen_plaw(&w, nvars, argu(argc, argv, "m", 50000000), atof(arg(argc, argv, "alpha", "0.8")));
    } else { fprintf(stderr, "unknown kind %s\n", kind); return 2; }
    if (write_bcnf(out, &w, nvars)) { fprintf(stderr, "cannot write %s\n", out); return 4; }
    printf("{\"kind\": \"%s\", \"nvars\": %llu, \"nclauses\": %llu, \"nwords\": %llu}\n", kind, (unsigned long long)nvars,
           (unsigned long long)w.ncls, (unsigned long long)w.n);
    return 0;
}
