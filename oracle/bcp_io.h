/*
 * bcp_io.h -- TEST INFRASTRUCTURE (oracle side). Tiny binary file formats shared by the
 * reference harness (oracle/ref_harness.cc), the C restatement (oracle/bcp_oracle.c),
 * the synthetic generators (tools/cnfgen.c) and the Python tests/bench (numpy reads them).
 *
 * All integers little-endian.
 *
 *  instance  (.bcnf): u64 magic "BCNF0001", u64 nvars_hint, u64 nclauses, u64 nwords,
 *                     i32 words[nwords]            -- clauses, each 0-terminated
 *  cubes     (.cubes): u64 magic "CUBE0001", u64 ncubes, u64 nlits,
 *                     u64 off[ncubes+1], i32 lits[nlits]
 *  results   (.bres): u64 magic "BRES0001", u64 n, u64 nlits, u64 status, u64 nlevel0,
 *                     u64 implications, u64 assignments, f64 seconds,
 *                     u8 flag[n] (padded to 8), u32 props[n] (padded to 8),
 *                     u64 off[n+1], i32 lits[nlits] (padded to 8), i32 level0[nlevel0]
 *
 * Text instances (the reference's regression_instances cnf files, Main.cc:207-221 format) are read
 * with the convention of SURVEY.md 8c: every clause line in file order; lines starting with
 * s / r / o / l / b / n / c / p are ignored.
 */
#ifndef BCP_IO_H
#define BCP_IO_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BCNF_MAGIC 0x31303030464e4342ull /* "BCNF0001" */
#define CUBE_MAGIC 0x3130303045425543ull /* "CUBE0001" */
#define BRES_MAGIC 0x3130303053455242ull /* "BRES0001" */

/* probe / cube result flags (shared with include/bbcp.h) */
enum {
    BCP_FLAG_OK = 0,          /* fixpoint reached, no conflict */
    BCP_FLAG_CONFLICT = 1,    /* UP derives a conflict (incl. contradictory cube) */
    BCP_FLAG_TRIVIAL_L0 = 2,  /* every cube literal already true at level 0: nothing to do */
    BCP_FLAG_FALSE_L0 = 3,    /* some cube literal is false at level 0 */
    BCP_FLAG_UNMAPPED = 4     /* cube names a variable that occurs in no clause; not propagated */
};

typedef struct {
    int32_t *words;   /* 0-terminated clauses */
    uint64_t nwords;
    uint64_t nclauses;
    uint64_t nvars_hint;
} bcp_instance;

static inline int bcp_has_suffix(const char *s, const char *suf)
{
    size_t n = strlen(s), m = strlen(suf);
    return n >= m && strcmp(s + n - m, suf) == 0;
}

/* returns 0 on success */
static inline int bcp_read_instance(const char *path, bcp_instance *out)
{
    memset(out, 0, sizeof(*out));
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    if (bcp_has_suffix(path, ".bcnf")) {
        uint64_t hdr[4];
        if (fread(hdr, 8, 4, f) != 4 || hdr[0] != BCNF_MAGIC) { fclose(f); return -2; }
        out->nvars_hint = hdr[1];
        out->nclauses = hdr[2];
        out->nwords = hdr[3];
        out->words = (int32_t *)malloc(out->nwords * 4 + 4);
        if (fread(out->words, 4, out->nwords, f) != out->nwords) { fclose(f); return -3; }
        fclose(f);
        return 0;
    }
    /* text */
    size_t cap = 1 << 16;
    out->words = (int32_t *)malloc(cap * 4);
    char *line = NULL;
    size_t lcap = 0;
    ssize_t len;
    while ((len = getline(&line, &lcap, f)) >= 0) {
        char *p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (!(*p == '-' || (*p >= '0' && *p <= '9'))) continue; /* s r o l b n c p, blank */
        int any = 0;
        for (;;) {
            char *e;
            long v = strtol(p, &e, 10);
            if (e == p) break;
            p = e;
            if (out->nwords + 2 > cap) { cap *= 2; out->words = (int32_t *)realloc(out->words, cap * 4); }
            any = 1;
            out->words[out->nwords++] = (int32_t)v;
            if (v == 0) { out->nclauses++; any = 0; break; }
            long a = v < 0 ? -v : v;
            if ((uint64_t)a > out->nvars_hint) out->nvars_hint = (uint64_t)a;
        }
        if (any) { /* clause line without trailing 0: terminate it */
            out->words[out->nwords++] = 0;
            out->nclauses++;
        }
    }
    free(line);
    fclose(f);
    return 0;
}

typedef struct {
    uint64_t ncubes, nlits;
    uint64_t *off;
    int32_t *lits;
} bcp_cubes;

static inline int bcp_read_cubes(const char *path, bcp_cubes *c)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    uint64_t hdr[3];
    if (fread(hdr, 8, 3, f) != 3 || hdr[0] != CUBE_MAGIC) { fclose(f); return -2; }
    c->ncubes = hdr[1];
    c->nlits = hdr[2];
    c->off = (uint64_t *)malloc((c->ncubes + 1) * 8);
    c->lits = (int32_t *)malloc(c->nlits * 4 + 4);
    if (fread(c->off, 8, c->ncubes + 1, f) != c->ncubes + 1) { fclose(f); return -3; }
    if (fread(c->lits, 4, c->nlits, f) != c->nlits) { fclose(f); return -3; }
    fclose(f);
    return 0;
}

static inline void bcp_pad8(FILE *f, uint64_t nbytes)
{
    static const char z[8] = {0};
    uint64_t r = nbytes & 7;
    if (r) fwrite(z, 1, 8 - r, f);
}

static inline int bcp_write_results(const char *path, uint64_t n, uint64_t status,
                                    const uint8_t *flag, const uint32_t *props, const uint64_t *off,
                                    const int32_t *lits, const int32_t *level0, uint64_t nlevel0,
                                    uint64_t implications, uint64_t assignments, double seconds)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    uint64_t nlits = off ? off[n] : 0;
    uint64_t hdr[7] = {BRES_MAGIC, n, nlits, status, nlevel0, implications, assignments};
    fwrite(hdr, 8, 7, f);
    fwrite(&seconds, 8, 1, f);
    fwrite(flag, 1, n, f); bcp_pad8(f, n);
    fwrite(props, 4, n, f); bcp_pad8(f, n * 4);
    if (off) fwrite(off, 8, n + 1, f); else { uint64_t z = 0; for (uint64_t i = 0; i <= n; ++i) fwrite(&z, 8, 1, f); }
    fwrite(lits, 4, nlits, f); bcp_pad8(f, nlits * 4);
    fwrite(level0, 4, nlevel0, f);
    fclose(f);
    return 0;
}

#endif
