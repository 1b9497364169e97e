/*
 * oracle_main.c -- TEST INFRASTRUCTURE. Command-line driver of the plain-C restatement with the same
 * arguments and output format as oracle/_ref/ref_bcp (oracle/ref_harness.cc), so either can serve as
 * the CPU baseline process (one process per host core over probe shards).
 *
 * usage: oracle_bcp <instance(.cnf|.bcnf)> --out <file.bres> [--probe-all | --cubes <f.cubes>]
 *                   [--shard i n] [--stride k] [--flags-only]
 */
#include "bcp_oracle.h"
#include "bcp_io.h"
#include <time.h>

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <instance> --out f [--probe-all|--cubes f] [--shard i n] [--stride k] [--flags-only]\n", argv[0]); return 2; }
    const char *out = NULL, *cubes_path = NULL;
    uint64_t shard_i = 0, shard_n = 1, stride = 1, repeat = 1;
    int flags_only = 0;
    for (int i = 2; i < argc; ++i) {
        if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--cubes") && i + 1 < argc) cubes_path = argv[++i];
        else if (!strcmp(argv[i], "--probe-all")) cubes_path = NULL;
        else if (!strcmp(argv[i], "--shard") && i + 2 < argc) { shard_i = strtoull(argv[i + 1], 0, 10); shard_n = strtoull(argv[i + 2], 0, 10); i += 2; }
        else if (!strcmp(argv[i], "--stride") && i + 1 < argc) stride = strtoull(argv[++i], 0, 10);
        else if (!strcmp(argv[i], "--flags-only")) flags_only = 1;
        else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = strtoull(argv[++i], 0, 10);
        else { fprintf(stderr, "bad arg %s\n", argv[i]); return 2; }
    }
    bcp_instance ins;
    if (bcp_read_instance(argv[1], &ins)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 3; }
    orc *o = orc_create();
    orc_add_words(o, ins.words, ins.nwords);
    free(ins.words);
    bcp_cubes cubes; memset(&cubes, 0, sizeof(cubes));
    if (cubes_path && bcp_read_cubes(cubes_path, &cubes)) { fprintf(stderr, "cannot read %s\n", cubes_path); return 3; }
    const int status = orc_level0(o);
    const uint64_t n = cubes_path ? cubes.ncubes : 2ull * (uint64_t)orc_max_user_var(o);
    uint8_t *flag = (uint8_t *)malloc(n + 8); memset(flag, 255, n + 8);
    uint32_t *props = (uint32_t *)calloc(n + 2, 4);
    uint64_t *off = (uint64_t *)calloc(n + 1, 8);
    int32_t *lits = NULL;
    uint64_t nl0 = orc_level0_lits(o, NULL, 0);
    int32_t *l0 = (int32_t *)malloc(nl0 * 4 + 4);
    orc_level0_lits(o, l0, nl0);
    const uint64_t impl0 = orc_implications(o), asg0 = orc_assignments(o);
    double seconds = 0;
    char slist[4096] = "";
    int slen = 0;
    uint64_t ran = 0, conflicts = 0;
    if (!status) {
        const uint64_t lo = n * shard_i / shard_n, hi = n * (shard_i + 1) / shard_n;
        for (uint64_t rep = 0; rep < repeat; ++rep) {
            if (lits) { orc_free(lits); lits = NULL; }
            const double t0 = now_s();
            orc_run_cubes(o, cubes_path ? cubes.lits : NULL, cubes.off, n, lo, hi, stride, flag, props, off, flags_only ? NULL : &lits);
            seconds = now_s() - t0;
            slen += snprintf(slist + slen, sizeof(slist) - slen, "%s%.6f", rep ? ", " : "", seconds);
            if (slen > (int)sizeof(slist) - 32) break;
        }
        for (uint64_t i = 0; i < n; ++i) { ran += flag[i] != 255; conflicts += flag[i] == BCP_FLAG_CONFLICT; }
    } else nl0 = 0;
    const uint64_t impl = orc_implications(o) - impl0, asg = orc_assignments(o) - asg0;
    if (out && bcp_write_results(out, n, (uint64_t)status, flag, props, off, lits, l0, nl0, impl, asg, seconds)) { fprintf(stderr, "cannot write %s\n", out); return 4; }
    printf("{\"status\": %d, \"n\": %llu, \"ran\": %llu, \"conflicts\": %llu, \"implications\": %llu, \"assignments\": %llu, \"implied_lits\": %llu, \"level0\": %llu, \"delayed_triggers\": 0, \"seconds\": %.6f, \"seconds_list\": [%s]}\n",
           status, (unsigned long long)n, (unsigned long long)ran, (unsigned long long)conflicts, (unsigned long long)impl,
           (unsigned long long)asg, (unsigned long long)off[n], (unsigned long long)nl0, seconds, slist);
    orc_release(o);
    return 0;
}
