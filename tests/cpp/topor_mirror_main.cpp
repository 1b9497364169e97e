// Drives the engine through the C++ mirror of the reference facade (include/bbcp_topor.hpp), the way the
// reference's own users drive CTopor: AddClause per clause, Solve(assumps), GetLitValue; then one ProbeAll.
// usage: topor_mirror <words.i32> <out.bin>
//   words.i32: int32 stream of 0-terminated clauses.  out.bin: u64 n, u64 nlits, u8 flag[n], u64 off[n+1], i32 lits[nlits]
// exit code: 0 ok, 3 no CUDA device / engine error (message on stderr), 4 contradictory instance
#include <cstdio>
#include <vector>

#include "bbcp_topor.hpp"

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    std::vector<int32_t> words;
    if (FILE *f = fopen(argv[1], "rb")) {
        int32_t buf[4096];
        size_t k;
        while ((k = fread(buf, 4, 4096, f)) > 0) words.insert(words.end(), buf, buf + k);
        fclose(f);
    } else return 2;
    bbcp::CBatchedTopor t;
    std::vector<int32_t> cur;
    for (int32_t w : words) {
        if (w) { cur.push_back(w); continue; }
        t.AddClause(cur);
        cur.clear();
    }
    const int rc = t.Finalize();
    if (rc == BBCP_CONTRADICTORY) return 4;
    if (rc != BBCP_OK || t.IsError()) { fprintf(stderr, "engine error %d: %s\n", rc, t.GetStatusExplanation().c_str()); return 3; }
    // single-cube path: assume +1, read back its value
    const bbcp::TToporReturnVal r = t.Solve({1});
    if (r != bbcp::TToporReturnVal::RET_UNSAT && t.GetLitValue(1) == bbcp::TToporLitVal::VAL_UNSATISFIED) return 5;
    t.Backtrack(0);
    bbcp::BatchResult res = t.ProbeAll();
    if (res.rc != BBCP_OK) { fprintf(stderr, "probe error %d: %s\n", res.rc, t.GetStatusExplanation().c_str()); return 3; }
    FILE *o = fopen(argv[2], "wb");
    if (!o) return 2;
    const uint64_t n = res.info.n, nl = res.info.n_implied;
    fwrite(&n, 8, 1, o);
    fwrite(&nl, 8, 1, o);
    fwrite(res.flag.data(), 1, n, o);
    fwrite(res.impl_off.data(), 8, n + 1, o);
    fwrite(res.impl_lits.data(), 4, nl, o);
    fclose(o);
    printf("{\"n\": %llu, \"n_implied\": %llu, \"failed\": %zu, \"propagations\": %llu}\n", (unsigned long long)n, (unsigned long long)nl,
           t.FailedLiterals().size(), (unsigned long long)t.GetPropagations());
    return 0;
}
