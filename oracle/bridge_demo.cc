// bridge_demo.cc -- TEST INFRASTRUCTURE. Binds include/bbcp_bridge.hpp to the UNMODIFIED reference solver
// (Topor::CTopor<int32_t,uint32_t,false>, compiled where it lies by oracle/Makefile) and checks, on one
// instance, that the reference answers the same with and without the GPU probing pass in front of it:
//   for every query (the file's own "s ..." lines, or one query without assumptions)
//      plain:  fresh CTopor, AddClause*, Solve(assumps)
//      bridge: CToporWithGpuProbing<CTopor>, AddClause*, GpuProbe() [GPU: probing to a fixpoint + products,
//              units fed back through AddClause / GetNextUnitClause], Solve(assumps)
// and that the cube filter keeps every cube the reference finds satisfiable.
// With --learnts the CPU solver's learnt clauses (SetCbNewLearntCls) flow into the GPU database and a GPU pass runs inside
// the running Solve every 20 learnt clauses (units back through GetNextUnitClause); with --drat <file> every unit fed to the
// CPU solver is logged as a DRAT text line in derivation order (tests/test_bridge.py re-checks each one by unit propagation).
// usage: bridge_demo <instance(.cnf | .bcnf)> [--cubes f.cubes] [--learnts] [--drat f]   -> one JSON line; exit 0 iff all answers agree
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "Topor.hpp"
#include "bbcp_bridge.hpp"
#include "bcp_io.h"

using Ref = Topor::CTopor<int32_t, uint32_t, false>;

static std::vector<std::vector<int32_t>> clauses_of(const bcp_instance &ins)
{
    std::vector<std::vector<int32_t>> out;
    std::vector<int32_t> cur;
    for (uint64_t i = 0; i < ins.nwords; ++i) {
        if (ins.words[i]) { cur.push_back(ins.words[i]); continue; }
        out.push_back(cur);
        cur.clear();
    }
    return out;
}

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    bcp_instance ins;
    if (bcp_read_instance(argv[1], &ins) != 0) { fprintf(stderr, "cannot read %s\n", argv[1]); return 3; }
    const auto clauses = clauses_of(ins);
    std::vector<std::vector<int32_t>> queries;
    bool learnts = false;
    std::ofstream drat;
    for (int a = 2; a < argc; ++a) {
        if (!strcmp(argv[a], "--cubes") && a + 1 < argc) {
            bcp_cubes cubes;
            memset(&cubes, 0, sizeof(cubes));
            if (bcp_read_cubes(argv[++a], &cubes) != 0) { fprintf(stderr, "cannot read %s\n", argv[a]); return 3; }
            for (uint64_t i = 0; i < cubes.ncubes; ++i) queries.emplace_back(cubes.lits + cubes.off[i], cubes.lits + cubes.off[i + 1]);
        } else if (!strcmp(argv[a], "--learnts")) learnts = true;
        else if (!strcmp(argv[a], "--drat") && a + 1 < argc) drat.open(argv[++a]);
    }
    if (queries.empty()) queries.push_back({});

    bbcp::CToporWithGpuProbing<Ref> bridge;
    if (learnts) bridge.IngestLearntClauses(Topor::TStopTopor::VAL_CONTINUE, 8, 20);
    if (drat.is_open()) bridge.SetDratLog(&drat);
    for (const auto &c : clauses) bridge.AddClause(c);
    const long fed = bridge.GpuProbe();
    if (fed == -2) { fprintf(stderr, "engine error: %s\n", bridge.Gpu().GetStatusExplanation().c_str()); return 4; }

    // cube filter on the GPU (before the CPU sees any cube)
    std::vector<int32_t> lits;
    std::vector<uint64_t> off = {0};
    for (const auto &q : queries) { lits.insert(lits.end(), q.begin(), q.end()); off.push_back(lits.size()); }
    std::vector<char> alive(queries.size(), 0);
    for (size_t i : bridge.FilterCubes(lits, off)) alive[i] = 1;

    size_t agree = 0, sat = 0, unsat = 0, filtered = 0, filter_wrong = 0;
    for (size_t qi = 0; qi < queries.size(); ++qi) {
        std::vector<int32_t> q = queries[qi];
        Ref plain;
        for (auto c : clauses) plain.AddClause(std::span<int32_t>(c));
        const auto a = plain.Solve(std::span<int32_t>(q));
        const auto b = bridge.Solve(q, /*probeFirst=*/false);
        agree += a == b;
        sat += a == Topor::TToporReturnVal::RET_SAT;
        unsat += a == Topor::TToporReturnVal::RET_UNSAT;
        if (!alive[qi]) { ++filtered; filter_wrong += a != Topor::TToporReturnVal::RET_UNSAT; }
    }
    if (drat.is_open()) drat.close();
    printf("{\"queries\": %zu, \"agree\": %zu, \"sat\": %zu, \"unsat\": %zu, \"units_fed\": %ld, \"cubes_filtered\": %zu, \"filter_wrong\": %zu, \"learnts_taken\": %llu, \"in_solve_passes\": %llu, \"units_total\": %llu}\n",
           queries.size(), agree, sat, unsat, fed, filtered, filter_wrong, (unsigned long long)bridge.LearntsTaken(), (unsigned long long)bridge.InSolvePasses(),
           (unsigned long long)bridge.UnitsFed());
    return agree == queries.size() && filter_wrong == 0 ? 0 : 1;
}
