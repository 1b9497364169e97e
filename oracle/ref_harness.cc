// ref_harness.cc -- TEST INFRASTRUCTURE. "Route A" oracle driver: derives from the
// reference's own solver class and calls its own, unmodified, protected
// NewDecLevel()/Assign()/BCP()/Backtrack() per probe or cube. Nothing of the reference is copied
// into this repository: the reference sources are compiled where they lie (oracle/Makefile,
// outputs under oracle/_ref/ only) and this file only #includes the reference's public header.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// execute the resulting binary. The product (intel_sat_solver_b200/) never does.
//
// Reference entry points driven (file:line in /root/reference):
//   AddUserClause            Topi.cc:377-653      clause intake / normalisation
//   ReserveVarAndLitData     TopiCompression.cc:13-35   (sizes per-decision-level arrays)
//   BCP                      TopiBcp.cc:103-410   the hot path
//   NewDecLevel              Topi.cc:1529-1536
//   Assign                   TopiAsg.cc:46-139
//   Backtrack                TopiBacktrack.cc:18-53
// Cube semantics restated from HandleAssumptions/AssignAssumptions (Topi.cc:797-938, 715-769):
// a literal false at level 0 => UNSAT; true at level 0 => dropped; duplicate => dropped; both
// polarities => UNSAT; then one decision level + one BCP per still-unassigned literal, and a cube
// literal found falsified => UNSAT.
//
// usage: ref_bcp <instance(.cnf text | .bcnf)> --out <file.bres>
//                [--probe-all | --cubes <file.cubes>] [--shard i n] [--stride k] [--flags-only] [--repeat r]
//        prints one JSON line with counts and the probe-loop wall time.

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "Topi.hpp"
#include "bcp_io.h"

using namespace std;

struct Harness : public Topor::CTopi<int32_t, uint32_t, false>
{
    using Base = Topor::CTopi<int32_t, uint32_t, false>;
    using TULit = uint32_t;
    using TUVar = uint32_t;

    vector<int32_t> i2e; // internal var -> external var (non-aliased owner)
    int32_t maxUserVar = 0;

    bool Load(const bcp_instance& ins)
    {
        vector<int32_t> cls;
        for (uint64_t i = 0; i < ins.nwords; ++i)
        {
            const int32_t w = ins.words[i];
            cls.push_back(w);
            if (w == 0)
            {
                AddUserClause(span<int32_t>(cls.data(), cls.size()));
                cls.clear();
            }
            else
            {
                maxUserVar = max(maxUserVar, w < 0 ? -w : w);
            }
        }
        return !IsUnrecoverable();
    }

    inline bool Mapped(int32_t v) const { return (size_t)v < m_E2ILitMap.cap() && m_E2ILitMap[v] != 0; }
    inline TULit ToInt(int32_t l) const { return l > 0 ? m_E2ILitMap[l] : Negate(m_E2ILitMap[-l]); }

    void BuildI2E()
    {
        i2e.assign((size_t)GetNextVar() + 1, 0);
        for (int32_t v = 1; v <= maxUserVar; ++v)
        {
            if (!Mapped(v)) continue;
            const TULit il = m_E2ILitMap[v];
            const TUVar iv = GetVar(il);
            // aliased external variables (Topi.cc:559-572) share an internal var with an older
            // owner; only level-0 vars can be aliased, and those never show up in an implied set
            if (iv < i2e.size() && i2e[iv] == 0) i2e[iv] = IsNeg(il) ? -v : v;
        }
    }

    inline int32_t ToExt(TULit il) const
    {
        const int32_t e = i2e[GetVar(il)];
        return IsNeg(il) ? -e : e;
    }

    // returns false iff contradictory at level 0
    bool Level0(size_t maxDepth)
    {
        ReserveVarAndLitData(maxDepth);
        if (IsUnrecoverable()) return false;
        auto ci = BCP();
        if (ci.IsContradiction() || IsUnrecoverable()) return false;
        return true;
    }

    inline bool IsL0(TULit l) const { return IsAssigned(l) && GetAssignedDecLevel(l) == 0; }

    void CollectLevel0(vector<int32_t>& out) const
    {
        for (int32_t v = 1; v <= maxUserVar; ++v)
        {
            if (!Mapped(v)) continue;
            const TULit il = m_E2ILitMap[v];
            if (IsL0(il)) out.push_back(IsSatisfied(il) ? v : -v);
        }
    }

    // one cube; returns flag, appends the implied set (external lits, sorted by |var|)
    uint8_t RunCube(const int32_t* lits, size_t n, vector<int32_t>& implied, bool collect, vector<TULit>& tmp)
    {
        tmp.clear();
        for (size_t i = 0; i < n; ++i)
        {
            const int32_t l = lits[i];
            if (l == 0) break;
            const int32_t v = l < 0 ? -l : l;
            if (!Mapped(v)) return BCP_FLAG_UNMAPPED;
            tmp.push_back(ToInt(l));
        }
        bool anyLive = false;
        for (TULit il : tmp)
        {
            if (IsL0(il))
            {
                if (IsFalsified(il)) return BCP_FLAG_FALSE_L0;
            }
            else anyLive = true;
        }
        if (!anyLive) return BCP_FLAG_TRIVIAL_L0;

        uint8_t flag = BCP_FLAG_OK;
        for (TULit il : tmp)
        {
            if (IsAssigned(il))
            {
                if (IsFalsified(il)) { flag = BCP_FLAG_CONFLICT; break; }
                continue; // true at level 0, duplicate, or already implied
            }
            NewDecLevel();
            Assign(il, BadClsInd, BadULit, m_DecLevel);
            auto ci = BCP();
            if (ci.IsContradiction()) { flag = BCP_FLAG_CONFLICT; break; }
        }
        if (flag == BCP_FLAG_OK && collect)
        {
            const size_t start = implied.size();
            for (TUVar v = m_TrailEnd; v != BadUVar && m_VarInfo[v].m_DecLevel != 0; v = m_VarInfo[v].m_TrailPrev)
            {
                implied.push_back(ToExt(GetAssignedLitForVar(v)));
            }
            sort(implied.begin() + start, implied.end(), [](int32_t a, int32_t b) { return abs(a) < abs(b); });
        }
        Backtrack(0);
        return flag;
    }

    uint64_t Implications() const { return m_Stat.m_Implications; }
    uint64_t Assignments() const { return m_Stat.m_Assignments; }
    uint64_t DelayedTriggers() const { return m_Stat.m_DelayedImplicationsTrigerring; }
};

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <instance> --out f [--probe-all|--cubes f] [--shard i n] [--stride k] [--flags-only]\n", argv[0]); return 2; }
    const char* inst = argv[1];
    const char* out = nullptr; const char* cubesPath = nullptr;
    uint64_t shardI = 0, shardN = 1, stride = 1, repeat = 1; bool flagsOnly = false;
    for (int i = 2; i < argc; ++i)
    {
        string a = argv[i];
        if (a == "--out" && i + 1 < argc) out = argv[++i];
        else if (a == "--cubes" && i + 1 < argc) cubesPath = argv[++i];
        else if (a == "--probe-all") cubesPath = nullptr;
        else if (a == "--shard" && i + 2 < argc) { shardI = strtoull(argv[++i], 0, 10); shardN = strtoull(argv[++i], 0, 10); }
        else if (a == "--stride" && i + 1 < argc) stride = strtoull(argv[++i], 0, 10);
        else if (a == "--flags-only") flagsOnly = true;
        else if (a == "--repeat" && i + 1 < argc) repeat = strtoull(argv[++i], 0, 10);
        else { fprintf(stderr, "bad arg %s\n", argv[i]); return 2; }
    }

    bcp_instance ins;
    if (bcp_read_instance(inst, &ins) != 0) { fprintf(stderr, "cannot read %s\n", inst); return 3; }

    Harness h;
    bool ok = h.Load(ins);
    free(ins.words);

    bcp_cubes cubes; memset(&cubes, 0, sizeof(cubes));
    size_t maxDepth = 1;
    uint64_t n = 0;
    if (cubesPath)
    {
        if (bcp_read_cubes(cubesPath, &cubes) != 0) { fprintf(stderr, "cannot read %s\n", cubesPath); return 3; }
        n = cubes.ncubes;
        for (uint64_t i = 0; i < n; ++i) maxDepth = max(maxDepth, (size_t)(cubes.off[i + 1] - cubes.off[i]));
    }
    else n = 2ull * (uint64_t)h.maxUserVar;

    ok = ok && h.Level0(maxDepth + 1);
    h.BuildI2E();

    vector<uint8_t> flag(n, 255);
    vector<uint32_t> props(n, 0);
    vector<uint64_t> off(n + 1, 0);
    vector<int32_t> implied, level0;
    vector<uint32_t> tmp;
    uint64_t ran = 0, conflicts = 0;
    const uint64_t impl0 = h.Implications(), asg0 = h.Assignments();
    double seconds = 0;

    string secondsList;
    if (ok)
    {
        h.CollectLevel0(level0);
        const uint64_t lo = n * shardI / shardN, hi = n * (shardI + 1) / shardN;
        // --repeat R: R timed passes over the same probes in one process (Backtrack(0) restores the state);
        // the results written are those of the last pass, `seconds` is the last pass, all passes are listed
        for (uint64_t rep = 0; rep < repeat; ++rep)
        {
            implied.clear(); ran = 0; conflicts = 0;
            auto t0 = chrono::steady_clock::now();
            for (uint64_t i = 0; i < n; ++i)
            {
                off[i] = implied.size();
                if (i < lo || i >= hi || (i - lo) % stride != 0) continue;
                const uint64_t before = h.Implications();
                if (cubesPath)
                {
                    flag[i] = h.RunCube(cubes.lits + cubes.off[i], cubes.off[i + 1] - cubes.off[i], implied, !flagsOnly, tmp);
                }
                else
                {
                    const int32_t v = (int32_t)(i / 2 + 1);
                    const int32_t l = (i & 1) ? -v : v;
                    flag[i] = h.RunCube(&l, 1, implied, !flagsOnly, tmp);
                }
                props[i] = (uint32_t)(h.Implications() - before);
                ++ran;
                conflicts += flag[i] == BCP_FLAG_CONFLICT;
            }
            off[n] = implied.size();
            seconds = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
            secondsList += (rep ? ", " : "") + to_string(seconds);
        }
    }

    const uint64_t implications = h.Implications() - impl0, assignments = h.Assignments() - asg0;
    if (out)
    {
        if (bcp_write_results(out, n, ok ? 0 : 1, flag.data(), props.data(), off.data(), implied.data(), level0.data(), level0.size(), implications, assignments, seconds) != 0)
        {
            fprintf(stderr, "cannot write %s\n", out); return 4;
        }
    }
    printf("{\"status\": %d, \"n\": %llu, \"ran\": %llu, \"conflicts\": %llu, \"implications\": %llu, \"assignments\": %llu, \"implied_lits\": %llu, \"level0\": %llu, \"delayed_triggers\": %llu, \"seconds\": %.6f, \"seconds_list\": [%s]}\n",
        ok ? 0 : 1, (unsigned long long)n, (unsigned long long)ran, (unsigned long long)conflicts, (unsigned long long)implications, (unsigned long long)assignments,
        (unsigned long long)implied.size(), (unsigned long long)level0.size(), (unsigned long long)h.DelayedTriggers(), seconds, secondsList.c_str());
    return 0;
}
