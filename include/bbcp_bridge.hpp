// bbcp_bridge.hpp -- live-solver bridge (survey "next" row N3): run the GPU engine beside an UNMODIFIED
// Topor-style CPU solver. The bridge shadows AddClause into the GPU engine, runs a GPU probing pass at the
// instants the reference's own inprocessing hook allows -- between Solve calls (InprocessIfRequired is only
// entered at assumption level, TopiInprocess.cc:73-76) -- and feeds what it learns back through the solver's
// public API only:
//   * failed-literal units (and the necessary literals of bbcp_probe_products) as unit clauses: AddClause({l});
//     a failed literal's negation is a RUP consequence of the clause set, so a DRAT trace stays checkable;
//   * during a running Solve the same units are offered through SetParallelData's GetNextUnitClause callback,
//     which the reference polls at every restart (Topi.cc:1306-1345);
//   * a cube batch is filtered by unit-propagation conflict on the GPU before the survivors go to CPU
//     Solve(cube) one by one (cube-and-conquer entry).
//   * learnt clauses can flow the other way: IngestLearntClauses() registers the reference's SetCbNewLearntCls callback
//     (Topi.hpp:120, emitted at TopiConflictAnalysis.cc:1176-1183); what the CPU solver learns is added to the GPU
//     engine's database before its next pass (a learnt clause is implied by the clause set, so it only strengthens unit
//     propagation), and every `everyConflicts`-th learnt clause can trigger a GPU pass INSIDE the running Solve -- the
//     reference's own inprocessing runs on a conflict-count schedule (TopiInprocess.cc:65-89) -- whose units reach the
//     solver through GetNextUnitClause at its next restart;
//   * SetDratLog(): every unit handed to the CPU solver is also written to a DRAT text stream, in derivation order
//     (bbcp_unit_log), so the reference's drat-trim flow still verifies: a failed literal's negation is RUP.
// TSolver is any class with the CTopor surface used below (Topor.hpp:26-121): AddClause(std::span<int32_t>),
// Solve(std::span<int32_t>), SetParallelData(...). The header does not include the reference: the reference is
// bound where the template is instantiated (oracle/bridge_demo.cc does that for the tests).
#ifndef BBCP_BRIDGE_HPP
#define BBCP_BRIDGE_HPP

#include <cstdint>
#include <deque>
#include <ostream>
#include <span>
#include <vector>

#include "bbcp_topor.hpp"

namespace bbcp {

template <class TSolver>
class CToporWithGpuProbing {
public:
    explicit CToporWithGpuProbing(int device = -1) : m_Gpu(0, device)
    {
        // units found by the GPU while a Solve is running are picked up by the reference at its restarts
        m_Cpu.SetParallelData(0, [](unsigned, int) {}, [this](unsigned, bool) {
            if (m_Pending.empty()) return 0;
            const int l = m_Pending.front();
            m_Pending.pop_front();
            return l;
        });
    }

    TSolver &Cpu() { return m_Cpu; }
    CBatchedTopor &Gpu() { return m_Gpu; }

    // Learnt clauses of the CPU solver (at most maxLen literals) join the GPU database; with everyConflicts > 0 a GPU pass
    // runs inside the Solve at every everyConflicts-th learnt clause. `cont` is the callback's "keep going" value
    // (Topor::TStopTopor::VAL_CONTINUE) -- passed in because this header does not include the reference.
    template <class TStop>
    void IngestLearntClauses(TStop cont, size_t maxLen = 8, uint64_t everyConflicts = 0)
    {
        m_Cpu.SetCbNewLearntCls([this, cont, maxLen, everyConflicts](const std::span<int32_t> c) {
            ++m_Conflicts;
            if (c.size() <= maxLen) { m_Learnt.emplace_back(c.begin(), c.end()); ++m_LearntTaken; }
            if (everyConflicts && m_Conflicts % everyConflicts == 0) InSolvePass();
            return cont;
        });
    }

    // DRAT text lines ("l 0") for every unit given to the CPU solver. Necessary literals (implied by both polarities of a
    // variable) are not RUP, so the products are left out while a log is attached.
    void SetDratLog(std::ostream *os) { m_Drat = os; }

    void AddClause(std::vector<int32_t> c)
    {
        m_Gpu.AddClause(c);
        m_Cpu.AddClause(std::span<int32_t>(c));
        m_Stale = true;
    }

    // One GPU inprocessing pass: failed-literal probing to a fixpoint, optionally the necessary literals of the last
    // pass too. Returns the number of new unit clauses handed to the CPU solver; -1 if the GPU refuted the instance
    // (the CPU solver has been given the empty-clause equivalent: both polarities of a variable), -2 on an engine error.
    long GpuProbe(bool withProducts = true)
    {
        if (m_Drat) withProducts = false;
        TakeLearnts();
        if (m_Gpu.Finalize() < 0) return -2;
        std::vector<int32_t> before = Level0();
        const int rc = m_Gpu.ProbeToFixpoint();
        if (rc < 0) return -2;
        m_Stale = false;
        if (rc == BBCP_CONTRADICTORY) {
            m_Cpu.AddClause(std::span<int32_t>(m_False));   // {1}, {-1}: contradictory at level 0
            std::vector<int32_t> t = {-m_False[0]};
            m_Cpu.AddClause(std::span<int32_t>(t));
            return -1;
        }
        std::vector<int32_t> after = Level0();
        if (withProducts) {
            // x implied by +v and by -v is a unit as well
            BatchResult r = m_Gpu.ProbeAll(true, false);
            if (r.rc == BBCP_OK) {
                std::vector<uint32_t> nec(bbcp_bitmap_words(m_Gpu.Handle()));
                uint64_t n = 0;
                if (bbcp_probe_products(m_Gpu.Handle(), nec.data(), &n, nullptr, 0, nullptr) == BBCP_OK && n)
                    for (size_t w = 0; w < nec.size(); ++w)
                        for (uint32_t m = nec[w]; m; m &= m - 1) {
                            const uint32_t il = (uint32_t)(w * 32) + (uint32_t)__builtin_ctz(m);
                            after.push_back((il & 1) ? -(int32_t)(il >> 1) : (int32_t)(il >> 1));
                        }
            }
        }
        long fed = 0;
        std::vector<bool> had;
        for (int32_t l : before) { const size_t v = (size_t)(l < 0 ? -l : l); if (had.size() <= v) had.resize(v + 1, false); had[v] = true; }
        // derivation order first (the failed-literal units, each RUP after the ones before), then what they propagate to
        std::vector<int32_t> ordered(bbcp_unit_log(m_Gpu.Handle(), nullptr, 0));
        if (!ordered.empty()) bbcp_unit_log(m_Gpu.Handle(), ordered.data(), ordered.size());
        ordered.insert(ordered.end(), after.begin(), after.end());
        after.swap(ordered);
        for (int32_t l : after) {
            const size_t v = (size_t)(l < 0 ? -l : l);
            if (had.size() <= v) had.resize(v + 1, false);
            if (had[v]) continue;
            had[v] = true;
            std::vector<int32_t> u = {l};
            if (m_Drat) *m_Drat << l << " 0\n";
            m_Cpu.AddClause(std::span<int32_t>(u));
            m_Gpu.AddClause(u);
            m_Pending.push_back(l);
            ++fed;
        }
        m_UnitsFed += (uint64_t)fed;
        return fed;
    }

    // Solve with the CPU solver; a GPU pass runs first when clauses arrived since the last one
    auto Solve(std::vector<int32_t> assumps = {}, bool probeFirst = true)
    {
        if (probeFirst && m_Stale) GpuProbe();
        return m_Cpu.Solve(std::span<int32_t>(assumps));
    }

    // cube-and-conquer entry: indices of the cubes that unit propagation does not refute
    std::vector<size_t> FilterCubes(const std::vector<int32_t> &lits, const std::vector<uint64_t> &off)
    {
        std::vector<size_t> alive;
        if (m_Gpu.Finalize() < 0) return alive;
        BatchResult r = m_Gpu.PropagateBatch(lits, off, /*implied=*/false);
        for (size_t i = 0; i < r.flag.size(); ++i)
            if (r.flag[i] != BBCP_FLAG_CONFLICT && r.flag[i] != BBCP_FLAG_FALSE_L0) alive.push_back(i);
        return alive;
    }

    uint64_t UnitsFed() const { return m_UnitsFed; }
    uint64_t LearntsTaken() const { return m_LearntTaken; }
    uint64_t InSolvePasses() const { return m_InSolvePasses; }

private:
    void TakeLearnts()
    {
        for (auto &c : m_Learnt) m_Gpu.AddClause(c);
        if (!m_Learnt.empty()) m_Stale = true;
        m_Learnt.clear();
    }
    // A GPU pass while the CPU solver is searching (called from its learnt-clause callback, at whatever decision level it
    // is): the engine works on its own copy of the clauses, so nothing of the solver's state is touched; new level-0
    // literals are queued for GetNextUnitClause (AddClause is not allowed inside Solve).
    void InSolvePass()
    {
        TakeLearnts();
        if (m_Gpu.Finalize() < 0) return;
        std::vector<int32_t> before = Level0();
        const int rc = m_Gpu.ProbeToFixpoint();
        ++m_InSolvePasses;
        if (rc != BBCP_OK) return;     // (a refutation found here is found by the solver's own search as well)
        std::vector<bool> had;
        for (int32_t l : before) { const size_t v = (size_t)(l < 0 ? -l : l); if (had.size() <= v) had.resize(v + 1, false); had[v] = true; }
        for (int32_t l : Level0()) {
            const size_t v = (size_t)(l < 0 ? -l : l);
            if (had.size() <= v) had.resize(v + 1, false);
            if (had[v]) continue;
            m_Pending.push_back(l);
            m_Gpu.AddClause(std::vector<int32_t>{l});
            ++m_UnitsFed;
        }
    }
    std::vector<int32_t> Level0()
    {
        std::vector<int32_t> out(bbcp_level0_lits(m_Gpu.Handle(), nullptr, 0));
        if (!out.empty()) bbcp_level0_lits(m_Gpu.Handle(), out.data(), out.size());
        return out;
    }
    TSolver m_Cpu;
    CBatchedTopor m_Gpu;
    std::deque<int> m_Pending;
    std::vector<int32_t> m_False = {1};
    bool m_Stale = true;
    uint64_t m_UnitsFed = 0, m_Conflicts = 0, m_LearntTaken = 0, m_InSolvePasses = 0;
    std::vector<std::vector<int32_t>> m_Learnt;
    std::ostream *m_Drat = nullptr;
};

}  // namespace bbcp
#endif
