// bbcp_topor.hpp -- C++ host mirror of the reference's public facade for the batched-BCP path.
//
// bbcp::CBatchedTopor keeps the method names, argument meaning, value conventions and error behaviour of
// Topor::CTopor (Topor.hpp:26-121) for the subset of the facade that sits on this path, and adds the batch
// entry points the GPU engine exists for. It is header-only and only forwards to the C ABI of bbcp.h
// (libbbcp.so); no propagation happens on the host. Like the reference facade it never throws: errors
// are sticky and visible through IsError() / GetStatusExplanation() (Topor.hpp:111-114).
//
//   reference (Topor.hpp)                           here
//   CTopor(TLit varsNumHint = 0)            :31      CBatchedTopor(int32_t varsNumHint = 0, int device = -1)
//   void AddClause(span<TLit>)              :38      AddClause(const int32_t*, size_t) / vector / init-list / variadic
//   TToporReturnVal Solve(span<TLit> assumps) :48    Solve(assumps): RET_UNSAT iff UP of the assumptions derives a
//                                                    conflict; RET_SAT iff every variable got assigned; else
//                                                    RET_USER_INTERRUPT (fixpoint reached, search not run: what the
//                                                    reference returns with SetCbStopNow -> VAL_STOP)
//   TToporLitVal GetLitValue(TLit) const    :80      same four values
//   TLit GetLitDecLevel(TLit) const         :88      0 level-0 literal, 1 assigned by the last Solve's cube
//   void Backtrack(TLit)                    :68      Backtrack(0) forgets the last Solve's assignment
//   uint64_t GetPropagations() const        :108     reference convention (TopiBcp.cc:193-198)
//   bool IsError() / GetStatusExplanation() :111-114 same
//   TLit GetMaxUserVar() const              :94      same
#ifndef BBCP_TOPOR_HPP
#define BBCP_TOPOR_HPP

#include <array>
#include <cstdint>
#include <initializer_list>
#include <string>
#include <vector>
#if __cplusplus >= 202002L
#include <span>
#endif

#include "bbcp.h"

namespace bbcp {

// ToporExternalTypes.hpp:52-78 (same enumerator order, hence same numeric values)
enum class TToporReturnVal : uint8_t {
    RET_SAT, RET_UNSAT, RET_TIMEOUT_LOCAL, RET_CONFLICT_OUT, RET_MEM_OUT, RET_USER_INTERRUPT, RET_INDEX_TOO_NARROW,
    RET_PARAM_ERROR, RET_ASSUMPTION_REQUIRED_ERROR, RET_TIMEOUT_GLOBAL, RET_DRAT_FILE_PROBLEM, RET_EXOTIC_ERROR
};
// ToporExternalTypes.hpp:81-87
enum class TToporLitVal : uint8_t { VAL_SATISFIED, VAL_UNSATISFIED, VAL_UNASSIGNED, VAL_DONT_CARE };

// results of one batch, fetched into host vectors
struct BatchResult {
    bbcp_batch_info info{};
    std::vector<uint8_t> flag;        // [n] BBCP_FLAG_*
    std::vector<uint64_t> impl_off;   // [n+1]
    std::vector<int32_t> impl_lits;   // external literals, sorted by |var| inside each set
    int rc = BBCP_OK;
};

class CBatchedTopor {
public:
    explicit CBatchedTopor(int32_t varsNumHint = 0, int device = -1) : m_H(bbcp_create(varsNumHint))
    {
        if (m_H && device >= 0) bbcp_set_device(m_H, device);
    }
    ~CBatchedTopor() { bbcp_release(m_H); }
    CBatchedTopor(const CBatchedTopor &) = delete;
    CBatchedTopor &operator=(const CBatchedTopor &) = delete;

    // ---- CTopor-shaped surface
    void AddClause(const int32_t *lits, size_t n) { if (m_H) bbcp_add_clause(m_H, lits, n); }
    void AddClause(const std::vector<int32_t> &c) { AddClause(c.data(), c.size()); }
    void AddClause(std::initializer_list<int32_t> lits) { std::vector<int32_t> v(lits); AddClause(v); }
    template <class... T> void AddClause(int32_t lit1, T... lits) { std::array<int32_t, 1 + sizeof...(T)> v = {lit1, (int32_t)lits...}; AddClause(v.data(), v.size()); }
#if __cplusplus >= 202002L
    void AddClause(std::span<const int32_t> c) { AddClause(c.data(), c.size()); }
#endif

    TToporReturnVal Solve(const int32_t *assumps = nullptr, size_t n = 0)
    {
        if (!m_H) return TToporReturnVal::RET_MEM_OUT;
        for (size_t i = 0; i < n && assumps[i] != 0; ++i) bbcp_assume(m_H, assumps[i]);   // may be 0-ended (Topi.cc:824-827)
        const int r = bbcp_solve(m_H);
        const int st = bbcp_status(m_H);
        if (st == BBCP_ERR_ALLOC) return TToporReturnVal::RET_MEM_OUT;
        if (st == BBCP_ERR_INDEX_TOO_NARROW) return TToporReturnVal::RET_INDEX_TOO_NARROW;
        if (st < 0) return TToporReturnVal::RET_EXOTIC_ERROR;
        return r == 20 ? TToporReturnVal::RET_UNSAT : r == 10 ? TToporReturnVal::RET_SAT : TToporReturnVal::RET_USER_INTERRUPT;
    }
    TToporReturnVal Solve(const std::vector<int32_t> &assumps) { return Solve(assumps.data(), assumps.size()); }

    TToporLitVal GetLitValue(int32_t l) const
    {
        switch (m_H ? bbcp_val(m_H, l) : 2) {
        case 1: return TToporLitVal::VAL_SATISFIED;
        case -1: return TToporLitVal::VAL_UNSATISFIED;
        case 0: return TToporLitVal::VAL_UNASSIGNED;
        default: return TToporLitVal::VAL_DONT_CARE;
        }
    }
    int32_t GetLitDecLevel(int32_t l) const { return m_H ? bbcp_declevel(m_H, l) : -1; }
    void Backtrack(int32_t decLevel) { if (m_H) bbcp_backtrack(m_H, decLevel); }
    uint64_t GetPropagations() const { return m_H ? bbcp_propagations(m_H) : 0; }
    bool IsError() const { return !m_H || bbcp_status(m_H) < 0; }
    std::string GetStatusExplanation() const { return m_H ? bbcp_error_string(m_H) : "allocation of the engine handle failed"; }
    int32_t GetMaxUserVar() const { bbcp_db_info d; return m_H && bbcp_get_db_info(m_H, &d) == BBCP_OK ? d.max_var : 0; }

    // ---- batch surface (what TopiInprocess / a cube-and-conquer front end would call)
    // normalise + level-0 propagation + upload; BBCP_OK, BBCP_CONTRADICTORY, or an error code
    int Finalize() { return m_H ? bbcp_finalize(m_H) : BBCP_ERR_ALLOC; }
    // failed-literal probing of every literal (+1,-1,+2,-2,...)
    BatchResult ProbeAll(bool implied = true, bool fetch = true)
    {
        bbcp_db_info d{};
        if (m_H) bbcp_get_db_info(m_H, &d);
        return ProbeRange(0, 2ull * (uint64_t)d.max_var, implied, fetch);
    }
    BatchResult ProbeRange(uint64_t first, uint64_t last, bool implied = true, bool fetch = true)
    {
        BatchResult r;
        bbcp_opts o{};
        o.flags = implied ? BBCP_OPT_IMPLIED : 0u;
        r.rc = m_H ? bbcp_probe_range(m_H, first, last, &o, &r.info) : BBCP_ERR_ALLOC;
        if (r.rc == BBCP_OK && fetch) Fetch(r, implied);
        return r;
    }
    // cubes in CSR form: cube i = lits[off[i] .. off[i+1])
    BatchResult PropagateBatch(const std::vector<int32_t> &lits, const std::vector<uint64_t> &off, bool implied = true, bool fetch = true)
    {
        BatchResult r;
        bbcp_opts o{};
        o.flags = implied ? BBCP_OPT_IMPLIED : 0u;
        static const int32_t none = 0;
        r.rc = (m_H && !off.empty()) ? bbcp_propagate_batch(m_H, lits.empty() ? &none : lits.data(), off.data(), off.size() - 1, &o, &r.info) : BBCP_ERR_ARG;
        if (r.rc == BBCP_OK && fetch) Fetch(r, implied);
        return r;
    }
    // failed-literal probing with unit feedback to a fixpoint; BBCP_OK or BBCP_CONTRADICTORY
    int ProbeToFixpoint(uint32_t maxRounds = 0, uint32_t *rounds = nullptr, uint64_t *units = nullptr)
    {
        return m_H ? bbcp_probe_to_fixpoint(m_H, maxRounds, rounds, units) : BBCP_ERR_ALLOC;
    }
    // external literals whose depth-1 probe hit a conflict so far (ascending |var|, positive first)
    std::vector<int32_t> FailedLiterals()
    {
        std::vector<int32_t> out;
        if (!m_H) return out;
        std::vector<uint32_t> f(bbcp_bitmap_words(m_H));
        if (bbcp_fetch_bitmaps(m_H, f.data(), nullptr) != BBCP_OK) return out;
        for (size_t w = 0; w < f.size(); ++w)
            for (uint32_t m = f[w]; m; m &= m - 1) {
                const uint32_t il = (uint32_t)(w * 32) + (uint32_t)__builtin_ctz(m);
                out.push_back((il & 1) ? -(int32_t)(il >> 1) : (int32_t)(il >> 1));
            }
        return out;
    }
    bbcp_handle *Handle() { return m_H; }

private:
    void Fetch(BatchResult &r, bool implied)
    {
        r.flag.resize(r.info.n);
        r.rc = bbcp_fetch_flags(m_H, r.flag.data());
        if (r.rc != BBCP_OK || !implied) return;
        r.impl_off.resize(r.info.n + 1);
        r.impl_lits.resize(r.info.n_implied ? r.info.n_implied : 1);
        r.rc = bbcp_fetch_implied(m_H, r.impl_off.data(), r.impl_lits.data());
        r.impl_lits.resize(r.info.n_implied);
    }
    bbcp_handle *m_H;
};

}  // namespace bbcp
#endif
