// bbcp_db.h -- host side of the clause database: intake/normalisation result-equivalent to the reference's
// AddUserClause (Topi.cc:377-653, CHandleNewCls Topi.hpp:3110-3185), level-0 simplification, and the
// flattening of all clauses into read-only CSR rows + a clause arena (replaces the mutable watch arena of
// TopiWL.cc and the clause buffer written by AddClsToBufferAndWatch, TopiConflictAnalysis.cc:14-113).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace bbcp {

struct RowHdr { uint32_t start, nbin, ntern, nlong; };   // == uint4 on the device
struct Slot { uint32_t x, y; };                           // == uint2 on the device

struct HostDb {
    // ---- intake (raw user clauses, external literals, 0-terminated)
    std::vector<int32_t> raw;
    std::vector<int32_t> pending;     // bbcp_add accumulates one clause here
    int32_t max_var = 0;

    // ---- after normalise()
    bool contradictory = false;
    std::vector<uint8_t> varinfo;     // [max_var+1] bits 0-1: level-0 value (1 true, 2 false), bit 2: mapped
    std::vector<uint32_t> cls;        // kept clauses of size >= 2, internal literals, [len, lits...]*
    uint64_t ncls = 0;
    bool intake_by_chunk = false;     // the last normalise() took the order-free multi-threaded path
    std::vector<uint32_t> units;      // internal literals assigned directly by unit clauses, in order

    // ---- after build_rows()
    std::vector<RowHdr> hdr;          // [2*(max_var+1)]
    std::vector<Slot> slots;
    std::vector<uint32_t> arena;      // 4-word aligned clauses [size, lits..., 0 pad]
    uint64_t n_bin = 0, n_tern = 0, n_long = 0, n_long_lits = 0;

    // ---- after build_order(): the internal literals 2..2*(max_var+1)-1 sorted by their height in the binary
    // implication graph (children first). Empty when the database has no binary clause.
    std::vector<uint32_t> order;
    uint32_t max_height = 0;

    void add_words(const int32_t *w, uint64_t n);
    // AddUserClause semantics over the whole raw stream; fills varinfo (mapped + direct units), cls, units
    void normalise();
    void normalise_serial();     // the defining pass, in stream order
    bool normalise_parallel();   // the same result by clause chunk when no clause is (or becomes) a unit; false = not applicable, nothing decided
    // drop clauses satisfied / strip literals false under varinfo's level-0 values (the
    // level-0-simplified DB); returns false if a clause becomes empty (cannot happen after a level-0 fixpoint)
    bool simplify_level0();
    // flatten cls into hdr/slots/arena. Returns an error string ("" = ok)
    std::string build_rows();
    // Probe schedule for closure reuse (bbcp_propagate.cuh, expand_closures): p -> q is an edge when the binary clause
    // (not-p or q) exists, so closure(q) is a subset of closure(p). Strongly connected components (equivalent literals)
    // are found with an iterative Tarjan pass, which completes them in reverse topological order, so the height of a
    // component (1 + the largest height among its successors) is known the moment it completes. Literals are then
    // counting-sorted by height. Needs build_rows().
    void build_order();
    uint64_t n_mapped() const;
    uint64_t n_level0() const;
};

}  // namespace bbcp
