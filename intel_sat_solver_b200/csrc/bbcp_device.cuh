// bbcp_device.cuh -- device-side views shared by the kernels and the host launcher.
//
// Database layout in HBM (built by bbcp_db.cpp, read-only for every probe):
//   hdr[ilit]   uint4 {start, nbin, ntern, nlong}: the row scanned when internal literal `ilit` becomes
//               FALSE (the reference indexes m_Watches[Negate(p)] the same way, TopiBcp.cc:192).
//   slots[]     8-byte slots. Row = ceil(nbin/2) binary slots {other, other|0}  (Topi.hpp:1952-1962 keeps
//               4-byte binary entries too), then ntern ternary slots {a, b} (the clause is ilit v a v b,
//               held inline: no arena fetch), then nlong long slots {blocker, clause16} -- one per OCCURRENCE
//               of ilit in a clause of >= 4 literals, blocker = the next literal of that clause.
//   arena[]     16-byte units; clause = [size, l0, l1, ...] padded with 0 to a multiple of 4 words.
//   litinfo[l]  per internal literal l, the facts a probe of l needs: true / false at level 0, variable
//               mapped, row of the NEGATION of l has binary clauses / is non-empty (LI_* bits).
// Internal literal = 2*var + neg with var = the external (DIMACS) variable.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bbcp {

struct DbView {
    const uint4 *hdr;
    const uint2 *slots;
    const uint4 *arena;
    const uint8_t *litinfo;     // per internal literal l (as a probe): LI_* bits
    int32_t max_var;
    uint32_t n_slots, n_arena_units;   // extents (bounds checks of debug builds: -DBBCP_CHECKS)
    uint32_t nvars_words16;    // words of a 2-bit-per-var dense assignment: ceil((max_var+1)/16)
    // 1: literals became true / false at level 0 AFTER the rows were built (unit-feedback rounds, bbcp_probe_to_fixpoint): the
    // rows still hold them, so a literal the probe has not assigned is also looked up in litinfo (value_of)
    uint32_t l0_dirty;
};

constexpr uint8_t LI_TRUE_L0 = 1, LI_FALSE_L0 = 2, LI_MAPPED = 4, LI_HAS_BIN = 8, LI_NONEMPTY = 16;
constexpr uint32_t LEN_SELF = 0x80000000u;   // len[] flag: the implied set is the item's own literal (not stored)

enum Counter { C_PROPS_ALL = 0, C_PROPS_REF, C_SLOTS, C_CLAUSE_FETCH, C_BYTES, C_LARGE, C_MID, C_OK, C_CONFLICT, C_SKIPPED, C_ADOPTED, C_ADOPT_LITS, C_INHERITED, C_PROPS_REF_OK, C_ASSIGN, C_ASSIGN_OK, C_BCPS, C_HOT_ROWS, C_N };

struct BatchView {
    // input: cubes (external literals, CSR); or, when cube_off == nullptr, a list of probe literals
    // (cube_lits[i] is item i); or, when both are nullptr, the probe universe starting at `first`
    const int32_t *cube_lits;
    const uint64_t *cube_off;
    // bbcp_probe_lits_to_host with pinned caller buffers: the triage reads the probe list straight out of the caller's
    // (mapped) host memory and leaves the device copy `cube_lits` for the kernels after it; it also writes the flags it
    // decides into the caller's flag buffer. Both nullptr otherwise.
    const int32_t *host_lits;
    uint8_t *host_flag;
    uint64_t first;             // universe index of item 0 (probe mode)
    uint64_t n;                 // items in the batch
    // output
    uint8_t *flag;              // [n]
    uint64_t *raw_off;          // [n] position of the item's implied set in out_lits
    uint32_t *len;              // [n] size of the implied set (| LEN_SELF)
    int32_t *out_lits;          // external literals, sorted by |var| inside each set
    uint64_t out_cap;
    unsigned long long *out_cursor;
    uint32_t *failed, *unit;    // bitmaps over internal literals (probe mode only)
    unsigned long long *counters;
    uint32_t *mid_list;         // items that did not fit the warp-private shared-memory state
    uint32_t *mid_count;
    uint32_t *big_list;         // items that outgrew the second tier
    uint32_t *big_count;
    uint32_t *huge_list;        // items that outgrew the third tier (dense last tier)
    uint32_t *huge_count;
    unsigned int *ticket;       // [4] dynamic schedulers of the four tiers
    uint32_t *deep_list;        // items queued by the triage for propagation
    uint32_t *deep_count;
    uint32_t *sort_list;        // second-tier items whose implied set waits in the pool unsorted (k_sort_sets)
    uint32_t *sort_count;
    uint2 *chunk_list;          // (item, chunk) pairs of giant implied sets, filled by k_scan_gather for k_gather_long
    uint32_t *chunk_count;
    // result compaction (k_scan_gather): canonical CSR offsets + literals
    void *off;                  // [n+1] u64, or u32 with OPT_OFF32
    unsigned long long off_base;   // added to every offset written (chunked batches that share one literal array)
    int32_t *impl;              // implied literals in item order
    unsigned long long *tile_sum;     // [ceil(n / 2048)] sum of the set sizes of each 2048-item tile (k_triage stores, propagate kernels add)
    uint32_t *tile_ticket;            // +1: finished-CTA count of k_scan_gather
    unsigned long long *total;  // off[n]
    unsigned long long *host_report;  // mapped pinned host copy of the counter block, written by k_scan_gather (or nullptr)
    unsigned long long *zero_next;    // the counter block of the NEXT launch sequence, zeroed by k_triage
    uint32_t zero_words;
    uint32_t opts;
    // closure reuse (universe batches only): order[] = every internal literal sorted by its height in the binary
    // implication graph, children first (HostDb::build_order); the first tier takes its probes in that order, and a probe
    // that reaches a literal whose own probe has finished adopts that result (expand_closures). nullptr: triage queue order.
    const uint32_t *order;
    uint32_t order_n;
    uint32_t reuse;             // 1: item i of this batch is the probe of internal literal first + i + 2 and may be looked up
};

constexpr uint32_t OPT_IMPLIED = 1u;
constexpr uint32_t OPT_NO_LEN_PRUNE = 2u;
constexpr uint32_t OPT_OFF32 = 4u;
constexpr uint32_t OPT_COMPACT_SELF = 128u;
constexpr uint32_t OPT_BINARY_ONLY = 256u;   // propagate through the binary clauses only (the binary-implication closure; bbcp_probe_hbr)
constexpr uint8_t FLAG_SELF = 0x80;         // OPT_COMPACT_SELF: flag OK + "the implied set is the probe's own literal" (not stored)
constexpr uint8_t FLAG_OK = 0, FLAG_CONFLICT = 1, FLAG_TRIVIAL_L0 = 2, FLAG_FALSE_L0 = 3, FLAG_UNMAPPED = 4;
constexpr uint8_t FLAG_OVERFLOW = 250;      // internal: outgrew the first tier, queued for the second (global hash state)
constexpr uint8_t FLAG_OVERFLOW2 = 249;     // internal: outgrew the second tier, queued for the third
constexpr uint8_t FLAG_OVERFLOW3 = 248;     // internal: outgrew the third tier, queued for the last (dense state, any size)
constexpr uint8_t FLAG_OVERFLOW_LAST = 248;
constexpr uint8_t FLAG_QUEUED = 255;        // internal: written by the triage for items that wait for the first tier
// Trail entries carry a 2-bit mode above the 30-bit literal (variables stop below 2^29):
//   0            an implication found by a row scan: its whole row is walked when it is dequeued;
//   TRAIL_CLOSED the literal came in with a finished closure, so every binary clause of its row is already satisfied:
//                only the ternary / long part of the row is walked;
//   TRAIL_SKIP   it came in with a finished closure that was LARGER than the state it was entered into: its row is not
//                walked at all -- instead the literals that were there before get a second walk (TRAIL_RESCAN);
//   TRAIL_RESCAN a second walk (ternary / long part) of a literal that is already on the trail: NOT a member of the
//                implied set, not looked up for adoption.
// See expand_closures in bbcp_propagate.cuh for why that is complete.
constexpr uint32_t TRAIL_LIT = 0x3fffffffu, TRAIL_RESCAN = 0x40000000u, TRAIL_CLOSED = 0x80000000u, TRAIL_SKIP = 0xc0000000u;
constexpr uint8_t FLAG_OUT_OVERFLOW = 251;  // internal: implied-literal buffer full, re-run in a later launch

// last tier: the trail holds every variable once plus the second-walk entries (TRAIL_RESCAN: each SKIP adoption at least
// quadruples the member count, so they stay below a third of the members)
__host__ __device__ inline uint32_t big_trail_words(uint32_t max_var) { return max_var + 1u + (max_var + 1u) / 2u + 64u; }

struct LargeState {
    uint32_t *bits;    // [slabs][nvars_words16] 2 bits per var, all zero between probes
    uint32_t *trail;   // [slabs][big_trail_words(max_var)]
};

void launch_triage(const BatchView &b, const DbView &db, int sm_count, cudaStream_t s);
void launch_small(const DbView &db, const BatchView &b, uint32_t trail_cap, int width, int blocks, cudaStream_t s, bool pdl);
void launch_mid(const DbView &db, const BatchView &b, uint32_t *slab, uint32_t cap, int width, int blocks, cudaStream_t s);
void launch_sort_sets(const BatchView &b, uint32_t cap, int sm_count, cudaStream_t s);
size_t mid_slab_words_per_group(uint32_t cap);
int mid_blocks_per_sm(int width);
void launch_big(const DbView &db, const BatchView &b, const LargeState &ls, int blocks, bool from_huge, cudaStream_t s);
// CTA-per-probe tiers with the hash in shared memory: cls 0 = second tier (reads mid_list), 1 = third tier (reads big_list)
struct BlockTier { int threads; uint32_t hs, cap; };   // CTA size, hash entries (shared memory), trail capacity
BlockTier block_tier_shape(int cls, uint32_t cap);
int block_tier_blocks_per_sm(const BlockTier &t);
void launch_block_tier(const DbView &db, const BatchView &b, const BlockTier &t, int cls, uint32_t *trail_slab, int blocks, cudaStream_t s);
constexpr uint32_t SG_TILE = 2048;   // items per step of the compaction kernel
size_t small_smem_bytes(uint32_t trail_cap, int width);
int small_blocks_per_sm(uint32_t trail_cap, int width);
struct MergeArgs;
void set_debug_buffer(unsigned long long *mapped);
void read_phase_ticks(unsigned long long *out);
void build_order_device(const DbView &db, uint32_t nl, uint8_t *d_height, uint32_t *d_order, int sm_count, cudaStream_t s);
bool launch_scan_gather(const BatchView &b, int pass, int max_blocks, bool giant_sets_possible, const MergeArgs &mg, cudaStream_t s, bool pdl);
int scan_gather_max_blocks(int sm_count);
// hyper-binary resolvents: x in full[i] \ bin[i] for every item i that reached a fixpoint in both passes
void launch_hbr(const uint8_t *flag_full, const uint64_t *off_full, const int32_t *impl_full, const uint8_t *flag_bin, const uint64_t *off_bin, const int32_t *impl_bin,
                uint64_t n, uint64_t first, int32_t *pairs, unsigned long long cap, unsigned long long *count, int sm_count, cudaStream_t s);
void launch_products(bool off32, const uint8_t *flag, const void *off, const int32_t *impl, uint64_t npairs, uint64_t var_base, uint32_t *necessary, int32_t *equiv,
                     unsigned long long equiv_cap, unsigned long long *counts, int sm_count, cudaStream_t s);
struct MergeArgs {
    uint32_t *const *peer_base;   // [world] every rank's [failed | unit | control] block (own entry = own memory)
    int rank, world;              // world <= 1: nothing to merge
    uint64_t stride;              // words between the failed and the unit bitmap
    uint64_t wp;                  // bitmap words owned by each rank (multiple of 4): rank r owns [r*wp, (r+1)*wp); 0 = rendezvous only
    uint32_t epoch;
    uint32_t timeout_ms;          // in-kernel wait limit for a peer's flag
    uint32_t dirty_before;        // this rank's slice may hold failed literals from earlier batches (no empty-slice hint)
};
constexpr int MERGE_CTRL_WORDS = 256;      // control words behind the two bitmaps, see merge_signal in bbcp_kernels.cu
constexpr int MERGE_CTRL_FINISHED = 64, MERGE_CTRL_ERROR = 65, MERGE_CTRL_DONE = 128, MERGE_CTRL_EMPTY = 192;
void launch_check_fixpoints(const DbView &db, bool off32, const uint8_t *flag, const void *off, const int32_t *impl, uint64_t first, uint64_t count, uint64_t stride,
                            uint32_t *scratch, int blocks, unsigned long long *out, cudaStream_t s);
// marks the n external literals of `lits` true at level 0 in litinfo (and their negations false)
void launch_mark_level0(uint8_t *litinfo, const int32_t *lits, uint64_t n, cudaStream_t s);
void launch_merge(const MergeArgs &m, int sm_count, cudaStream_t s);
void launch_merge_drain(const MergeArgs &m, cudaStream_t s);
constexpr int SMALL_THREADS = 128;  // first tier: SMALL_THREADS / W probes in flight per CTA (W lanes per probe)
constexpr int MID2_THREADS = 256;   // second tier: MID2_THREADS / W probes per CTA, state in a global slab
constexpr int BIG_THREADS = 256;   // third tier: 8 warps share one probe, state in a global slab
// (Other shapes were tried on the full-size Tseitin workload at the end of round 2: 128 threads x 8 CTAs per SM is 10 % slower;
// 128 x 10 and 64 x 16 did not finish their pass and the cause was not found -- re-validate with tools/bench_workloads.py c3
// before changing this or the launch bound of k_big.)

}  // namespace bbcp
