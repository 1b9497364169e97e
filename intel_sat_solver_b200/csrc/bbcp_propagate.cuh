// bbcp_propagate.cuh -- the propagation core shared by every tier: per-probe state, the row scan, the commit
// of implied literals. Included by bbcp_kernels.cu only.
//
// A probe is owned by a group of W lanes (Grp<W>); every collective carries the group's lane mask. The engine
// instantiates W = 32 (one warp per probe) only. W = 8 / 16 (four / two probes per warp, diverging freely) was
// built to put more probes in flight on structured instances, whose rows keep ~10 % of a warp's lanes busy, and
// measured on the Tseitin/BMC shape on a B200: 2.3x SLOWER at W = 8, 1.5x slower at W = 16 -- divergent groups of one
// warp are issued one after the other, so the instruction stream per warp quadruples while the latency it was
// meant to hide does not overlap any better. (The "rare fault under divergence" noted then was found in round 2: lanes
// of one warp read a CTA-wide abort latch at different moments and the warp split; see BlockTail::aborted.) Rejected;
// the template parameter stays because the code reads no worse for it.
#pragma once
#include "bbcp_device.cuh"

// Diagnostic build (-DBBCP_CHECKS, `make checks`): bounds checks and a loop watchdog. A failing check records its source
// line, CTA and thread in a mapped host buffer (device printf output is lost when the context dies) and traps; the
// host prints the record with the CUDA error (bbcp_capi.cu, cuda_ok). Compiled out of the product build.
#ifdef BBCP_CHECKS
#include <cstdio>
namespace bbcp { __device__ unsigned long long *g_dbg = nullptr; }   // (this header is included by bbcp_kernels.cu only)
#define BBCP_FAIL(v0, v1) do { if (bbcp::g_dbg && bbcp::g_dbg[0] == 0ull) { bbcp::g_dbg[0] = (unsigned long long)__LINE__; \
        bbcp::g_dbg[1] = blockIdx.x; bbcp::g_dbg[2] = threadIdx.x; bbcp::g_dbg[3] = (unsigned long long)(v0); bbcp::g_dbg[4] = (unsigned long long)(v1); \
        __threadfence_system(); } __trap(); } while (0)
#define BBCP_CHECK(cond, ...) do { if (!(cond)) { printf("BBCP_CHECK failed: " #cond " | " __VA_ARGS__); BBCP_FAIL(0, 0); } } while (0)
#define BBCP_MARK(id) do { if (bbcp::g_dbg && blockIdx.x < 480 && blockDim.x == 256 && (threadIdx.x & 31) == 0) bbcp::g_dbg[256 + blockIdx.x * 8 + (threadIdx.x >> 5)] = ((unsigned long long)(id) << 32) | __LINE__; } while (0)
// phase timers of the CTA tiers (thread 0 of a CTA, clock64 ticks summed over all CTAs; read with bbcp_debug_phase_ticks)
namespace bbcp { __device__ unsigned long long g_phase[8]; }
#define BBCP_PHASE_DECL() long long bbcp_ph_t = clock64()
#define BBCP_PHASE_DECL2() bbcp_ph_t = clock64()
#define BBCP_PHASE(i) do { if (threadIdx.x == 0) { const long long bbcp_ph_n = clock64(); atomicAdd(&bbcp::g_phase[i], (unsigned long long)(bbcp_ph_n - bbcp_ph_t)); bbcp_ph_t = bbcp_ph_n; } } while (0)
#define BBCP_WATCH_BEGIN() const unsigned long long bbcp_watch_t0 = bbcp::watch_ns()
#define BBCP_WATCH(v0, v1) do { if (bbcp::watch_ns() - bbcp_watch_t0 > 4000000000ull) BBCP_FAIL(v0, v1); } while (0)
namespace bbcp { __device__ __forceinline__ unsigned long long watch_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; } }
#else
#define BBCP_CHECK(cond, ...) do { } while (0)
#define BBCP_MARK(id) do { } while (0)
#define BBCP_PHASE_DECL() do { } while (0)
#define BBCP_PHASE_DECL2() do { } while (0)
#define BBCP_PHASE(i) do { } while (0)
#define BBCP_WATCH_BEGIN() do { } while (0)
#define BBCP_WATCH(v0, v1) do { } while (0)
#endif

namespace bbcp {
namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int32_t to_ext(uint32_t il) { return (il & 1u) ? -(int32_t)(il >> 1) : (int32_t)(il >> 1); }

template <int W>
struct Grp {
    static_assert(W == 8 || W == 16 || W == 32, "group width");
    static constexpr int width = W;
    __device__ __forceinline__ static int lane() { return threadIdx.x & (W - 1); }
    __device__ __forceinline__ static int shift() { return (threadIdx.x & 31) & ~(W - 1); }   // first warp lane of the group
    __device__ __forceinline__ static unsigned mask()
    {
        if (W == 32) return FULL;
        return ((1u << (W & 31)) - 1u) << shift();
    }
    __device__ __forceinline__ static unsigned rel(unsigned warp_bits)
    {
        if (W == 32) return warp_bits;
        return (warp_bits >> shift()) & ((1u << (W & 31)) - 1u);
    }
    template <class T> __device__ __forceinline__ static T shfl(T v, int src) { return __shfl_sync(mask(), v, src, W); }
    template <class T> __device__ __forceinline__ static T shfl_up(T v, int d) { return __shfl_up_sync(mask(), v, d, W); }
    template <class T> __device__ __forceinline__ static T shfl_xor(T v, int d) { return __shfl_xor_sync(mask(), v, d, W); }
    __device__ __forceinline__ static unsigned ballot(bool p) { return rel(__ballot_sync(mask(), p)); }
    __device__ __forceinline__ static bool any(bool p) { return __any_sync(mask(), p) != 0; }
    __device__ __forceinline__ static unsigned match_any(uint32_t v) { return rel(__match_any_sync(mask(), v)); }
    __device__ __forceinline__ static void sync() { __syncwarp(mask()); }
    __device__ __forceinline__ static unsigned lt() { return (1u << lane()) - 1u; }   // group lanes below mine
};

// ---------------------------------------------------------------- per-probe state, first tier (shared memory)
// Bulk-async staging of long rows (the "TMA stages hot watch blocks" of the north star), measured in round 2 and OFF by
// default (BBCP_TMA_ROWS=1 turns it on; profiles/r02): when a step walks one long row, its slots come into shared memory
// STAGE_SLOTS at a time with ONE cp.async.bulk (UBLKCP in SASS) completing on an mbarrier, instead of eight 8-byte LDGs
// per lane. The walk is bound by the two hash look-ups per slot, not by fetching the slots, and the 2 KB per warp cost a
// resident CTA per SM.
constexpr uint32_t STAGE_SLOTS = 256, STAGE_MIN = 512;
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar))); }
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src), "r"(bytes),
                 "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

struct SmemState {
    uint32_t *hash;    // [1 << log_hs] true literals, 0 = empty
    uint32_t *trail;   // [cap]
    uint16_t *tslot;   // [cap] hash slot of each trail entry (for the O(|trail|) clean-up)
    uint32_t log_hs, cap;
    uint2 *stage = nullptr;               // [STAGE_SLOTS + 2] staging buffer of this warp (16-byte aligned), or nullptr
    unsigned long long *stage_bar = nullptr;
    uint32_t stage_phase = 0;

    __device__ __forceinline__ uint32_t h(uint32_t var) const { return (var * 0x9E3779B1u) >> (32 - log_hs); }
    // 0 unassigned, 1 true, 2 false
    __device__ __forceinline__ int lookup(uint32_t lit) const
    {
        const uint32_t var = lit >> 1, mask = (1u << log_hs) - 1u;
        uint32_t s = h(var);
        for (;;) {
            const uint32_t k = *(volatile uint32_t *)&hash[s];
            if (k == 0) return 0;
            if ((k >> 1) == var) return k == lit ? 1 : 2;
            s = (s + 1) & mask;
        }
    }
    // 0 newly assigned, 1 already true, 2 already false (conflict)
    __device__ __forceinline__ int insert(uint32_t lit, uint32_t &slot)
    {
        const uint32_t var = lit >> 1, mask = (1u << log_hs) - 1u;
        uint32_t s = h(var);
        for (;;) {
            const uint32_t old = atomicCAS(&hash[s], 0u, lit);
            if (old == 0) { slot = s; return 0; }
            if ((old >> 1) == var) return old == lit ? 1 : 2;
            s = (s + 1) & mask;
        }
    }
    __device__ __forceinline__ uint32_t tget(uint32_t i) const { return *(volatile uint32_t *)&trail[i]; }
    __device__ __forceinline__ void tset(uint32_t i, uint32_t lit, uint32_t slot) { trail[i] = lit; tslot[i] = (uint16_t)slot; }
    template <class G> __device__ __forceinline__ void cleanup(uint32_t tail)
    {
        for (uint32_t i = G::lane(); i < tail; i += G::width) { const uint32_t s = tslot[i]; if (s != 0xffffu) hash[s] = 0; }   // (0xffff: a second-walk entry, it owns no slot)
        G::sync();
    }
    // after an overflow some table entries have no trail entry: wipe everything
    template <class G> __device__ __forceinline__ void wipe()
    {
        for (uint32_t i = G::lane(); i < (1u << log_hs); i += G::width) hash[i] = 0;
        G::sync();
    }
    template <class G> __device__ __forceinline__ void maybe_grow(uint32_t) {}
};

// ---------------------------------------------------------------- per-probe state, second tier (global hash)
// Same open-addressed set + trail, but in a per-group slab in global memory (L2-resident while hot): a probe
// with a trail of thousands of literals still occupies one group of lanes, so hundreds of them run per SM,
// where a shared-memory state of that size would allow four.
// The table GROWS with the probe: it starts at 2^GHASH_LOG_MIN entries and is rebuilt four times larger from the trail
// whenever it is half full, up to 2^log_max (= twice the trail capacity). A fixed full-size table made every look-up a
// DRAM sector once the slabs of all resident warps (thousands x 96 KB) outgrew the 126 MB L2 -- ncu on the Tseitin
// shape: 18 % L2 hit rate, 323 GB of DRAM reads for 1.1 G implied literals; most second-tier probes hold a few
// thousand literals and now touch a few tens of KB. Invariant: the slab is all zero beyond the live table.
constexpr uint32_t GHASH_LOG_MIN = 11;
struct GhashState {
    uint32_t *hash;   // [1 << log_hs] true literals, 0 = empty
    uint32_t *trail;  // [cap]
    uint32_t log_hs, cap;   // log_hs: the live size
    uint32_t log_max;

    __device__ __forceinline__ uint32_t h(uint32_t var) const { return (var * 0x9E3779B1u) >> (32 - log_hs); }
    __device__ __forceinline__ int lookup(uint32_t lit) const
    {
        const uint32_t var = lit >> 1, mask = (1u << log_hs) - 1u;
        uint32_t s = h(var);
        for (;;) {
            const uint32_t k = __ldcg(&hash[s]);
            if (k == 0) return 0;
            if ((k >> 1) == var) return k == lit ? 1 : 2;
            s = (s + 1) & mask;
        }
    }
    __device__ __forceinline__ int insert(uint32_t lit, uint32_t &slot)
    {
        const uint32_t var = lit >> 1, mask = (1u << log_hs) - 1u;
        uint32_t s = h(var);
        slot = 0;
        for (;;) {
            const uint32_t old = atomicCAS(&hash[s], 0u, lit);
            if (old == 0) return 0;
            if ((old >> 1) == var) return old == lit ? 1 : 2;
            s = (s + 1) & mask;
        }
    }
    __device__ __forceinline__ uint32_t tget(uint32_t i) const { return __ldcg(&trail[i]); }
    __device__ __forceinline__ void tset(uint32_t i, uint32_t lit, uint32_t) { __stcg(&trail[i], lit); }
    // leave the table empty for the next probe: short trails are removed entry by entry (slots found first, then
    // cleared, so that no probe chain is cut while it is still needed), long ones by wiping the table
    template <class G> __device__ __forceinline__ void cleanup(uint32_t tail)
    {
        const uint32_t hs = 1u << log_hs, mask = hs - 1u;
        if (tail * 8u >= hs) { wipe<G>(); return; }
        for (uint32_t i = G::lane(); i < tail; i += G::width) {
            const uint32_t lit = __ldcg(&trail[i]) & TRAIL_LIT;
            uint32_t s = h(lit >> 1);
            while (__ldcg(&hash[s]) != lit) s = (s + 1) & mask;
            __stcg(&trail[i], s);
        }
        G::sync();
        for (uint32_t i = G::lane(); i < tail; i += G::width) __stcg(&hash[__ldcg(&trail[i])], 0u);
        G::sync();
    }
    template <class G> __device__ __forceinline__ void wipe()
    {
        uint4 *p = reinterpret_cast<uint4 *>(hash);
        for (uint32_t i = G::lane(); i < (1u << log_hs) / 4; i += G::width) __stcg(&p[i], make_uint4(0, 0, 0, 0));
        G::sync();
    }
    // call between commits (every table entry is on the trail then): rebuild a larger table once this one is half full
    template <class G> __device__ __forceinline__ void maybe_grow(uint32_t tail)
    {
        if (log_hs >= log_max || 2u * tail <= (1u << log_hs)) return;
        wipe<G>();
        while (log_hs < log_max && 2u * tail > (1u << log_hs)) log_hs = min(log_hs + 2u, log_max);
        for (uint32_t i = G::lane(); i < tail; i += G::width) { uint32_t slot; insert(__ldcg(&trail[i]) & TRAIL_LIT, slot); }   // (a TRAIL_RESCAN duplicate finds its literal present)
        G::sync();
    }
    // after cleanup / wipe: the next probe starts small again
    __device__ __forceinline__ void reset_size() { log_hs = min(GHASH_LOG_MIN, log_max); }
};

// ---------------------------------------------------------------- per-probe state, CTA-wide (shared hash, global trail)
// One CTA owns the probe. Only the randomly accessed part of the state -- the set of true literals -- is in shared
// memory (up to ~200 KB: tens of thousands of literals); the trail, which is appended and consumed sequentially, streams
// through a per-CTA slab in global memory with coalesced accesses. The per-warp global hash of k_deep2 paid a DRAM
// sector for every look-up once the slabs of all resident warps outgrew the L2 (profiles/r02: 323 GB of DRAM reads for
// 1.1 G implied literals on the Tseitin shape); this state pays none. Any table size (multiply-shift range reduction).
struct BhashState {
    uint32_t *hash;   // shared memory [hs], 0 = empty
    uint32_t *trail;  // global [cap]
    uint32_t hs, cap;

    __device__ __forceinline__ uint32_t h(uint32_t var) const { return __umulhi(var * 0x9E3779B1u, hs); }
    __device__ __forceinline__ int lookup(uint32_t lit) const
    {
        const uint32_t var = lit >> 1;
        uint32_t s = h(var);
        for (uint32_t guard = 0;; ++guard) {
            BBCP_CHECK(guard <= hs, "shared hash full on lookup (hs %u cap %u blk %d)\n", hs, cap, blockIdx.x);
#ifdef BBCP_CHECKS
            if (guard >= 64 && bbcp::g_dbg) bbcp::g_dbg[8] = ((unsigned long long)lit << 32) | guard;   // long probe sequences
#endif
            const uint32_t k = *(volatile uint32_t *)&hash[s];
            if (k == 0) return 0;
            if ((k >> 1) == var) return k == lit ? 1 : 2;
            if (++s == hs) s = 0;
        }
    }
    __device__ __forceinline__ int insert(uint32_t lit, uint32_t &slot)
    {
        const uint32_t var = lit >> 1;
        uint32_t s = h(var);
        slot = 0;
        for (uint32_t guard = 0;; ++guard) {
            BBCP_CHECK(guard <= hs, "shared hash full on insert (hs %u cap %u blk %d)\n", hs, cap, blockIdx.x);
            const uint32_t old = atomicCAS(&hash[s], 0u, lit);
            if (old == 0) return 0;
            if ((old >> 1) == var) return old == lit ? 1 : 2;
            if (++s == hs) s = 0;
        }
    }
    __device__ __forceinline__ uint32_t tget(uint32_t i) const { return __ldcg(&trail[i]); }
    __device__ __forceinline__ void tset(uint32_t i, uint32_t lit, uint32_t) { __stcg(&trail[i], lit); }
    template <class G> __device__ __forceinline__ void wipe() {}   // the kernel clears the whole table after every probe
    template <class G> __device__ __forceinline__ void maybe_grow(uint32_t) {}
};

// ---------------------------------------------------------------- per-probe state, last tier (dense, global slab)
struct GmemState {
    uint32_t *bits;   // 2 bits per var: 1 = var true, 2 = var false
    uint32_t *trail;
    uint32_t cap;

    __device__ __forceinline__ int lookup(uint32_t lit) const
    {
        const uint32_t var = lit >> 1;
        const uint32_t o = (__ldcg(&bits[var >> 4]) >> ((var & 15u) * 2u)) & 3u;
        if (o == 0) return 0;
        return (o == 1u) == ((lit & 1u) == 0u) ? 1 : 2;
    }
    __device__ __forceinline__ int insert(uint32_t lit, uint32_t &slot)
    {
        const uint32_t var = lit >> 1, sh = (var & 15u) * 2u, bit = (lit & 1u) ? 2u : 1u;
        const uint32_t o = (atomicOr(&bits[var >> 4], bit << sh) >> sh) & 3u;
        slot = 0;
        if (o == 0) return 0;
        return o == bit ? 1 : 2;
    }
    __device__ __forceinline__ uint32_t tget(uint32_t i) const { return __ldcg(&trail[i]); }
    __device__ __forceinline__ void tset(uint32_t i, uint32_t lit, uint32_t) { __stcg(&trail[i], lit); }
    template <class G> __device__ __forceinline__ void wipe() {}   // cannot overflow: the trail holds every variable
    template <class G> __device__ __forceinline__ void maybe_grow(uint32_t) {}
};

// per-group work counters, kept in registers (uniform over the group) and flushed once
struct WorkCounters {
    unsigned long long props_all = 0, props_ref = 0, slots = 0, fetches = 0, bytes = 0, adopt_lits = 0;
    // the reference's counters: Assign calls, BCP() calls, and m_Implications restricted to the probes that
    // reach a fixpoint (the only part that does not depend on the propagation order)
    unsigned long long assign = 0, assign_ok = 0, props_ref_ok = 0, mark_ref = 0;
    uint32_t n_ok = 0, n_conf = 0, n_skip = 0, adopted = 0, inherited = 0, bcps = 0, hot_rows = 0;
    __device__ __forceinline__ void begin_item() { mark_ref = props_ref; }
    // tail = literals on the trail when the item ended (0 from the warps of a CTA that do not own the tail)
    __device__ __forceinline__ void end_item(bool ok, uint32_t tail, bool propagated)
    {
        assign += tail;
        bcps += propagated && tail;
        if (ok) { assign_ok += tail; props_ref_ok += props_ref - mark_ref; }
    }
    template <class G> __device__ void flush(unsigned long long *c)
    {
        if (G::lane() == 0) {
            if (props_all) atomicAdd(&c[C_PROPS_ALL], props_all);
            if (props_ref) atomicAdd(&c[C_PROPS_REF], props_ref);
            if (slots) atomicAdd(&c[C_SLOTS], slots);
            if (fetches) atomicAdd(&c[C_CLAUSE_FETCH], fetches);
            if (bytes) atomicAdd(&c[C_BYTES], bytes);
            if (n_ok) atomicAdd(&c[C_OK], (unsigned long long)n_ok);
            if (n_conf) atomicAdd(&c[C_CONFLICT], (unsigned long long)n_conf);
            if (n_skip) atomicAdd(&c[C_SKIPPED], (unsigned long long)n_skip);
            if (adopted) atomicAdd(&c[C_ADOPTED], (unsigned long long)adopted);
            if (adopt_lits) atomicAdd(&c[C_ADOPT_LITS], adopt_lits);
            if (inherited) atomicAdd(&c[C_INHERITED], (unsigned long long)inherited);
            if (assign) atomicAdd(&c[C_ASSIGN], assign);
            if (assign_ok) atomicAdd(&c[C_ASSIGN_OK], assign_ok);
            if (props_ref_ok) atomicAdd(&c[C_PROPS_REF_OK], props_ref_ok);
            if (bcps) atomicAdd(&c[C_BCPS], (unsigned long long)bcps);
            if (hot_rows) atomicAdd(&c[C_HOT_ROWS], (unsigned long long)hot_rows);
        }
    }
};

// ---------------------------------------------------------------- where newly implied literals go
// GroupTail: one group owns the probe, the trail tail is a register. BlockTail: the warps of a CTA share one
// probe, the tail and the conflict / overflow latches live in shared memory.
struct GroupTail {
    static constexpr bool HOT_ROWS = true;
    uint32_t tail;
    uint32_t nres = 0;   // TRAIL_RESCAN entries on the trail (not members of the implied set)
    // hot rows (process_batch): ternary / long slots of the rows walked so far, "a TRAIL_SKIP set is on the trail", and
    // "the batch just walked skipped a hot row: the caller owes the second walks"
    unsigned long long lsum = 0;
    bool skipped_set = false, hot_pending = false;
    __device__ __forceinline__ uint32_t get() const { return tail; }
    __device__ __forceinline__ uint32_t members() const { return tail - nres; }
    __device__ __forceinline__ bool aborted() const { return false; }
    template <class G> __device__ __forceinline__ void latch_conflict() {}
    template <class G> __device__ __forceinline__ void drop() { tail = 0; nres = 0; }   // the state was wiped: nothing is on the trail
    // Commit one round of candidate implications (one candidate per lane, 0 = none). Group-converged.
    // ADOPT: the candidates are the literals of a finished closure -- pairwise distinct (no duplicate search) and
    // entered on the trail with `mark` (TRAIL_CLOSED / TRAIL_SKIP). Returns 0 ok, 1 conflict, 2 state overflow.
    template <class G, class S, bool ADOPT = false>
    __device__ __forceinline__ int commit(S &st, uint32_t cand, uint32_t mark = TRAIL_CLOSED)
    {
        const bool has = cand != 0;
        bool leader = has;
        if (!ADOPT) {
            // same literal implied by several clauses in this step: keep the lowest lane (deterministic trail order)
            const unsigned same = G::match_any(has ? cand : (0x80000000u | (uint32_t)G::lane()));
            leader = has && (G::lane() == __ffs(same) - 1);
        }
        uint32_t slot = 0;
        const int r = leader ? st.insert(cand, slot) : 1;
        const bool any_conf = G::any(r == 2);
        const unsigned newm = G::ballot(leader && r == 0);
        const uint32_t nnew = __popc(newm);
        if (tail + nnew > st.cap) return 2;  // caller wipes the whole table: the new entries are not on the trail
        if (leader && r == 0) st.tset(tail + __popc(newm & G::lt()), ADOPT ? (cand | mark) : cand, slot);
        tail += nnew;
        return any_conf ? 1 : 0;
    }
    // second half of an ADOPT commit whose inserts were issued by the caller (several at a time: closure_adopt)
    template <class G, class S>
    __device__ __forceinline__ int append_adopted(S &st, uint32_t cand, int r, uint32_t mark)
    {
        const bool any_conf = G::any(r == 2);
        const unsigned newm = G::ballot(r == 0);
        const uint32_t nnew = __popc(newm);
        if (tail + nnew > st.cap) return 2;
        if (r == 0) st.tset(tail + __popc(newm & G::lt()), cand | mark, 0xffffu);
        tail += nnew;
        return any_conf ? 1 : 0;
    }
    // Second walks (one TRAIL_RESCAN entry each) for the members among the trail entries [0, upto) that will not get a
    // walk of their own any more: the ones before `head` (walked already) and the TRAIL_SKIP ones. Returns 2 when they
    // do not fit.
    template <class G, class S>
    __device__ __forceinline__ int append_rescan(S &st, uint32_t upto, uint32_t head)
    {
        for (uint32_t base = 0; base < upto; base += G::width) {
            const uint32_t i = base + G::lane();
            const uint32_t e = i < upto ? st.tget(i) : TRAIL_RESCAN;
            const uint32_t mode = e & TRAIL_SKIP;
            const bool has = mode != TRAIL_RESCAN && (i < head || mode == TRAIL_SKIP);
            const unsigned m = G::ballot(has);
            const uint32_t n = __popc(m);
            if (tail + n > st.cap) return 2;
            if (has) st.tset(tail + __popc(m & G::lt()), (e & TRAIL_LIT) | TRAIL_RESCAN, 0xffffu);
            tail += n;
            nres += n;
        }
        return 0;
    }
};

struct BlockTail {
    static constexpr bool HOT_ROWS = false;   // (the CTA tiers hold the large probes: their rows are walked as they are)
    uint32_t *ctl;
    unsigned long long lsum = 0;
    bool skipped_set = false, hot_pending = false;  // shared: [0] tail, [1] conflict latch, [2] overflow latch, [8] TRAIL_RESCAN entries on the trail
    __device__ __forceinline__ uint32_t get() const { return *(volatile uint32_t *)&ctl[0]; }
    // Warp-uniform by a vote: the latches are written by OTHER warps at any time, and the lanes of a warp do not
    // necessarily read them in the same cycle; a warp whose lanes disagreed here would split -- some lanes leave
    // process_batch, the rest wait for them at the next shuffle (seen on B200 as a CTA that never finishes)
    __device__ __forceinline__ bool aborted() const { return __any_sync(FULL, (*(volatile uint32_t *)&ctl[1] | *(volatile uint32_t *)&ctl[2]) != 0) != 0; }
    template <class G> __device__ __forceinline__ void latch_conflict() { if (G::lane() == 0) ctl[1] = 1; }
    template <class G> __device__ __forceinline__ void drop() {}   // (the dense third-tier state cannot overflow)
    template <class G, class S, bool ADOPT = false>
    __device__ __forceinline__ int commit(S &st, uint32_t cand, uint32_t mark = TRAIL_CLOSED)
    {
        const bool has = cand != 0;
        bool leader = has;
        if (!ADOPT) {
            const unsigned same = G::match_any(has ? cand : (0x80000000u | (uint32_t)G::lane()));
            leader = has && (G::lane() == __ffs(same) - 1);
        }
        uint32_t slot = 0;
        const int r = leader ? st.insert(cand, slot) : 1;   // duplicates across warps: only one insert returns 0
        const bool any_conf = G::any(r == 2);
        const unsigned newm = G::ballot(leader && r == 0);
        const uint32_t nnew = __popc(newm);
        int ret = 0;
        if (nnew) {
            uint32_t pos = 0;
            if (G::lane() == 0) pos = atomicAdd(&ctl[0], nnew);
            pos = G::shfl(pos, 0);
            if (pos + nnew > st.cap) { if (G::lane() == 0) ctl[2] = 1; ret = 2; }
            else if (leader && r == 0) st.tset(pos + __popc(newm & G::lt()), ADOPT ? (cand | mark) : cand, slot);
        }
        if (any_conf) { if (G::lane() == 0) ctl[1] = 1; if (!ret) ret = 1; }
        return ret;
    }
    template <class G, class S>
    __device__ __forceinline__ int append_adopted(S &st, uint32_t cand, int r, uint32_t mark)
    {
        const bool any_conf = G::any(r == 2);
        const unsigned newm = G::ballot(r == 0);
        const uint32_t nnew = __popc(newm);
        int ret = 0;
        if (nnew) {
            uint32_t pos = 0;
            if (G::lane() == 0) pos = atomicAdd(&ctl[0], nnew);
            pos = G::shfl(pos, 0);
            if (pos + nnew > st.cap) { if (G::lane() == 0) ctl[2] = 1; ret = 2; }
            else if (r == 0) st.tset(pos + __popc(newm & G::lt()), cand | mark, 0xffffu);
        }
        if (any_conf) { if (G::lane() == 0) ctl[1] = 1; if (!ret) ret = 1; }
        return ret;
    }
    // second walks for the trail entries first, first + stride, ... below upto (this warp's share); see GroupTail
    template <class G, class S>
    __device__ __forceinline__ int append_rescan(S &st, uint32_t upto, uint32_t head, uint32_t first = 0, uint32_t stride = 32)
    {
        for (uint32_t base = first; base < upto; base += stride) {
            const uint32_t i = base + G::lane();
            const uint32_t e = i < upto ? st.tget(i) : TRAIL_RESCAN;
            const uint32_t mode = e & TRAIL_SKIP;
            const bool has = mode != TRAIL_RESCAN && (i < head || mode == TRAIL_SKIP);
            const unsigned m = G::ballot(has);
            const uint32_t n = __popc(m);
            if (!n) continue;
            uint32_t pos = 0;
            if (G::lane() == 0) { pos = atomicAdd(&ctl[0], n); atomicAdd(&ctl[8], n); }
            pos = G::shfl(pos, 0);
            if (pos + n > st.cap) { if (G::lane() == 0) ctl[2] = 1; return 2; }
            if (has) st.tset(pos + __popc(m & G::lt()), (e & TRAIL_LIT) | TRAIL_RESCAN, 0xffffu);
        }
        return 0;
    }
};

// Value of a literal for this probe: the probe's own assignment, and -- only while level-0 facts are newer than the rows
// (DbView::l0_dirty) -- the level-0 value. 0 unassigned, 1 true, 2 false.
template <class S>
__device__ __forceinline__ int value_of(const DbView &db, const S &st, uint32_t lit)
{
    int v = st.lookup(lit);
    if (v == 0 && db.l0_dirty) {
        const uint8_t li = __ldg(&db.litinfo[lit]);
        v = (li & LI_TRUE_L0) ? 1 : (li & LI_FALSE_L0) ? 2 : 0;
    }
    return v;
}

template <class S> __device__ __forceinline__ uint2 st_stage_slot(const S &, uint32_t) { return make_uint2(0, 0); }
__device__ __forceinline__ uint2 st_stage_slot(const SmemState &st, uint32_t i) { return st.stage[i]; }
template <class S> struct Staging { static constexpr bool available = false; };
template <> struct Staging<SmemState> { static constexpr bool available = true; };

// Judge a long clause (>= 4 literals) from scratch: TopiBcp.cc:262-367 without the watch bookkeeping.
// Returns the literal to imply (0 = none); sets confl when every literal is false.
template <class S>
__device__ __forceinline__ uint32_t visit_long(const DbView &db, const S &st, uint32_t cref, bool &confl, uint32_t &units16)
{
    BBCP_CHECK(cref >= 1 && cref < db.n_arena_units, "clause ref %u of %u\n", cref, db.n_arena_units);
    const uint4 *cp = db.arena + cref;
    uint4 w = __ldg(cp);
    const uint32_t size = w.x;
    uint32_t nfree = 0, cand = 0, i = 0;
    bool sat = false;
    units16 = 1;
    uint32_t l[4] = {w.y, w.z, w.w, 0};
    int have = 3;
    for (;;) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < have && i < size && !sat) {
                const int v = value_of(db, st, l[k]);
                if (v == 1) sat = true;
                else if (v == 0) { ++nfree; cand = l[k]; }
                ++i;
            }
        }
        if (sat || nfree >= 2 || i >= size) break;
        w = __ldg(++cp);
        ++units16;
        l[0] = w.x; l[1] = w.y; l[2] = w.z; l[3] = w.w;
        have = 4;
    }
    if (sat || nfree >= 2) return 0;
    if (nfree == 0) { confl = true; return 0; }
    return cand;
}

// One group scans the rows of up to W newly true literals (lane i holds literal p for i < B): one header
// gather, then the union of the rows is flattened and walked W eight-byte slots at a time.
// Returns 0 ok, 1 conflict, 2 overflow.
// e = trail entries: literal | mode (bbcp_device.cuh: TRAIL_CLOSED / TRAIL_RESCAN walk the ternary / long part of the row
// only, TRAIL_SKIP walks nothing).
template <class G, class S, class T>
__device__ int process_batch(const DbView &db, S &st, T &tl, uint32_t e, uint32_t B, bool bins_only, WorkCounters &wc)
{
    constexpr int W = G::width;
    const int lane = G::lane();
    const uint32_t p = e & TRAIL_LIT, mode = e & TRAIL_SKIP;
    const bool in = lane < (int)B;
    uint4 hd = make_uint4(0, 0, 0, 0);
    BBCP_CHECK(!in || (p >= 2 && p < 2u * ((uint32_t)db.max_var + 1)), "trail literal %u (B %u lane %d blk %d thr %d)\n", p, B, lane, blockIdx.x, threadIdx.x);
    bool nonempty;
    if (in && mode != TRAIL_SKIP) {
        hd = __ldg(&db.hdr[p ^ 1u]);
        nonempty = (hd.y | hd.z | hd.w) != 0;
    } else nonempty = in && (__ldg(&db.litinfo[p]) & LI_NONEMPTY) != 0;   // (one byte, L2-resident: only the reference's counter wants it)
    // the reference counts a propagation per dequeued literal whose row was ever allocated; a second walk is not a dequeue
    wc.props_ref += __popc(G::ballot(nonempty && mode != TRAIL_RESCAN));
    wc.props_all += __popc(G::ballot(in && mode != TRAIL_RESCAN));
    if (mode != 0) { hd.x += (hd.y + 1u) >> 1; hd.y = 0; }
    if (T::HOT_ROWS && !bins_only) {
        // HOT ROW. The ternary / long part of the row of a literal that occurs in very many clauses (a power-law instance: a
        // million) is walked by every probe that falsifies it, although the probe has assigned a handful of literals.
        // A clause of that row can only be unit if ANOTHER of its literals is false: one assigned later sees this literal
        // when its own row is walked; one assigned earlier is given a second walk (TRAIL_RESCAN) now. So when the row is
        // several times longer than all rows walked so far together, it is not walked (its binary part always is) and the
        // members before this batch are re-walked instead. One such row per batch (the batch-mates are walked in this very
        // call and see the literal), and not after a TRAIL_SKIP adoption (those members are never walked).
        const uint32_t L = (in && mode != TRAIL_SKIP) ? hd.z + hd.w : 0u;
        const bool cand = !tl.skipped_set && (mode == 0 || mode == TRAIL_CLOSED) && L >= 256u && (unsigned long long)L > 4ull * (tl.lsum + 64ull);
        uint32_t key = cand ? ((min(L, 0x03ffffffu) << 5) | (uint32_t)lane) : 0u;
#pragma unroll
        for (int d = W / 2; d; d >>= 1) key = max(key, G::shfl_xor(key, d));
        uint32_t sum = (mode != TRAIL_RESCAN) ? L : 0u;
#pragma unroll
        for (int d = W / 2; d; d >>= 1) sum += G::shfl_xor(sum, d);
        tl.lsum += sum;
        if (key && (int)(key & 31u) == lane) { hd.z = 0; hd.w = 0; }
        if (key) tl.hot_pending = true;
    }
    const uint32_t nb2 = (hd.y + 1u) >> 1;
    const uint32_t cnt = bins_only ? nb2 : nb2 + hd.z + hd.w;
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < W; d <<= 1) {
        const uint32_t t = G::shfl_up(incl, d);
        if (lane >= d) incl += t;
    }
    const uint32_t excl = incl - cnt;
    const uint32_t total = G::shfl(incl, W - 1);
    wc.slots += total;
    wc.bytes += 16ull * __popc(G::ballot(in && mode != TRAIL_SKIP)) + __popc(G::ballot(in && mode == TRAIL_SKIP)) + 8ull * total;   // headers (or one byte of litinfo) + slots

    // one long row and a staging buffer: the slots come through shared memory, STAGE_SLOTS per bulk copy
    bool staged = false;
    uint32_t stage_first = 0;     // slot index (within the row) of stage[0]
    if constexpr (Staging<S>::available) staged = st.stage != nullptr && B == 1 && total >= STAGE_MIN;
    BBCP_WATCH_BEGIN();
    for (uint32_t base = 0; base < total; base += W) {
        BBCP_WATCH(total, base);
        if constexpr (Staging<S>::available) {
            if (staged && base % STAGE_SLOTS == 0) {
                // (the row lives at 8-byte granularity, the bulk copy wants 16: start one slot early if need be)
                const uint32_t row0 = G::shfl(hd.x, 0) + base;
                const uint32_t lead = row0 & 1u;
                const uint32_t n = min(STAGE_SLOTS, total - base) + lead;
                G::sync();     // everybody is done with the previous chunk
                if (lane == 0) bulk_load(st.stage, db.slots + (row0 - lead), ((n + 1u) & ~1u) * 8u, st.stage_bar);
                mbar_wait(st.stage_bar, st.stage_phase);
                st.stage_phase ^= 1u;
                stage_first = base - lead;   // stage[i] = row slot stage_first + i  (stage_first may be "-1" = 0xffffffff: lead slot)
            }
        }
        const uint32_t g = base + lane;
        const bool active = g < total;
        // owner row of slot g: the largest lane i with excl_i <= g
        int lo = 0;
#pragma unroll
        for (int step = W / 2; step; step >>= 1) {
            const int mid = lo + step;
            const uint32_t e = G::shfl(excl, mid & (W - 1));
            if (mid < W && e <= g) lo = mid;
        }
        const uint32_t r_start = G::shfl(hd.x, lo);
        const uint32_t r_nb2 = (G::shfl(hd.y, lo) + 1u) >> 1;
        const uint32_t r_nt = G::shfl(hd.z, lo);
        const uint32_t k = g - G::shfl(excl, lo);
        uint32_t c0 = 0, c1 = 0;
        bool confl = false;
        uint32_t fetched16 = 0;
        if (active) {
            BBCP_CHECK(r_start + k < db.n_slots, "slot %u + %u of %u (lo %d total %u base %u)\n", r_start, k, db.n_slots, lo, total, base);
            uint2 s;
            if (staged) s = st_stage_slot(st, k - stage_first); else s = __ldg(&db.slots[r_start + k]);
            if (k < r_nb2) {
                // binary clauses (TopiBcp.cc:207-246)
                int v = value_of(db, st, s.x);
                if (v == 0) c0 = s.x; else if (v == 2) confl = true;
                if (s.y) {
                    v = value_of(db, st, s.y);
                    if (v == 0) c1 = s.y; else if (v == 2) confl = true;
                }
            } else if (k < r_nb2 + r_nt) {
                // inline ternary clause (falsified literal, s.x, s.y)
                const int va = value_of(db, st, s.x);
                if (va != 1) {
                    const int vb = value_of(db, st, s.y);
                    if (vb != 1) {
                        if (va == 2 && vb == 2) confl = true;
                        else if (va == 2) c0 = s.y;
                        else if (vb == 2) c0 = s.x;
                    }
                }
            } else {
                // long clause: cached literal first (TopiBcp.cc:255-259), then the clause itself
                if (value_of(db, st, s.x) != 1) c0 = visit_long(db, st, s.y, confl, fetched16);
            }
        }
        const unsigned fm = G::ballot(fetched16 != 0);
        if (fm) {
            uint32_t f = fetched16;
#pragma unroll
            for (int d = W / 2; d; d >>= 1) f += G::shfl_xor(f, d);
            wc.fetches += __popc(fm);
            wc.bytes += 16ull * f;
        }
        if (G::any(confl)) { tl.template latch_conflict<G>(); return 1; }
        if (G::any(c0 != 0)) {
            const int r = tl.template commit<G>(st, c0);
            if (r) return r;
        }
        if (G::any(c1 != 0)) {
            const int r = tl.template commit<G>(st, c1);
            if (r) return r;
        }
        G::sync();
        if (tl.aborted()) return 1;  // another warp of the CTA hit a conflict / overflow
        st.template maybe_grow<G>(tl.get());
    }
    return 0;
}

// Closure reuse. The lanes hold newly true literals of this probe (`e`, lane < B). If the batch is an all-literal pass,
// literal q is itself an item of the batch, and when its own probe has already finished its result is adopted:
//   conflict    UP(F and q) derives a conflict, and q is implied here, so this probe conflicts too;
//   overflow    q's set outgrew this tier (it waits in a later tier's queue), and this probe's set contains it: hand
//               this probe on right away instead of propagating up to the capacity first;
//   fixpoint    every literal of q's implied set is implied here as well: the set is read back from the result pool
//               (internal literals, coalesced) and entered 32 literals per step, marked TRAIL_CLOSED -- the set is closed
//               under the binary clauses, so when those literals are dequeued only their ternary / long rows are walked.
// This adds literals of the fixpoint EARLIER than the row scans would find them, nothing else: every trail literal
// still gets its row walked, so the fixpoint, the conflict answer and the sorted implied set are unchanged (unit
// propagation is confluent). What it changes is the shape of the work: the breadth-first frontier of a probe is a few
// literals wide (the kernels were instruction-issue-bound at ~3 literals per 32-lane step), adopted sets fill the lanes.
// `tier`: 1 warp + shared-memory state, 2 / 3 CTA + shared hash (two sizes), 4 dense. Returns 0 / 1 conflict / 2 overflow.
__device__ __forceinline__ uint8_t ld_flag(const uint8_t *p) { return *(const volatile uint8_t *)p; }

// Per lane: the finished result of the lane's literal's own probe, if there is one. Returns 0, 1 (a conflict is
// inherited) or 2 (an overflow is inherited) for the whole group; len > 1 in the lanes that have a set to adopt.
template <class G, class T>
__device__ __forceinline__ int closure_lookup(const BatchView &b, T &tl, uint32_t e, uint32_t B, int tier, WorkCounters &wc, uint32_t &len,
                                              unsigned long long &off)
{
    const int lane = G::lane();
    len = 0;
    off = 0;
    const uint64_t u = (uint64_t)(e & TRAIL_LIT) - 2u - b.first;      // item index of the literal's own probe
    const bool want = lane < (int)B && !(e & TRAIL_SKIP) && (e & TRAIL_LIT) >= 2u + b.first && u < b.n;
    const uint8_t f = want ? ld_flag(&b.flag[u]) : (uint8_t)FLAG_QUEUED;
    if (G::any(f == FLAG_CONFLICT)) { tl.template latch_conflict<G>(); ++wc.inherited; return 1; }
    // tier t hands on what outgrew it with flag FLAG_OVERFLOW - (t - 1); a literal whose probe outgrew THIS tier or a later one
    // takes this probe with it (its set is a subset of this one)
    if (G::any(f >= FLAG_OVERFLOW_LAST && f <= (uint8_t)(FLAG_OVERFLOW + 1 - tier))) { ++wc.inherited; return 2; }
    if (!G::any(f == FLAG_OK) || !(b.opts & OPT_IMPLIED)) return 0;
    __threadfence();   // the set was published before the flag (publish_result)
    if (f == FLAG_OK) {
        len = __ldcg(&b.len[u]);
        if (len & LEN_SELF) len = 0;         // implies only itself
        if (len > 1) off = __ldcg((const unsigned long long *)&b.raw_off[u]); else len = 0;
        BBCP_CHECK(off + len <= b.out_cap, "finished set of item %llu: off %llu len %u cap %llu\n", (unsigned long long)u, off, len, (unsigned long long)b.out_cap);
    }
    return 0;
}

// The group enters literals [first, first + W), [first + stride, ...) ... of a finished set. Returns 0 / 1 / 2.
template <class G, class S, class T>
__device__ __forceinline__ int closure_adopt(const BatchView &b, S &st, T &tl, unsigned long long src, uint32_t n, uint32_t first, uint32_t stride,
                                             uint32_t mark = TRAIL_CLOSED)
{
    const int lane = G::lane();
    BBCP_CHECK(src + n <= b.out_cap, "adopting off %llu len %u of %llu\n", src, n, (unsigned long long)b.out_cap);
    BBCP_WATCH_BEGIN();
    // ADOPT_ILP literals per lane and step: their loads, then their inserts, are in flight together (on the global-memory
    // states an insert is a DRAM atomic; one per step left the adoption bound by that round trip). The first tier's state
    // records a slot index per trail entry at insert time, so it keeps the one-at-a-time form.
    constexpr int ADOPT_ILP = 4;
    uint32_t i = first;
    if constexpr (!Staging<S>::available) {
        for (; (unsigned long long)i + (ADOPT_ILP - 1ull) * stride < n; i += ADOPT_ILP * stride) {
            BBCP_WATCH(n, i);
            uint32_t c[ADOPT_ILP];
            int r[ADOPT_ILP];
#pragma unroll
            for (int k = 0; k < ADOPT_ILP; ++k) c[k] = i + k * stride + lane < n ? (uint32_t)__ldcg(&b.out_lits[src + i + k * stride + lane]) : 0u;
#pragma unroll
            for (int k = 0; k < ADOPT_ILP; ++k) { uint32_t slot; r[k] = c[k] ? st.insert(c[k], slot) : 1; }
            // (a conflict does not cut this short: every literal that was inserted must reach the trail, the clean-up of the
            // state goes by the trail; an overflow wipes the state anyway)
            int ret = 0;
#pragma unroll
            for (int k = 0; k < ADOPT_ILP; ++k) {
                const int rr = tl.template append_adopted<G>(st, c[k], r[k], mark);
                if (rr == 2) return 2;
                if (rr) ret = rr;
            }
            if (ret) return ret;
            if (tl.aborted()) return 1;
            st.template maybe_grow<G>(tl.get());
        }
    }
    for (; i < n; i += stride) {
        BBCP_WATCH(n, i);
        const uint32_t c = i + lane < n ? (uint32_t)__ldcg(&b.out_lits[src + i + lane]) : 0u;
        BBCP_CHECK(c == 0 || (c >= 2 && !(c & TRAIL_SKIP)), "adopted literal %u (off %llu + %u of %u)\n", c, src, i + lane, n);
        const int r = tl.template commit<G, S, true>(st, c, mark);
        if (r) return r;
        if (tl.aborted()) return 1;
        st.template maybe_grow<G>(tl.get());
    }
    return 0;
}

template <class G, class S, class T>
__device__ int expand_closures(const BatchView &b, S &st, T &tl, uint32_t e, uint32_t B, uint32_t head, int tier, WorkCounters &wc)
{
    uint32_t len;
    unsigned long long off;
    const int r0 = closure_lookup<G>(b, tl, e, B, tier, wc, len, off);
    if (r0) return r0;
    unsigned m = G::ballot(len > 1);
    while (m) {
        const int k = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t n = G::shfl(len, k);
        const unsigned long long src = G::shfl(off, k);
        if (n > st.cap) return 2;
        const uint32_t t0 = tl.get();
        // (the first tier's sets are a few hundred literals: not worth the second walks, and its slot index array has no
        // room for them)
        const bool skip = tier >= 2 && n >= 4u * t0 && 2ull * t0 + n <= st.cap;
        int r = closure_adopt<G>(b, st, tl, src, n, 0, G::width, skip ? TRAIL_SKIP : TRAIL_CLOSED);
        if (!r && skip) { r = tl.template append_rescan<G>(st, t0, head); tl.skipped_set = true; }
        if (r) return r;
        ++wc.adopted;
        wc.adopt_lits += n;
        wc.bytes += 4ull * n;
    }
    return 0;
}

// Breadth-first unit propagation of trail[0..tail) to fixpoint by ONE group. Returns 0 fixpoint, 1 conflict, 2 overflow.
template <class G, class S>
__device__ int propagate(const DbView &db, const BatchView &b, S &st, GroupTail &tl, bool prune, int tier, WorkCounters &wc)
{
    const int lane = G::lane();
    uint32_t head = 0;
    while (head < tl.tail) {
        const uint32_t B = min((uint32_t)G::width, tl.tail - head);
        const uint32_t e = lane < (int)B ? st.tget(head + lane) : 0u;
        // while a single literal is assigned only binary clauses can become unit
        const bool first_step = tl.tail == 1;
        if (b.reuse && !first_step) {
            const int r = expand_closures<G>(b, st, tl, e, B, head, tier, wc);
            if (r) return r;
        }
        const int r = process_batch<G>(db, st, tl, e, B, (b.opts & OPT_BINARY_ONLY) || (prune && first_step), wc);
        if (r) return r;
        if (tl.hot_pending) {
            // a hot row was left out: the members walked before this batch get their second walk
            tl.hot_pending = false;
            ++wc.hot_rows;
            if (tl.template append_rescan<G>(st, head, head)) return 2;
        }
        head += B;
    }
    return 0;
}

// A finished probe makes its result visible to the other probes of the launch (expand_closures): the set and its
// position first, the flag last, each behind a fence.
// Deferred form for the first tier, whose probes are a few microseconds long: the two fences of publish_result would be a
// fifth of a probe's life (and a gpu-scope fence also drops the SM's L1 lines, which is where the hot rows live). The warp
// writes set, size and position at once, keeps (item, flag) in a lane and publishes the flags of PUBLISH_BATCH finished probes
// behind ONE fence. A flag that shows a little later only means that a neighbour propagates instead of adopting.
constexpr uint32_t PUBLISH_BATCH = 8;
struct DeferredFlags {
    uint32_t item = 0, n = 0;     // lane k < n holds the k-th finished item
    uint8_t flag = 0;
    template <class G> __device__ __forceinline__ void flush(const BatchView &b)
    {
        if (!n) return;
        if (b.reuse) __threadfence();     // every lane: the sets (and lane 0's sizes / positions) are out
        G::sync();
        if (G::lane() < (int)n) *(volatile uint8_t *)&b.flag[item] = flag;
        n = 0;
    }
    template <class G> __device__ __forceinline__ void add(const BatchView &b, uint64_t it, uint8_t f, uint32_t len, unsigned long long off)
    {
        if (G::lane() == 0) {
            b.len[it] = len; b.raw_off[it] = off;
            if (len) atomicAdd(&b.tile_sum[it / SG_TILE], (unsigned long long)len);
        }
        if (G::lane() == (int)n) { item = (uint32_t)it; flag = f; }
        if (++n == PUBLISH_BATCH) flush<G>(b);
    }
};

template <class G>
__device__ __forceinline__ void publish_result(const BatchView &b, uint64_t item, uint8_t flag, uint32_t len, unsigned long long off)
{
    if (b.reuse) __threadfence();   // every lane: its part of the set is out
    G::sync();
    if (G::lane() == 0) {
        b.len[item] = len; b.raw_off[item] = off;
        if (len) atomicAdd(&b.tile_sum[item / SG_TILE], (unsigned long long)len);
        if (b.reuse) __threadfence();
        *(volatile uint8_t *)&b.flag[item] = flag;
    }
}

// Load a cube into the (empty) state with the semantics of HandleAssumptions (Topi.cc:797-938).
// Returns the flag to report if the cube is decided here (2,3,4 or 1), FLAG_OVERFLOW, or 0xff = propagate.
template <class G, class S, class T>
__device__ int load_cube(const DbView &db, S &st, const int32_t *lits, uint32_t n, T &tl)
{
    const int lane = G::lane();
    bool unmapped = false, false_l0 = false, conflict = false, overflow = false;
    for (uint32_t base = 0; base < n; base += G::width) {
        const uint32_t i = base + lane;
        uint32_t cand = 0;
        if (i < n) {
            const int32_t l = lits[i];
            const uint32_t v = (uint32_t)(l < 0 ? -l : l);
            if (l != 0) {
                const uint32_t il = 2u * v + (l < 0);
                const uint8_t li = v <= (uint32_t)db.max_var ? db.litinfo[il] : 0;
                if (!(li & LI_MAPPED)) unmapped = true;
                else if (li & LI_FALSE_L0) false_l0 = true;
                else if (!(li & LI_TRUE_L0)) cand = il;
            }
        }
        if (!overflow && !conflict) {
            const int r = tl.template commit<G>(st, cand);
            if (r == 1) conflict = true;
            if (r == 2) overflow = true;
            if (!r) st.template maybe_grow<G>(tl.get());
        }
        G::sync();
    }
    const bool decided = G::any(unmapped) || G::any(false_l0);
    if (decided && overflow) {
        // the cube is decided without propagation, but the commit that overflowed left table entries that are not on
        // the trail: the caller's trail-driven clean-up would miss them and they would leak into the next probe
        st.template wipe<G>();
        tl.template drop<G>();
    }
    if (G::any(unmapped)) return FLAG_UNMAPPED;
    if (G::any(false_l0)) return FLAG_FALSE_L0;
    if (overflow) return FLAG_OVERFLOW;
    if (conflict) return FLAG_CONFLICT;
    if (tl.get() == 0) return FLAG_TRIVIAL_L0;
    return 0xff;
}

// load_cube for a depth-1 probe (a single literal, no memory to read)
template <class G, class S, class T>
__device__ int load_probe(const DbView &db, S &st, int32_t l, T &tl)
{
    const uint32_t v = (uint32_t)(l < 0 ? -l : l);
    const uint8_t li = (l != 0 && v <= (uint32_t)db.max_var) ? db.litinfo[2u * v + (l < 0)] : 0;
    if (!(li & LI_MAPPED)) return FLAG_UNMAPPED;
    if (li & LI_FALSE_L0) return FLAG_FALSE_L0;
    if (li & LI_TRUE_L0) return FLAG_TRIVIAL_L0;
    const int r = tl.template commit<G>(st, G::lane() == 0 ? 2u * v + (l < 0) : 0u);
    G::sync();
    return r == 2 ? FLAG_OVERFLOW : 0xff;
}

// group bitonic sort of a[0..n2) ascending, n2 a power of two (shared memory)
template <class G>
__device__ void bitonic_sort(uint32_t *a, uint32_t n2)
{
    const int lane = G::lane();
    for (uint32_t k = 2; k <= n2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = lane; t < n2 / 2; t += G::width) {
                // t-th compare-exchange pair of this stage
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const uint32_t ixj = i | j;
                const bool up = (i & k) == 0;
                const uint32_t x = a[i], y = a[ixj];
                if ((x > y) == up) { a[i] = y; a[ixj] = x; }
            }
            G::sync();
        }
    }
}

// allocate `count` output literals for this group (uniform); returns ~0 if the pool is full
template <class G>
__device__ __forceinline__ unsigned long long out_alloc(const BatchView &b, uint32_t count)
{
    unsigned long long base = 0;
    if (G::lane() == 0) base = atomicAdd(b.out_cursor, (unsigned long long)count);
    base = G::shfl(base, 0);
    return base + count <= b.out_cap ? base : ~0ull;
}

__device__ __forceinline__ void mark_failed(const BatchView &b, uint32_t ilit)
{
    atomicOr(&b.failed[ilit >> 5], 1u << (ilit & 31));
    atomicOr(&b.unit[(ilit ^ 1u) >> 5], 1u << ((ilit ^ 1u) & 31));
}

struct Item {
    const int32_t *lits;
    uint32_t n;
    int32_t probe_lit;  // depth-1 item: its literal (0 otherwise)
};

__device__ __forceinline__ Item get_item(const BatchView &b, uint64_t item)
{
    Item it;
    if (b.cube_off) {
        const uint64_t o = b.cube_off[item];
        it.lits = b.cube_lits + o;
        it.n = (uint32_t)(b.cube_off[item + 1] - o);
        it.probe_lit = it.n == 1 ? it.lits[0] : 0;  // a depth-1 cube is a probe
    } else if (b.cube_lits) {
        it.probe_lit = b.cube_lits[item];   // probe-list mode; 0 is reported as unmapped by the triage
        it.lits = b.cube_lits + item;
        it.n = 1;
    } else {
        const uint64_t idx = b.first + item;
        const int32_t v = (int32_t)(idx / 2 + 1);
        it.probe_lit = (idx & 1) ? -v : v;
        it.lits = nullptr;
        it.n = 1;
    }
    return it;
}

}  // namespace
}  // namespace bbcp
