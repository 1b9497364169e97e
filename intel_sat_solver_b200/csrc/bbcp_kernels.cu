// bbcp_kernels.cu -- batched BCP to fixpoint on sm_100a. Hand-written CUDA; no tensor cores (there is no
// dense contraction on this path); what bounds the kernels (profiles/r02): instruction issue in
// the first tier, random 4-byte accesses to the per-probe state in the later tiers, launch / first-load latency in the
// streaming kernels.
//
// What one probe/cube does here replaces one NewDecLevel/Assign/BCP()/Backtrack(0) round of the reference
// (TopiBcp.cc:103-410, TopiAsg.cc:46-139, TopiBacktrack.cc:18-53):
//   * the database is immutable and shared by all probes; the per-probe state is a sparse assignment
//     (open-addressed hash set of true literals) and a trail -- in shared memory (first tier, k_deep), in a
//     per-warp slab in global memory (second tier, k_deep2) -- or a dense 2-bit-per-variable assignment and a
//     full-length trail in a per-CTA slab (third tier, k_big, any size); the propagation core itself is in
//     bbcp_propagate.cuh;
//   * the trail is also the propagation queue; it is consumed breadth-first, 32 literals per step, one
//     literal per lane for the row-header gather, then the union of the 32 rows is flattened and scanned
//     32 eight-byte slots at a time (coalesced within a row);
//   * rows are OCCURRENCE lists (every clause is listed under each of its literals), so no watch has to
//     move: assignments inside one probe only grow, hence a clause can be judged from scratch whenever one
//     of its literals becomes false, and the visit made for its last-falsified literal sees every earlier
//     assignment. This replaces FindBestWLCand/SwapCurrWatch/WLAddLongWatch/WLRemoveLongWatch
//     (TopiBcp.cc:27-72, TopiWL.cc:115-208), which mutate shared rows;
//   * binary clauses are 4-byte entries as in the reference (Topi.hpp:1952-1962); ternary clauses are held
//     inline in the row ({a,b} per occurrence), long clauses as {blocker, clause} with the reference's
//     cached-literal short-cut (TopiBcp.cc:255-259);
//   * duplicate implications found in the same step are collapsed with __match_any_sync, the insertion into
//     the hash set (atomicCAS) is the linearisation point that detects x / not-x conflicts;
//   * while exactly one literal is assigned, only binary clauses can become unit, so the first step of a
//     depth-1 probe reads the binary part of its row only ("length pruning"); a probe whose literal has no
//     binary clause ends right there. That triage is its own streaming kernel (k_triage: eight probes per
//     thread, one byte of per-literal facts each, a table look-up, flag + set size written, one atomic per
//     warp only where probes are queued); only the probes that can propagate are queued for k_deep;
//   * in an all-literal pass a probe ADOPTS the finished results of the literals it reaches (closure reuse: conflicts
//     and tier overflows are inherited, fixpoints are entered 32 literals per step from the result pool), probes are
//     taken in the order of their height in the binary implication graph, an adopted set larger than the state it is
//     entered into is not walked at all (the smaller side gets second walks instead), and a row several times longer
//     than everything walked so far is left out the same way -- expand_closures / process_batch in bbcp_propagate.cuh,
//     DESIGN.md section 2 for why the fixpoint is unchanged;
//   * k_scan_gather turns the set sizes into CSR offsets and moves every set to its canonical place in one
//     pass (the per-tile sums it needs are produced on the way by k_triage and the propagate kernels), and
//     carries the multi-GPU merge of the bitmaps when peers are attached.
// A conflict ends the probe (NewContradiction's same-level branch, TopiBcp.cc:141-158); only the flag is
// part of the contract for conflicting probes.
#include "bbcp_device.cuh"
#include "bbcp_propagate.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace bbcp {

// diagnostic builds: where a failed check leaves its record (mapped host memory); a no-op in the product build
#ifdef BBCP_CHECKS
__device__ __forceinline__ unsigned long long &bbcp_sweep_words() { return g_phase[7]; }
#else
__device__ unsigned long long g_unused_sweep;
__device__ __forceinline__ unsigned long long &bbcp_sweep_words() { return g_unused_sweep; }
#endif
// diagnostic builds: clock64 ticks per phase of the CTA tiers, summed over CTAs, since the last call
// [0] narrow-frontier mode, [1] closure look-up + adoption (whole CTA), [2] row walks (whole CTA), [3] load, [4] emission,
// [5] clean-up + publication, [6] ticket / idle
void read_phase_ticks(unsigned long long *out)
{
#ifdef BBCP_CHECKS
    cudaMemcpyFromSymbol(out, g_phase, 64);
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_phase, z, 64);
#else
    for (int i = 0; i < 8; ++i) out[i] = 0;
#endif
}

void set_debug_buffer(unsigned long long *mapped)
{
#ifdef BBCP_CHECKS
    cudaMemcpyToSymbol(g_dbg, &mapped, sizeof(mapped));
#else
    (void)mapped;
#endif
}

namespace {

__device__ __forceinline__ uint32_t lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

// Programmatic dependent launch (sm_90+): the three kernels of a batch's first sequence are launched with
// programmatic stream serialization, so the CTAs of the next kernel are scheduled while the previous one drains.
// pdl_wait() = "everything the previous kernel of the stream wrote is visible from here on" (a no-op when the
// kernel was launched the ordinary way); pdl_trigger() = "the next kernel may be scheduled now".
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

// ---------------------------------------------------------------- first tier: shared-memory state
// Result of one probe of the first tier: flag, implied set (sorted in shared memory, written to the pool as
// external literals) or hand-over to the second tier.
template <class G>
__device__ void finish_small(const BatchView &b, SmemState &st, uint64_t item, const Item &it, int res, uint32_t tail, uint32_t nmem,
                             WorkCounters &wc, DeferredFlags &df)
{
    const int lane = G::lane();
    if (res == 2) st.template wipe<G>(); else st.template cleanup<G>(tail);
    uint8_t flag;
    uint32_t len = 0;
    unsigned long long off = 0;
    if (res == 2) {
        flag = FLAG_OVERFLOW;
        if (lane == 0) b.mid_list[atomicAdd(b.mid_count, 1u)] = (uint32_t)item;   // queued before the flag shows (expand_closures)
    } else if (res == 1) {
        flag = FLAG_CONFLICT;
        if (it.probe_lit && lane == 0) mark_failed(b, 2u * (uint32_t)abs(it.probe_lit) + (it.probe_lit < 0));
    } else {
        flag = FLAG_OK;
        if (b.opts & OPT_IMPLIED) {
            uint32_t n2 = 1;
            while (n2 < tail) n2 <<= 1;
            // (second-walk entries are not members: they sort behind the set)
            for (uint32_t i = lane; i < n2; i += G::width) {
                const uint32_t ent = i < tail ? st.trail[i] : TRAIL_RESCAN;
                st.trail[i] = (ent & TRAIL_SKIP) != TRAIL_RESCAN ? (ent & TRAIL_LIT) : 0xffffffffu;
            }
            G::sync();
            if (tail > 1) bitonic_sort<G>(st.trail, n2);
            off = out_alloc<G>(b, nmem);
            if (off == ~0ull) { flag = FLAG_OUT_OVERFLOW; off = 0; }
            else {
                // the pool holds INTERNAL literals (ascending = ascending variable): other probes adopt the set as it is
                // (expand_closures); k_scan_gather turns it into external literals on its way to the canonical place
                len = nmem;
                for (uint32_t i = lane; i < nmem; i += G::width) b.out_lits[off + i] = (int32_t)st.trail[i];
                wc.bytes += 4ull * nmem;
            }
        }
    }
    df.template add<G>(b, item, flag, len, off);
    wc.n_ok += flag == FLAG_OK;
    wc.n_conf += flag == FLAG_CONFLICT;
    if (flag != FLAG_OVERFLOW) wc.end_item(flag == FLAG_OK, nmem, true);   // (a probe handed on is counted by the tier that finishes it)
    G::sync();
}

// Triage: eight consecutive items per thread. Decides every depth-1 item that needs no propagation (unmapped /
// assigned at level 0 / no binary clause on the negation => the implied set is the literal itself, marked
// LEN_SELF so that nothing but the size is written here) and queues the rest for k_deep. The kernel is
// instruction-issue-bound, not bandwidth-bound (6 bytes per item), so the per-item work is a table look-up:
// the five litinfo bits index a 32-entry shared-memory table of (flag, self, counts-as-propagation); in
// universe mode the eight litinfo bytes of a thread are two aligned 32-bit loads + funnel shifts, flags leave
// as one 8-byte store and sizes as two 16-byte stores.
constexpr int TRIAGE_ITEMS = 8;
constexpr int TM_UNIVERSE = 0, TM_LIST = 1, TM_CUBES = 2;
constexpr uint32_t TQ_SELF = 0x100u, TQ_REF = 0x200u, TQ_DEEP = 0x400u, TQ_NONE = 0x8000u;   // NONE: past the end of the batch

template <int MODE>
__global__ void __launch_bounds__(256) k_triage(const DbView db, const BatchView b)
{
    __shared__ unsigned int s_cnt[4];  // ok, skipped, props_ref
    __shared__ uint16_t s_lut[32];
    pdl_trigger();   // k_deep's CTAs may take their places; they wait for this grid to complete before they read anything
    const bool prune = !(b.opts & OPT_NO_LEN_PRUNE) || (b.opts & OPT_BINARY_ONLY);
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x < 32) {
        const uint32_t li = threadIdx.x;
        uint32_t q;
        if (!(li & LI_MAPPED)) q = FLAG_UNMAPPED;
        else if (li & LI_FALSE_L0) q = FLAG_FALSE_L0;
        else if (li & LI_TRUE_L0) q = FLAG_TRIVIAL_L0;
        else if (prune && !(li & LI_HAS_BIN)) q = ((b.opts & OPT_COMPACT_SELF) ? FLAG_SELF : FLAG_OK) | TQ_SELF | ((li & LI_NONEMPTY) ? TQ_REF : 0u);  // implies exactly itself
        else q = 0xffu | TQ_DEEP;
        s_lut[li] = (uint16_t)q;
    }
    // housekeeping: the counter block of the NEXT launch sequence starts from zero (no memset nodes on the stream)
    if (blockIdx.x == 0) for (uint32_t i = threadIdx.x; i < b.zero_words; i += 256) b.zero_next[i] = 0ull;
    __syncthreads();
    // (compact read-back: a self-implying probe is told by a flag bit, its one-literal set is not materialised)
    const uint32_t selflen = (b.opts & OPT_IMPLIED) && !(b.opts & OPT_COMPACT_SELF) ? (LEN_SELF | 1u) : 0u;
    const uint32_t n = (uint32_t)b.n;   // run_batch refuses n >= 2^31
    uint32_t n_ok = 0, n_skip = 0, n_ref = 0;
    const uint32_t stride = gridDim.x * 256u * TRIAGE_ITEMS;
    for (uint32_t wb = blockIdx.x * 256u * TRIAGE_ITEMS; wb < n; wb += stride) {
        // every lane of the warp walks every iteration (the queueing below uses full-warp collectives); lanes past
        // the end of the batch carry TQ_NONE items
        const uint32_t base = wb + threadIdx.x * TRIAGE_ITEMS;
        const bool full = base + TRIAGE_ITEMS <= n;
        uint32_t q[TRIAGE_ITEMS];
        if (MODE == TM_UNIVERSE) {
            // item i is internal literal first + i + 2: eight consecutive litinfo bytes
            const uint64_t il0 = b.first + base + 2;
            const uint32_t *w = reinterpret_cast<const uint32_t *>(db.litinfo) + (il0 >> 2);
            const uint32_t sh = (uint32_t)(il0 & 3u) * 8u;
            const bool in = base < n;   // litinfo is padded by 16 bytes past its last literal, no further
            const uint32_t w0 = in ? __ldg(w) : 0u, w1 = in ? __ldg(w + 1) : 0u, w2 = (in && sh) ? __ldg(w + 2) : 0u;
            const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
#pragma unroll
            for (int k = 0; k < TRIAGE_ITEMS; ++k) {
                const uint32_t li = ((k < 4 ? lo : hi) >> ((k & 3) * 8)) & 31u;
                q[k] = (full || base + k < n) ? s_lut[li] : TQ_NONE;
            }
        } else {
            int32_t pulled[TRIAGE_ITEMS];
            if (MODE == TM_LIST) {
                // the list is in the caller's pinned memory (two 16-byte loads per thread over PCIe, the device copy written
                // for the kernels after this one) or already on the device
                const int32_t *src = b.host_lits ? b.host_lits : b.cube_lits;
                int32_t *copy = const_cast<int32_t *>(b.cube_lits);
                if (full && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(copy)) & 15u) == 0) {
                    const int4 a = __ldcs(reinterpret_cast<const int4 *>(src + base)), c = __ldcs(reinterpret_cast<const int4 *>(src + base) + 1);
                    pulled[0] = a.x; pulled[1] = a.y; pulled[2] = a.z; pulled[3] = a.w;
                    pulled[4] = c.x; pulled[5] = c.y; pulled[6] = c.z; pulled[7] = c.w;
                    if (b.host_lits) { reinterpret_cast<int4 *>(copy + base)[0] = a; reinterpret_cast<int4 *>(copy + base)[1] = c; }
                } else {
#pragma unroll
                    for (int k = 0; k < TRIAGE_ITEMS; ++k) {
                        pulled[k] = (full || base + k < n) ? src[base + k] : 0;
                        if (b.host_lits && (full || base + k < n)) copy[base + k] = pulled[k];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < TRIAGE_ITEMS; ++k) {
                const uint32_t item = base + k;
                q[k] = TQ_NONE;   // past the end
                if (full || item < n) {
                    int32_t lit;
                    bool single = true;
                    if (MODE == TM_LIST) lit = pulled[k];
                    else {
                        const uint64_t o = b.cube_off[item];
                        single = b.cube_off[item + 1] - o == 1;
                        lit = single ? b.cube_lits[o] : 0;
                    }
                    if (lit != 0) {
                        const uint32_t v = (uint32_t)(lit < 0 ? -lit : lit);
                        q[k] = s_lut[v <= (uint32_t)db.max_var ? (db.litinfo[2u * v + (lit < 0)] & 31u) : 0u];
                    } else q[k] = (MODE == TM_LIST) ? FLAG_UNMAPPED : (0xffu | TQ_DEEP);   // a 0 in a probe list names no variable
                }
            }
        }
        uint32_t deepmask = 0, selfs = 0;
#pragma unroll
        for (int k = 0; k < TRIAGE_ITEMS; ++k) {
            const uint32_t x = q[k];
            selfs += (x >> 8) & 1u;
            n_ok += (x >> 8) & 1u;
            n_ref += (x >> 9) & 1u;
            n_skip += (x & 0xffu) >= FLAG_TRIVIAL_L0 && (x & 0xffu) <= FLAG_UNMAPPED;
            deepmask |= ((x >> 10) & 1u) << k;
        }
        if (full) {
            // deep items get their flag / size from the propagate kernels; the placeholder written here is overwritten
            const uint2 fv = make_uint2((q[0] & 0xffu) | ((q[1] & 0xffu) << 8) | ((q[2] & 0xffu) << 16) | (q[3] << 24),
                                        (q[4] & 0xffu) | ((q[5] & 0xffu) << 8) | ((q[6] & 0xffu) << 16) | (q[7] << 24));
            *reinterpret_cast<uint2 *>(b.flag + base) = fv;
            uint4 *lp = reinterpret_cast<uint4 *>(b.len + base);
            lp[0] = make_uint4((q[0] & TQ_SELF) ? selflen : 0u, (q[1] & TQ_SELF) ? selflen : 0u, (q[2] & TQ_SELF) ? selflen : 0u, (q[3] & TQ_SELF) ? selflen : 0u);
            lp[1] = make_uint4((q[4] & TQ_SELF) ? selflen : 0u, (q[5] & TQ_SELF) ? selflen : 0u, (q[6] & TQ_SELF) ? selflen : 0u, (q[7] & TQ_SELF) ? selflen : 0u);
            // the caller's flag buffer gets the same eight bytes (final unless an item was queued: the host checks the queue count)
            if (MODE == TM_LIST && b.host_flag) *reinterpret_cast<uint2 *>(b.host_flag + base) = fv;
        } else {
#pragma unroll
            for (int k = 0; k < TRIAGE_ITEMS; ++k)
                if (base + k < n) {
                    b.flag[base + k] = (uint8_t)q[k]; b.len[base + k] = (q[k] & TQ_SELF) ? selflen : 0u;
                    if (MODE == TM_LIST && b.host_flag) b.host_flag[base + k] = (uint8_t)q[k];
                }
        }
        // the tile's sum of set sizes so far (one literal per self-implying probe): stored, not added, so the array needs
        // no clearing between batches; the propagate kernels add the sizes of the sets they store
        {
            const uint32_t wsum = __reduce_add_sync(FULL, selfs);
            if (lane_id() == 0 && wsum) atomicAdd(&s_cnt[3], wsum);
            __syncthreads();
            if (threadIdx.x == 0) { b.tile_sum[wb / SG_TILE] = selflen ? (unsigned long long)s_cnt[3] : 0ull; s_cnt[3] = 0; }
            __syncthreads();
        }
        // queue the deep items (the warp reserves one run of the list)
        if (__any_sync(FULL, deepmask != 0)) {
            const uint32_t cnt = __popc(deepmask);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, d);
                if (lane_id() >= d) incl += t;
            }
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            uint32_t pos = 0;
            if (lane_id() == 0) pos = atomicAdd(b.deep_count, total);
            pos = __shfl_sync(FULL, pos, 0) + incl - cnt;
            uint32_t m = b.order ? 0u : deepmask;   // with a probe schedule (b.order) the first tier finds the queued items by their flag
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                b.deep_list[pos++] = base + k;
            }
        }
    }
    // one set of counter atomics per block
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        n_ok += __shfl_xor_sync(FULL, n_ok, d);
        n_skip += __shfl_xor_sync(FULL, n_skip, d);
        n_ref += __shfl_xor_sync(FULL, n_ref, d);
    }
    if (lane_id() == 0) {
        if (n_ok) atomicAdd(&s_cnt[0], n_ok);
        if (n_skip) atomicAdd(&s_cnt[1], n_skip);
        if (n_ref) atomicAdd(&s_cnt[2], n_ref);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0]) {
            // a probe decided here assigns its own literal and runs one (empty) propagation
            atomicAdd(&b.counters[C_OK], (unsigned long long)s_cnt[0]); atomicAdd(&b.counters[C_PROPS_ALL], (unsigned long long)s_cnt[0]);
            atomicAdd(&b.counters[C_ASSIGN], (unsigned long long)s_cnt[0]); atomicAdd(&b.counters[C_ASSIGN_OK], (unsigned long long)s_cnt[0]);
            atomicAdd(&b.counters[C_BCPS], (unsigned long long)s_cnt[0]);
        }
        if (s_cnt[1]) atomicAdd(&b.counters[C_SKIPPED], (unsigned long long)s_cnt[1]);
        if (s_cnt[2]) { atomicAdd(&b.counters[C_PROPS_REF], (unsigned long long)s_cnt[2]); atomicAdd(&b.counters[C_PROPS_REF_OK], (unsigned long long)s_cnt[2]); }
    }
}

// Group-per-probe propagation of the queued items, per-probe state in shared memory. W lanes per probe
// (bbcp_propagate.cuh), SMALL_THREADS / W probes in flight per CTA.
template <int W>
__global__ void __launch_bounds__(SMALL_THREADS)
k_deep(const DbView db, const BatchView b, const uint32_t log_hs, const uint32_t cap, const uint32_t stage_rows)
{
    using G = Grp<W>;
    extern __shared__ uint32_t smem[];
    constexpr int GROUPS = SMALL_THREADS / W;
    const int lane = G::lane(), grp = threadIdx.x / W;
    pdl_wait();
    pdl_trigger();
    const uint32_t ndeep = *b.deep_count;
    if ((uint64_t)blockIdx.x * GROUPS >= ndeep) return;
    const uint32_t hs = 1u << log_hs;
    const uint32_t per_grp = hs + cap + cap / 2 + (stage_rows ? 2 * (STAGE_SLOTS + 2) + 4 : 0);  // words: hash, trail, tslot (u16) [, staging buffer, mbarrier]
    SmemState st;
    st.hash = smem + grp * per_grp;
    st.trail = st.hash + hs;
    st.tslot = reinterpret_cast<uint16_t *>(st.trail + cap);
    st.log_hs = log_hs;
    st.cap = cap;
    if (stage_rows) {
        st.stage = reinterpret_cast<uint2 *>(st.trail + cap + cap / 2);      // (hs, cap multiples of 32 words: 16-byte aligned)
        st.stage_bar = reinterpret_cast<unsigned long long *>(st.stage + STAGE_SLOTS + 2);
        if (lane == 0) mbar_init(st.stage_bar);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = lane; i < hs; i += W) st.hash[i] = 0;
    G::sync();

    const bool prune = !(b.opts & OPT_NO_LEN_PRUNE);
    WorkCounters wc;
    // work: the triage queue in queue order, or -- all-literal batches over a database with binary clauses -- the
    // height-sorted schedule b.order, ORDER_CHUNK positions per ticket, of which the ones that name a queued item of
    // this batch are taken (a sharded range sees only its own part of the schedule)
    constexpr uint32_t ORDER_CHUNK = 8;
    unsigned pending = 0;
    uint32_t mine = 0;
    DeferredFlags df;
    for (;;) {
        uint64_t it_idx;
        if (b.order) {
            while (!pending) {
                unsigned t = 0;
                if (lane == 0) t = atomicAdd(b.ticket, ORDER_CHUNK);
                t = G::shfl(t, 0);
                if (t >= b.order_n) break;
                bool ok = false;
                if (lane < (int)ORDER_CHUNK && t + lane < b.order_n) {
                    const uint64_t u = (uint64_t)__ldg(&b.order[t + lane]) - 2u;
                    if (u >= b.first && u - b.first < b.n) { mine = (uint32_t)(u - b.first); ok = b.flag[mine] == FLAG_QUEUED; }
                }
                pending = G::ballot(ok);
            }
            if (!pending) break;
            const int k = __ffs(pending) - 1;
            pending &= pending - 1;
            it_idx = G::shfl(mine, k);
        } else {
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(b.ticket, 1u);
            t = G::shfl(t, 0);
            if (t >= ndeep) break;
            it_idx = b.deep_list[t];
        }
        BBCP_CHECK(it_idx < b.n, "queued item %llu of %llu (ndeep %u)\n", (unsigned long long)it_idx, (unsigned long long)b.n, ndeep);
        const Item it = get_item(b, it_idx);
        GroupTail tl{0};
        wc.begin_item();
        int f = it.probe_lit ? load_probe<G>(db, st, it.probe_lit, tl) : load_cube<G>(db, st, it.lits, it.n, tl);
        int res;
        if (f == 0xff) res = propagate<G>(db, b, st, tl, prune, 1, wc);
        else if (f == FLAG_OVERFLOW) res = 2;
        else if (f == FLAG_CONFLICT) res = 1;
        else {
            st.template cleanup<G>(tl.tail);
            df.template add<G>(b, it_idx, (uint8_t)f, 0, 0);
            ++wc.n_skip;
            G::sync();
            continue;
        }
        finish_small<G>(b, st, it_idx, it, res, tl.tail, tl.members(), wc, df);
    }
    df.template flush<G>(b);
    wc.template flush<G>(b.counters);
}

// ---------------------------------------------------------------- second tier: group per probe, global hash state
// Probes that outgrew the shared-memory state. Same propagation code, state in a per-group slab (GhashState).
// The implied set is written to the pool UNSORTED as internal literals and the item is queued for k_sort_sets.
template <int W>
__global__ void __launch_bounds__(MID2_THREADS, 4)
k_deep2(const DbView db, const BatchView b, uint32_t *slab, const uint32_t log_hs, const uint32_t cap)
{
    using G = Grp<W>;
    constexpr int GROUPS = MID2_THREADS / W;
    const int lane = G::lane();
    const uint32_t count = *b.mid_count;
    const uint32_t ggrp = blockIdx.x * GROUPS + threadIdx.x / W;
    if (ggrp >= count) return;
    GhashState st;
    st.hash = slab + (size_t)ggrp * ((1u << log_hs) + cap);
    st.trail = st.hash + (1u << log_hs);
    st.log_max = log_hs;
    st.cap = cap;
    st.reset_size();
    const bool prune = !(b.opts & OPT_NO_LEN_PRUNE);
    WorkCounters wc;
    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(b.ticket + 1, 1u);
        t = G::shfl(t, 0);
        if (t >= count) break;
        const uint64_t item = b.mid_list[t];
        const Item it = get_item(b, item);
        GroupTail tl{0};
        wc.begin_item();
        const int f = it.probe_lit ? load_probe<G>(db, st, it.probe_lit, tl) : load_cube<G>(db, st, it.lits, it.n, tl);
        int res;
        if (f == 0xff) res = propagate<G>(db, b, st, tl, prune, 2, wc);
        else if (f == FLAG_OVERFLOW) res = 2;
        else if (f == FLAG_CONFLICT) res = 1;
        else {
            st.template cleanup<G>(tl.tail);
            st.reset_size();
            publish_result<G>(b, item, (uint8_t)f, 0, 0);
            if (lane == 0) atomicAdd(&b.counters[C_MID], 1ull);
            ++wc.n_skip;
            G::sync();
            continue;
        }
        uint8_t flag;
        uint32_t len = 0;
        unsigned long long off = 0;
        const uint32_t tail = tl.tail;
        if (res == 2) {
            flag = FLAG_OVERFLOW2;
            if (lane == 0) b.big_list[atomicAdd(b.big_count, 1u)] = (uint32_t)item;
            st.template wipe<G>();   // some table entries have no trail entry
            st.reset_size();
        } else {
            if (res == 1) {
                flag = FLAG_CONFLICT;
                if (it.probe_lit && lane == 0) mark_failed(b, 2u * (uint32_t)abs(it.probe_lit) + (it.probe_lit < 0));
            } else {
                flag = FLAG_OK;
                if (b.opts & OPT_IMPLIED) {
                    const uint32_t nmem = tl.members();
                    off = out_alloc<G>(b, nmem);
                    if (off == ~0ull) { flag = FLAG_OUT_OVERFLOW; off = 0; }
                    else {
                        // internal literals, trail order, without the second-walk entries
                        len = nmem;
                        uint32_t w = 0;
                        for (uint32_t base = 0; base < tail; base += W) {
                            const uint32_t ent = base + lane < tail ? st.tget(base + lane) : TRAIL_RESCAN;
                            const bool mem = (ent & TRAIL_SKIP) != TRAIL_RESCAN;
                            const unsigned mm = G::ballot(mem);
                            if (mem) b.out_lits[off + w + __popc(mm & G::lt())] = (int32_t)(ent & TRAIL_LIT);
                            w += __popc(mm);
                        }
                        if (lane == 0) b.sort_list[atomicAdd(b.sort_count, 1u)] = (uint32_t)item;
                        wc.bytes += 4ull * tail;
                    }
                }
            }
            G::sync();
            st.template cleanup<G>(tail);
            st.reset_size();
        }
        publish_result<G>(b, item, flag, len, off);
        if (lane == 0 && flag != FLAG_OVERFLOW2) atomicAdd(&b.counters[C_MID], 1ull);
        wc.n_ok += flag == FLAG_OK;
        wc.n_conf += flag == FLAG_CONFLICT;
        if (flag != FLAG_OVERFLOW2) wc.end_item(flag == FLAG_OK, tl.members(), true);
        G::sync();
    }
    wc.template flush<G>(b.counters);
}

// Sorts the implied sets the second tier left unsorted in the pool (ascending internal literal = ascending
// variable). One CTA per set (dynamic tickets: the sets differ 16-fold in size), bitonic network in shared memory; the
// compare-exchange steps at distances below 32 run on warp shuffles (a warp owns 32 consecutive keys), so a merge of
// 2^m keys costs m - 5 CTA barriers instead of m.
constexpr int SORT_THREADS = 256;
__device__ __forceinline__ uint32_t sort_cex(uint32_t x, uint32_t j, bool up, int lane)
{
    const uint32_t y = __shfl_xor_sync(FULL, x, j);
    const bool lower = (lane & j) == 0;
    return (lower == up) ? min(x, y) : max(x, y);
}
__global__ void __launch_bounds__(SORT_THREADS) k_sort_sets(const BatchView b, const uint32_t cap)
{
    extern __shared__ uint32_t s_keys[];
    __shared__ uint32_t s_e;
    const uint32_t count = *b.sort_count;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    constexpr uint32_t NW = SORT_THREADS / 32;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_e = atomicAdd(b.ticket + 2, 1u);   // (the CTA-per-probe tiers that own this ticket do not run beside the warp-per-probe second tier)
        __syncthreads();
        const uint32_t e = s_e;
        if (e >= count) break;
        const uint32_t item = b.sort_list[e];
        const uint32_t len = b.len[item] & 0x7fffffffu;
        int32_t *set = b.out_lits + b.raw_off[item];
        uint32_t n2 = 32;
        while (n2 < len) n2 <<= 1;
        // load + the merges up to 32 keys, all inside a warp's block of 32 keys
        for (uint32_t i = threadIdx.x; i < n2; i += SORT_THREADS) {
            uint32_t x = i < len ? (uint32_t)set[i] : 0xffffffffu;
#pragma unroll
            for (uint32_t k = 2; k <= 32; k <<= 1) {
                const bool up = (i & k) == 0;
#pragma unroll
                for (uint32_t j = k >> 1; j > 0; j >>= 1) x = sort_cex(x, j, up, lane);
            }
            s_keys[i] = x;
        }
        __syncthreads();
        for (uint32_t k = 64; k <= n2; k <<= 1) {
            for (uint32_t j = k >> 1; j >= 32; j >>= 1) {
                for (uint32_t t = threadIdx.x; t < n2 / 2; t += SORT_THREADS) {
                    const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const uint32_t ixj = i | j;
                    const bool up = (i & k) == 0;
                    const uint32_t x = s_keys[i], y = s_keys[ixj];
                    if ((x > y) == up) { s_keys[i] = y; s_keys[ixj] = x; }
                }
                __syncthreads();
            }
            for (uint32_t blk = warp; blk < n2 / 32; blk += NW) {
                const uint32_t i = blk * 32 + lane;
                const bool up = (i & k) == 0;
                uint32_t x = s_keys[i];
#pragma unroll
                for (uint32_t j = 16; j > 0; j >>= 1) x = sort_cex(x, j, up, lane);
                s_keys[i] = x;
            }
            __syncthreads();
        }
        for (uint32_t i = threadIdx.x; i < len; i += SORT_THREADS) set[i] = (int32_t)s_keys[i];
    }
}

// ---------------------------------------------------------------- third tier: one CTA per probe, dense state
// Probes of any size: the warps of a CTA share the per-probe state (dense 2-bit assignment + full-length trail in
// a per-CTA slab in global memory) and walk each breadth-first level of the trail together (level-synchronous,
// two barriers per level).
constexpr uint32_t ADOPT_MAX = 64;   // finished sets adopted CTA-wide per level; further ones are entered by the warp that found them
// narrow frontiers: while a level holds at most NARROW_IN literals, warp 0 walks the successive levels alone (no CTA
// barrier per level -- a deep implication chain costs a CTA five barriers per literal otherwise); it hands back to the
// whole CTA when the frontier widens beyond NARROW_OUT or a finished set of more than NARROW_ADOPT literals shows up
constexpr uint32_t NARROW_IN = 32, NARROW_OUT = 128, NARROW_ADOPT = 512;
constexpr uint32_t BULK_MEMBERS = 64;   // bulk SKIP adoption: at most this many trail entries before the set (one thread each looks its literal up)

// ctl: [0] tail, [1] conflict latch, [2] overflow latch, [6] "the next level is for the whole CTA", [7] head handed back by warp 0
template <class S>
__device__ void block_propagate(const DbView &db, const BatchView &b, S &st, uint32_t *ctl, bool prune, int tier, WorkCounters &wc)
{
    using G = Grp<32>;
    __shared__ unsigned long long s_src[ADOPT_MAX];
    __shared__ uint32_t s_len[ADOPT_MAX];
    __shared__ uint32_t s_na, s_ndup, s_base;
    __shared__ uint32_t s_dup[BULK_MEMBERS];
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    BlockTail tl{ctl};
    const volatile uint32_t *vctl = ctl;
    uint32_t head = 0;
    BBCP_WATCH_BEGIN();
    BBCP_PHASE_DECL();
    for (;;) {
        BBCP_WATCH(head, ctl[0]);
        __syncthreads();
        const uint32_t level_end = vctl[0];
        const bool stop = (vctl[1] | vctl[2]) != 0 || head >= level_end;
        const bool narrow = !stop && !vctl[6] && level_end - head <= NARROW_IN;
        if (threadIdx.x == 0) s_na = 0;
        __syncthreads();   // everybody has its snapshot before anyone appends
        if (stop) break;
        if (narrow) {
            if (warp == 0) {
                uint32_t hcur = head;
                for (;;) {
                    const uint32_t end = vctl[0];    // only this warp appends now
                    if (hcur >= end) break;
                    if (end - hcur > NARROW_OUT) { if (lane == 0) ctl[6] = 1; break; }
                    const uint32_t B = min(32u, end - hcur);
                    const uint32_t e = lane < (int)B ? st.tget(hcur + lane) : 0u;
                    if (b.reuse && end > 1) {
                        uint32_t len;
                        unsigned long long off;
                        const int r = closure_lookup<G>(b, tl, e, B, tier, wc, len, off);
                        if (r == 2 && lane == 0) ctl[2] = 1;          // an inherited overflow: hand the probe on
                        if (r) break;
                        if (G::any(len > NARROW_ADOPT)) { if (lane == 0) ctl[6] = 1; break; }   // a large set: all warps enter it
                        unsigned m = G::ballot(len > 1);
                        int ra = 0;
                        while (m && !ra) {
                            const int k = __ffs(m) - 1;
                            m &= m - 1;
                            const uint32_t n = G::shfl(len, k), t0 = vctl[0];
                            const bool skip = n >= 4u * t0 && 2ull * t0 + n <= st.cap;     // SKIP mode, see expand_closures
                            ra = closure_adopt<G>(b, st, tl, G::shfl(off, k), n, 0, 32, skip ? TRAIL_SKIP : TRAIL_CLOSED);
                            if (!ra && skip) ra = tl.template append_rescan<G>(st, t0, hcur);
                            if (!ra) { ++wc.adopted; wc.adopt_lits += n; wc.bytes += 4ull * n; }
                        }
                        if (ra) break;
                    }
                    if (process_batch<G>(db, st, tl, e, B, (b.opts & OPT_BINARY_ONLY) || (prune && end == 1), wc)) break;
                    hcur += B;
                }
                if (lane == 0) ctl[7] = hcur;
            }
            __syncthreads();
            head = vctl[7];
            BBCP_PHASE(0);
            continue;
        }
        if (b.reuse && level_end > 1) {
            // closure reuse, CTA-wide: the warps look up the literals of the level, the finished sets found are listed in
            // shared memory, then ALL warps enter each set together (a set of this tier has tens of thousands of literals;
            // entered by the one warp that found it, the other seven would wait at the barrier)
            for (uint32_t bse = head + warp * 32; bse < level_end; bse += nwarps * 32) {
                const uint32_t B = min(32u, level_end - bse);
                const uint32_t e = lane < (int)B ? st.tget(bse + lane) : 0u;
                uint32_t len;
                unsigned long long off;
                const int r = closure_lookup<G>(b, tl, e, B, tier, wc, len, off);
                if (r == 2 && lane == 0) ctl[2] = 1;                  // an inherited overflow: hand the probe on
                if (r) break;
                if (G::any(len > st.cap)) { if (lane == 0) ctl[2] = 1; break; }   // a set that cannot fit: hand the probe on right away
                uint32_t slot = ADOPT_MAX;
                if (len > 1) slot = atomicAdd(&s_na, 1u);
                if (slot < ADOPT_MAX) { s_src[slot] = off; s_len[slot] = len; }
                unsigned m = G::ballot(len > 1 && slot >= ADOPT_MAX);
                while (m) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    if (closure_adopt<G>(b, st, tl, G::shfl(off, k), G::shfl(len, k), 0, 32)) break;
                }
            }
            __syncthreads();
            const uint32_t na = min(s_na, ADOPT_MAX);
            // the largest set of the level may go in SKIP mode (expand_closures): its literals get no walk, the members
            // that were there before and will not be walked any more get a second one instead
            uint32_t best = 0;
            for (uint32_t a = 1; a < na; ++a) if (s_len[a] > s_len[best]) best = a;
            const uint32_t t0 = vctl[0];
            const bool skip = na && !((vctl[1] | vctl[2]) != 0) && s_len[best] >= 4u * t0 && 2ull * t0 + s_len[best] <= st.cap;
            __syncthreads();   // everybody has t0 before anyone appends
            if (skip && tl.template append_rescan<G>(st, t0, head, warp * 32, nwarps * 32)) { /* overflow latched */ }
            // BULK form of the SKIP adoption (dense state, few members so far): the set is sorted, so whether a member of the
            // state is in it -- or its negation: a conflict -- is a binary search per member; after that no insert needs its
            // answer: the set is copied to the trail as it is (a literal that was a member already is entered as a second
            // walk, which does not count as a member) and the bits are set with fire-and-forget reductions. A streaming copy
            // instead of a DRAM atomic round trip per 32 literals.
            bool bulk_done = false;
            if constexpr (std::is_same<S, GmemState>::value) {
                if (skip && t0 <= BULK_MEMBERS) {
                    const uint32_t n = s_len[best];
                    const unsigned long long src = s_src[best];
                    const uint32_t *A = reinterpret_cast<const uint32_t *>(b.out_lits) + src;
                    if (threadIdx.x == 0) s_ndup = 0;
                    __syncthreads();
                    if (threadIdx.x < t0) {
                        const uint32_t e0 = st.tget(threadIdx.x);
                        if ((e0 & TRAIL_SKIP) != TRAIL_RESCAN) {
                            const uint32_t lit = e0 & TRAIL_LIT, v2 = lit & ~1u;     // the set holds at most one literal of a variable
                            uint32_t lo = 0, hi = n;
                            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((__ldcg(&A[mid]) & ~1u) < v2) lo = mid + 1; else hi = mid; }
                            if (lo < n) {
                                const uint32_t x = __ldcg(&A[lo]);
                                if (x == lit) s_dup[atomicAdd(&s_ndup, 1u)] = lo;
                                else if (x == (lit ^ 1u)) ctl[1] = 1;               // the set holds the negation of a member
                            }
                        }
                    }
                    __syncthreads();
                    const uint32_t ndup = s_ndup;
                    if (!(vctl[1] | vctl[2])) {
                        if (threadIdx.x == 0) { s_base = atomicAdd(&ctl[0], n); atomicAdd(&ctl[8], ndup); }
                        __syncthreads();
                        const uint32_t base0 = s_base;
                        if ((unsigned long long)base0 + n > st.cap) { if (threadIdx.x == 0) ctl[2] = 1; }
                        else {
                            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                                const uint32_t c = __ldcg(&A[i]);
                                bool dup = false;
                                for (uint32_t q = 0; q < ndup; ++q) dup |= s_dup[q] == i;
                                st.tset(base0 + i, c | (dup ? TRAIL_RESCAN : TRAIL_SKIP), 0xffffu);
                                if (!dup) {
                                    const uint32_t var = c >> 1;
                                    atomicOr(&st.bits[var >> 4], ((c & 1u) ? 2u : 1u) << ((var & 15u) * 2u));   // result unused: a reduction
                                }
                            }
                            if (threadIdx.x == 0) { ++wc.adopted; wc.adopt_lits += n; wc.bytes += 8ull * n; }
                        }
                    }
                    bulk_done = true;
                    __syncthreads();
                }
            }
            for (uint32_t k = bulk_done ? 1u : 0u; k < na && !tl.aborted(); ++k) {
                const uint32_t a = k == 0 ? best : (k == best ? 0 : k);    // the largest first
                const uint32_t n = s_len[a];
                // successive sets start at successive warps, so that many short sets spread over the CTA as well
                if (closure_adopt<G>(b, st, tl, s_src[a], n, ((warp + nwarps - k % nwarps) % nwarps) * 32, nwarps * 32, skip && k == 0 ? TRAIL_SKIP : TRAIL_CLOSED)) break;
                if (warp == (int)(k % nwarps) && lane == 0) { ++wc.adopted; wc.adopt_lits += n; wc.bytes += 4ull * n; }
            }
            // (CTA-uniform: a warp that got past this point may latch a conflict before a slower thread has looked)
            const bool ab = __syncthreads_or((vctl[1] | vctl[2]) != 0);
            BBCP_PHASE(1);
            if (ab) continue;   // the loop head sees the latch and stops
        }
        const bool bins_only = (b.opts & OPT_BINARY_ONLY) || (prune && level_end == 1);
        for (uint32_t bse = head + warp * 32; bse < level_end; bse += nwarps * 32) {
            const uint32_t B = min(32u, level_end - bse);
            const uint32_t e = lane < (int)B ? st.tget(bse + lane) : 0u;
            if (process_batch<G>(db, st, tl, e, B, bins_only, wc)) break;
        }
        if (threadIdx.x == 0) ctl[6] = 0;   // (read again only behind the barrier at the loop head)
        head = level_end;
        BBCP_PHASE(2);
    }
}

constexpr int SWEEP_WORDS = 4;   // words of the dense assignment one thread sweeps per step: one 16-byte load
__global__ void __launch_bounds__(BIG_THREADS, 5) k_big(const DbView db, const BatchView b, const LargeState ls, const int from_huge)
{
    using G = Grp<32>;
    __shared__ uint32_t ctl[12];     // 0 tail, 1 conflict, 2 overflow, 3 load result, 4 ticket, 5 vmin, 6 vmax / whole-CTA level, 7 head, 8 second-walk entries
    __shared__ unsigned long long s_off;
    __shared__ uint32_t s_sweep[2][BIG_THREADS / 32];
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // the last tier takes what outgrew the tier before it: the third tier's hand-over list, or the second tier's when the
    // warp-per-probe second tier (k_deep2, BBCP_TIER2=warp) ran instead of the two CTA-per-probe tiers
    const uint32_t *list = from_huge ? b.huge_list : b.big_list;
    const uint32_t count = from_huge ? *b.huge_count : *b.big_count;
    if (blockIdx.x >= count) return;
    unsigned int *ticket = b.ticket + 3;

    GmemState gs;
    gs.bits = ls.bits + (size_t)blockIdx.x * db.nvars_words16;
    gs.cap = big_trail_words((uint32_t)db.max_var);
    gs.trail = ls.trail + (size_t)blockIdx.x * gs.cap;
    const bool prune = !(b.opts & OPT_NO_LEN_PRUNE);
    WorkCounters wc;
    BBCP_PHASE_DECL();
    for (;;) {
        __syncthreads();
        BBCP_PHASE(5);
        if (threadIdx.x == 0) {
            ctl[4] = atomicAdd(ticket, 1u);
            ctl[0] = ctl[1] = ctl[2] = ctl[8] = 0;
            ctl[3] = 0xff;
            ctl[5] = 0xffffffffu;
            ctl[6] = 0;
        }
        __syncthreads();
        const uint32_t t = ctl[4];
        if (t >= count) break;
        const uint64_t item = list[t];
        const Item it = get_item(b, item);
        wc.begin_item();
        if (warp == 0) {
            BlockTail tl{ctl};
            const int f = it.probe_lit ? load_probe<G>(db, gs, it.probe_lit, tl) : load_cube<G>(db, gs, it.lits, it.n, tl);
            if (lane == 0) ctl[3] = (uint32_t)f;
        }
        __syncthreads();
        const uint32_t f = ctl[3];
        BBCP_PHASE(3);
        if (f == 0xff) block_propagate(db, b, gs, ctl, prune, 4, wc);
        __syncthreads();
        BBCP_PHASE_DECL2();
        const uint32_t tail = min(ctl[0], gs.cap);
        const uint32_t nmem = tail - min(ctl[8], tail);     // members of the implied set (the trail also holds second-walk entries)
        const bool conflict = ctl[1] != 0 || f == FLAG_CONFLICT;
        uint8_t flag = (uint8_t)f;
        uint32_t len = 0;
        unsigned long long off = 0;
        bool swept = false;   // the emission sweep cleared the dense assignment
        if (f == 0xff || f == FLAG_CONFLICT) {
            if (conflict) {
                flag = FLAG_CONFLICT;
                if (it.probe_lit && threadIdx.x == 0) mark_failed(b, 2u * (uint32_t)abs(it.probe_lit) + (it.probe_lit < 0));
            } else {
                flag = FLAG_OK;
                if (b.opts & OPT_IMPLIED) {
                    if (threadIdx.x == 0) {
                        const unsigned long long o = atomicAdd(b.out_cursor, (unsigned long long)nmem);
                        s_off = o + nmem <= b.out_cap ? o : ~0ull;
                    }
                    __syncthreads();
                    off = s_off;
                    if (off == ~0ull) { flag = FLAG_OUT_OVERFLOW; off = 0; }
                    else {
                        // ordered emission: sweep the dense assignment between the smallest and largest assigned var
                        len = nmem;
                        swept = true;
                        uint32_t vmin = 0xffffffffu, vmax = 0;
                        for (uint32_t i = threadIdx.x; i < tail; i += blockDim.x) {
                            const uint32_t var = (gs.tget(i) & TRAIL_LIT) >> 1;
                            vmin = min(vmin, var);
                            vmax = max(vmax, var);
                        }
#pragma unroll
                        for (int d = 16; d; d >>= 1) {
                            vmin = min(vmin, __shfl_xor_sync(FULL, vmin, d));
                            vmax = max(vmax, __shfl_xor_sync(FULL, vmax, d));
                        }
                        if (lane == 0) { atomicMin(&ctl[5], vmin); atomicMax(&ctl[6], vmax); }
                        __syncthreads();
                        BBCP_PHASE(6);
                        const uint32_t w0 = (ctl[5] >> 4) & ~(uint32_t)(SWEEP_WORDS - 1), w1 = ctl[6] >> 4;
                        if (threadIdx.x == 0) atomicAdd(&bbcp_sweep_words(), (unsigned long long)(w1 - w0 + 1));
                        // The range is wide and sparse (on the Tseitin shape ~100 K words for ~18 K literals: a cone spans many
                        // frames). A thread takes SWEEP_WORDS consecutive words with one 16-byte load, so that one warp scan and
                        // one barrier order a whole step (the partial sums alternate between two buffers); the next step's load
                        // is in flight while this one is scanned and written out.
                        unsigned long long w = off;
                        const uint4 *bits4 = reinterpret_cast<const uint4 *>(gs.bits);     // slabs are whole 16-byte vectors (nvars_words16)
                        const uint32_t stride = SWEEP_WORDS * blockDim.x;
                        uint32_t wi = w0 + SWEEP_WORDS * threadIdx.x;
                        uint4 cur = wi <= w1 ? __ldcg(&bits4[wi >> 2]) : make_uint4(0, 0, 0, 0);
                        for (uint32_t wbase = w0, par = 0; wbase <= w1; wbase += stride, wi += stride, par ^= 1u) {
                            const uint4 nxt = wi + stride <= w1 ? __ldcg(&bits4[(wi + stride) >> 2]) : make_uint4(0, 0, 0, 0);
                            const uint32_t word[SWEEP_WORDS] = {cur.x, cur.y, cur.z, cur.w};
                            // the sweep also leaves the state empty for the next probe (every assigned variable lies in its range)
                            if (cur.x | cur.y | cur.z | cur.w) __stcg(reinterpret_cast<uint4 *>(gs.bits) + (wi >> 2), make_uint4(0, 0, 0, 0));
                            uint32_t c = 0;
#pragma unroll
                            for (int k = 0; k < SWEEP_WORDS; ++k) c += __popc((word[k] | (word[k] >> 1)) & 0x55555555u);  // one bit per assigned var
                            uint32_t incl = c;
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                const uint32_t x = __shfl_up_sync(FULL, incl, d);
                                if (lane >= d) incl += x;
                            }
                            if (lane == 31) s_sweep[par][warp] = incl;
                            __syncthreads();
                            uint32_t before = 0, total = 0;
                            for (int q = 0; q < nwarps; ++q) { const uint32_t x = s_sweep[par][q]; total += x; if (q < warp) before += x; }
                            unsigned long long o = w + before + incl - c;
#pragma unroll
                            for (int k = 0; k < SWEEP_WORDS; ++k) {
                                uint32_t m = (word[k] | (word[k] >> 1)) & 0x55555555u;
                                while (m) {
                                    const int bpos = __ffs(m) - 1;
                                    m &= m - 1;
                                    const uint32_t var = (wi + k) * 16u + (uint32_t)(bpos >> 1);
                                    const bool isneg = ((word[k] >> bpos) & 3u) == 2u;
                                    b.out_lits[o++] = (int32_t)(2u * var + (isneg ? 1u : 0u));   // internal literals, as every set in the pool
                                }
                            }
                            w += total;
                            cur = nxt;
                        }
                        wc.bytes += warp == 0 ? 4ull * tail + 4ull * (w1 - w0 + 1) : 0ull;
                    }
                }
            }
        }
        __syncthreads();
        BBCP_PHASE(4);
        // leave the state empty for the next probe
        for (uint32_t i = threadIdx.x; i < tail && !swept; i += blockDim.x) {
            const uint32_t var = (gs.tget(i) & TRAIL_LIT) >> 1;
            atomicAnd(&gs.bits[var >> 4], ~(3u << ((var & 15u) * 2u)));
        }
        if (b.reuse) __threadfence();   // every thread: its part of the set is out (publish_result, CTA-wide)
        __syncthreads();
        if (threadIdx.x == 0) {
            b.len[item] = len; b.raw_off[item] = off;
            if (len) atomicAdd(&b.tile_sum[item / SG_TILE], (unsigned long long)len);
            if (b.reuse) __threadfence();
            *(volatile uint8_t *)&b.flag[item] = flag;
            atomicAdd(&b.counters[C_LARGE], 1ull);
            wc.n_ok += flag == FLAG_OK;
            wc.n_conf += flag == FLAG_CONFLICT;
            wc.n_skip += flag >= FLAG_TRIVIAL_L0 && flag <= FLAG_UNMAPPED;
        }
        // every warp settles its own share of the row counters; the trail length is counted once
        wc.end_item(flag == FLAG_OK, threadIdx.x == 0 && flag <= FLAG_CONFLICT ? nmem : 0u, threadIdx.x == 0);
    }
    wc.flush<G>(b.counters);
}

// ---------------------------------------------------------------- second / third tier: one CTA per probe, shared hash
// Probes that outgrew the warp-private state. The CTA's warps share the probe as in k_big (level-synchronous rows,
// CTA-wide closure adoption), but the set of true literals is a hash table in SHARED memory and only the trail streams
// through global memory (BhashState). With closure reuse a probe of this size is mostly adoption -- thousands of
// literals entered per step by all warps -- which is what makes a CTA per probe pay; without it (frontiers a few
// literals wide) the warp-per-probe k_deep2 was the better second tier. The finished trail is sorted in the same
// shared memory (the table is no longer needed) and leaves as a sorted set of internal literals.
template <int THREADS>
__device__ void block_bitonic_sort(uint32_t *a, uint32_t n2)
{
    BBCP_WATCH_BEGIN();
    for (uint32_t k = 2; k <= n2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            BBCP_WATCH(n2, k);
            for (uint32_t t = threadIdx.x; t < n2 / 2; t += THREADS) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const uint32_t ixj = i | j;
                const bool up = (i & k) == 0;
                const uint32_t x = a[i], y = a[ixj];
                if ((x > y) == up) { a[i] = y; a[ixj] = x; }
            }
            __syncthreads();
        }
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_block(const DbView db, const BatchView b, uint32_t *trail_slab, const uint32_t hs, const uint32_t cap, const int cls)
{
    using G = Grp<32>;
    extern __shared__ uint32_t s_hash[];
    __shared__ uint32_t ctl[12];     // 0 tail, 1 conflict, 2 overflow, 3 load result, 4 ticket, 6 whole-CTA level, 7 head, 8 second-walk entries
    __shared__ unsigned long long s_off;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t *list = cls ? b.big_list : b.mid_list;
    const uint32_t count = cls ? *b.big_count : *b.mid_count;
    if (blockIdx.x >= count) return;
    unsigned int *ticket = b.ticket + 1 + cls;
    uint32_t *next_list = cls ? b.huge_list : b.big_list, *next_count = cls ? b.huge_count : b.big_count;
    const uint8_t next_flag = cls ? FLAG_OVERFLOW3 : FLAG_OVERFLOW2;

    BhashState st;
    st.hash = s_hash;
    st.trail = trail_slab + (size_t)blockIdx.x * cap;
    st.hs = hs;
    st.cap = cap;
    for (uint32_t i = threadIdx.x; i < hs; i += THREADS) s_hash[i] = 0;
    const bool prune = !(b.opts & OPT_NO_LEN_PRUNE);
    WorkCounters wc;
    BBCP_WATCH_BEGIN();
    for (;;) {
        BBCP_WATCH(ctl[4], count);
                __syncthreads();
        if (threadIdx.x == 0) {
            ctl[4] = atomicAdd(ticket, 1u);
            ctl[0] = ctl[1] = ctl[2] = ctl[6] = ctl[8] = 0;
            ctl[3] = 0xff;
        }
        __syncthreads();
        const uint32_t t = ctl[4];
        if (t >= count) break;
        const uint64_t item = list[t];
        const Item it = get_item(b, item);
        wc.begin_item();
        if (warp == 0) {
            BlockTail tl{ctl};
            const int f = it.probe_lit ? load_probe<G>(db, st, it.probe_lit, tl) : load_cube<G>(db, st, it.lits, it.n, tl);
            if (lane == 0) ctl[3] = (uint32_t)f;
        }
        __syncthreads();
        const uint32_t f = ctl[3];
        if (f == 0xff) block_propagate(db, b, st, ctl, prune, 2 + cls, wc);
        __syncthreads();
        const bool overflow = ctl[2] != 0 || f == FLAG_OVERFLOW;
        const bool conflict = !overflow && (ctl[1] != 0 || f == FLAG_CONFLICT);
        const uint32_t tail = min(ctl[0], cap);
        const uint32_t nmem = tail - min(ctl[8], tail);     // members of the implied set (the trail also holds second-walk entries)
        uint8_t flag = (uint8_t)f;
        uint32_t len = 0;
        unsigned long long off = 0;
        if (overflow) {
            flag = next_flag;
            if (threadIdx.x == 0) next_list[atomicAdd(next_count, 1u)] = (uint32_t)item;   // queued before the flag shows
        } else if (conflict) {
            flag = FLAG_CONFLICT;
            if (it.probe_lit && threadIdx.x == 0) mark_failed(b, 2u * (uint32_t)abs(it.probe_lit) + (it.probe_lit < 0));
        } else if (f == 0xff) {
            flag = FLAG_OK;
            if (b.opts & OPT_IMPLIED) {
                // the table has done its work: sort the trail in its place
                uint32_t n2 = 1;
                while (n2 < tail) n2 <<= 1;
                for (uint32_t i = threadIdx.x; i < n2; i += THREADS) {
                    const uint32_t ent = i < tail ? st.tget(i) : TRAIL_RESCAN;
                    s_hash[i] = (ent & TRAIL_SKIP) != TRAIL_RESCAN ? (ent & TRAIL_LIT) : 0xffffffffu;   // second-walk entries sort behind the set
                }
                if (threadIdx.x == 0) {
                    const unsigned long long o = atomicAdd(b.out_cursor, (unsigned long long)nmem);
                    s_off = o + nmem <= b.out_cap ? o : ~0ull;
                }
                __syncthreads();
                if (tail > 1) block_bitonic_sort<THREADS>(s_hash, n2);
                off = s_off;
                if (off == ~0ull) { flag = FLAG_OUT_OVERFLOW; off = 0; }
                else {
                    len = nmem;
                    for (uint32_t i = threadIdx.x; i < nmem; i += THREADS) b.out_lits[off + i] = (int32_t)s_hash[i];
                    if (threadIdx.x == 0) wc.bytes += 8ull * tail;
                }
            }
        }
        __syncthreads();
        // leave the table empty for the next probe (it holds the probe's literals, or the sorted trail)
        for (uint32_t i = threadIdx.x; i < hs; i += THREADS) s_hash[i] = 0;
        if (b.reuse) __threadfence();   // every thread: its part of the set is out (publish_result, CTA-wide)
        __syncthreads();
        if (threadIdx.x == 0) {
            b.len[item] = len; b.raw_off[item] = off;
            if (len) atomicAdd(&b.tile_sum[item / SG_TILE], (unsigned long long)len);
            if (b.reuse) __threadfence();
            *(volatile uint8_t *)&b.flag[item] = flag;
            if (!overflow) atomicAdd(&b.counters[C_MID], 1ull);
            wc.n_ok += flag == FLAG_OK;
            wc.n_conf += flag == FLAG_CONFLICT;
            wc.n_skip += flag >= FLAG_TRIVIAL_L0 && flag <= FLAG_UNMAPPED;
        }
        if (!overflow) wc.end_item(flag == FLAG_OK, threadIdx.x == 0 && flag <= FLAG_CONFLICT ? nmem : 0u, threadIdx.x == 0);
    }
    wc.flush<G>(b.counters);
}

// ---------------------------------------------------------------- multi-GPU merge of the bitmap slices
// All-gather over NVLink peer memory (one process per GPU, every rank runs this on its own GPU at the same
// point of its stream). peer_base[r] = rank r's [failed | unit | control] block, mapped through CUDA IPC; rank
// r owns bitmap words [r*wp, (r+1)*wp). PULL protocol, no system-scope fence anywhere (measured here: a
// fence.sys behind remote stores costs ~10 us, a remote load round trip ~2 us):
//   signal  as soon as this rank's bitmaps are final (they were written by EARLIER kernels of the stream, so
//           they are already performed in this GPU's L2, which also serves the peers' reads) one thread stores
//           the merge epoch into this rank's flag word in every peer;
//   wait    before reading, a thread spins on the local flag words until every peer's epoch has arrived;
//   pull    every thread then copies its share of every peer's slice of both bitmaps with 16-byte remote loads.
// Inside a batch the signal is the first thing k_scan_gather does and wait + pull are the last, so the flags
// travel while the compaction runs; k_merge is the stand-alone form.
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// control words behind the two bitmaps (MERGE_CTRL_*): [0..63] arrival epoch per rank ("rank r's slice is final"),
// [64] CTAs of the running merge that are done pulling (local), [65] time-out latch, [128..191] pull-done epoch per rank
// ("rank r has finished reading my slice": what bbcp_clear_bitmaps / a re-finalize / a detach wait for before they
// touch the bitmaps, k_merge_drain)
// slice_empty: this rank's slice holds no failed literal (the usual case on real instances): the peers are told so and
// clear their copy locally instead of pulling it. The hint travels as a second word; if it lands after the arrival
// flag has been seen the peer just pulls the (all-zero) slice -- slower, never wrong.
__device__ __forceinline__ void merge_signal(const MergeArgs &m, bool slice_empty)
{
    for (int p = 0; p < m.world; ++p) {
        if (p == m.rank) continue;
        uint32_t *ctrl = m.peer_base[p] + 2 * m.stride;
        if (slice_empty) *(volatile uint32_t *)(ctrl + MERGE_CTRL_EMPTY + m.rank) = m.epoch;
        *(volatile uint32_t *)(ctrl + m.rank) = m.epoch;
    }
}

// spin until every peer's word [base + p] has reached the epoch; one thread; false on time-out (a peer died or never
// made the matching call; the host call reports it)
__device__ bool merge_spin(const MergeArgs &m, int base)
{
    uint32_t *ctrl = m.peer_base[m.rank] + 2 * m.stride;
    const unsigned long long t0 = global_ns(), limit = (unsigned long long)m.timeout_ms * 1000000ull;
    for (int p = 0; p < m.world; ++p) {
        if (p == m.rank) continue;
        uint32_t spins = 0;
        while ((int32_t)(*(volatile uint32_t *)&ctrl[base + p] - m.epoch) < 0) {
            if ((++spins & 0xfffu) == 0 && global_ns() - t0 > limit) { ctrl[MERGE_CTRL_ERROR] = 1; return false; }
        }
    }
    __threadfence();   // the loads of the pull stay behind the flag reads
    return true;
}

__device__ __forceinline__ bool merge_wait(const MergeArgs &m) { return merge_spin(m, 0); }

// the same wait by a whole warp: lane p (and p + 32) watches peer p's flag, so eight peers cost one L2 round trip, not eight
__device__ bool merge_wait_warp(const MergeArgs &m)
{
    uint32_t *ctrl = m.peer_base[m.rank] + 2 * m.stride;
    const int lane = lane_id();
    const unsigned long long t0 = global_ns(), limit = (unsigned long long)m.timeout_ms * 1000000ull;
    bool ok = true;
    for (int p = lane; p < m.world; p += 32) {
        if (p == m.rank) continue;
        uint32_t spins = 0;
        while ((int32_t)(*(volatile uint32_t *)&ctrl[p] - m.epoch) < 0) {
            if ((++spins & 0xfffu) == 0 && global_ns() - t0 > limit) { ctrl[MERGE_CTRL_ERROR] = 1; ok = false; break; }
        }
    }
    ok = __all_sync(FULL, ok);
    __threadfence();   // the loads of the pull stay behind the flag reads
    return ok;
}

// tells every peer that this rank has finished reading their slices (called once, by the last CTA of a merging kernel)
__device__ __forceinline__ void merge_done_signal(const MergeArgs &m)
{
    for (int p = 0; p < m.world; ++p)
        if (p != m.rank) *(volatile uint32_t *)(m.peer_base[p] + 2 * m.stride + MERGE_CTRL_DONE + m.rank) = m.epoch;
}

// called by thread 0 of every CTA of a merging kernel when the CTA's share of the pull is complete (its remote loads
// have returned: their values have been stored). The last CTA tells every peer that this rank is done reading.
__device__ __forceinline__ void merge_cta_done(const MergeArgs &m)
{
    uint32_t *ctrl = m.peer_base[m.rank] + 2 * m.stride;
    __threadfence();
    if (atomicAdd(&ctrl[MERGE_CTRL_FINISHED], 1u) == gridDim.x - 1) {
        ctrl[MERGE_CTRL_FINISHED] = 0;   // the next merging kernel of the stream starts after this one has drained
        for (int p = 0; p < m.world; ++p)
            if (p != m.rank) *(volatile uint32_t *)(m.peer_base[p] + 2 * m.stride + MERGE_CTRL_DONE + m.rank) = m.epoch;
    }
}

__device__ __forceinline__ uint4 ld_remote(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// element i of the pull (0 <= i < merge_total): 16 bytes of a peer's slice of the FAILED bitmap; the unit bitmap is its
// image under "negate the literal" = swap the two bits of every variable, so it is derived here instead of being pulled
__device__ __forceinline__ uint32_t swap_polarity(uint32_t f) { return ((f & 0x55555555u) << 1) | ((f & 0xaaaaaaaau) >> 1); }

__device__ __forceinline__ uint64_t merge_total(const MergeArgs &m) { return (m.wp / 4) * (uint64_t)(m.world - 1); }

__device__ __forceinline__ void merge_pull(const MergeArgs &m, uint64_t gtid, uint64_t gsize)
{
    const uint64_t wp4 = m.wp / 4, s4 = m.stride / 4;           // both multiples of 4 words
    const uint64_t total = merge_total(m);
    uint4 *mine = reinterpret_cast<uint4 *>(m.peer_base[m.rank]);
    const volatile uint32_t *empty = m.peer_base[m.rank] + 2 * m.stride + MERGE_CTRL_EMPTY;
    for (uint64_t i = gtid; i < total; i += gsize) {
        const int pi = (int)(i / wp4);
        const int p = pi + (pi >= m.rank);
        const uint64_t v = (uint64_t)p * wp4 + (i - (uint64_t)pi * wp4);   // vector index inside the bitmap
        if (v >= s4) continue;                                             // the last rank's slice may be short
        const uint4 f = empty[p] == m.epoch ? make_uint4(0, 0, 0, 0) : ld_remote(reinterpret_cast<const uint4 *>(m.peer_base[p]) + v);
        mine[v] = f;
        mine[s4 + v] = make_uint4(swap_polarity(f.x), swap_polarity(f.y), swap_polarity(f.z), swap_polarity(f.w));
    }
}

__global__ void __launch_bounds__(256) k_merge(const MergeArgs m)
{
    __shared__ int s_ok;
    if (blockIdx.x == 0 && threadIdx.x == 0) merge_signal(m, false);
    if (threadIdx.x == 0) s_ok = merge_wait(m);
    __syncthreads();
    if (s_ok && m.wp) merge_pull(m, (uint64_t)blockIdx.x * 256 + threadIdx.x, (uint64_t)gridDim.x * 256);
    __syncthreads();
    if (threadIdx.x == 0) merge_cta_done(m);
}

// waits until every peer has finished reading this rank's slice in the merge of epoch m.epoch
__global__ void k_merge_drain(const MergeArgs m)
{
    if (threadIdx.x == 0) merge_spin(m, MERGE_CTRL_DONE);
}

// ---------------------------------------------------------------- result compaction into canonical CSR order
// One pass: sizes -> exclusive prefix sum (every CTA derives its first offset from the per-tile sums, see below)
// -> CSR offsets written, every implied set copied from its allocation-order position in the pool to its
// canonical position, self-implied literals regenerated from the item index. Sets above GATHER_GIANT are cut
// into chunks for k_gather_long so that no CTA is stuck with one of them.
constexpr int SG_THREADS = 256;
constexpr int SG_ITEMS = 8;
static_assert(SG_THREADS * SG_ITEMS == SG_TILE, "tile shape");
constexpr uint32_t GATHER_WARP = 8;       // sets up to this size are copied by the owning thread
constexpr uint32_t GATHER_BLOCK = 1024;   // up to this by a warp, above by the whole CTA
constexpr uint32_t GATHER_GIANT = 16384;  // above: chunked through k_gather_long
constexpr uint32_t GATHER_CHUNK = 2048;

__device__ __forceinline__ int32_t item_self_lit(const BatchView &b, uint64_t item)
{
    if (b.cube_off) return b.cube_lits[b.cube_off[item]];
    if (b.cube_lits) return b.cube_lits[item];
    const uint64_t idx = b.first + item;
    const int32_t pv = (int32_t)(idx / 2 + 1);
    return (idx & 1) ? -pv : pv;
}

// The items are cut into one contiguous segment of whole 2048-item tiles per CTA. The sum of the set sizes of
// every tile is already known when this kernel starts (k_triage stores it, the propagate kernels add to it), so a
// CTA gets the offset of its first item by adding up the tile sums before its segment (a few hundred to a few
// thousand 8-byte loads spread over 256 threads) -- no pass over the sizes, no grid barrier, no look-back chain.
// It then walks its segment, 2048 items per step in a warp-striped arrangement -- lane l of warp w owns items
// w*256 + j*32 + l, j = 0..7, so every load and store of a warp is one contiguous 128/256-byte run -- turns sizes
// into offsets with warp shuffles, writes the offsets and moves the literals.
template <bool OFF32, bool MERGE>
__global__ void __launch_bounds__(SG_THREADS) k_scan_gather(const BatchView b, const int pass, const uint64_t seg, const MergeArgs mg)
{
    pdl_wait();
    // first pass of a batch: if some probes still wait for the later tiers the sizes are not final
    if (pass == 0 && *(volatile uint32_t *)b.mid_count != 0) return;
    // the bitmaps are final (every propagate kernel of the batch is done): start sending this rank's slice to the
    // peers now -- unless the implied-literal pool overflowed and the whole batch is about to be re-run
    // (MERGE is a template parameter so that the single-GPU kernel does not carry the merge's registers)
    const bool merging = MERGE && mg.world > 1 && *(volatile unsigned long long *)b.out_cursor <= b.out_cap;
    if (merging && blockIdx.x == 0 && threadIdx.x == 0)
        merge_signal(mg, !mg.dirty_before && ((volatile unsigned long long *)b.counters)[C_CONFLICT] == 0ull);
    __shared__ unsigned long long s_warp[SG_THREADS / 32], s_red[SG_THREADS / 32];
    __shared__ uint32_t s_nbig, s_pos;
    __shared__ uint16_t s_big[SG_TILE];
    __shared__ unsigned long long s_bigdst[SG_TILE];
    const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const uint64_t n = b.n;
    constexpr bool off32 = OFF32;
    int32_t *__restrict__ dst = b.impl;
    const uint64_t lo = min((unsigned long long)blockIdx.x * seg, (unsigned long long)n), hi = min((unsigned long long)(lo + seg), (unsigned long long)n);

    // my prefix = the sizes of all items before my segment = sum of the 2048-item tile sums before it. The tile sums
    // were initialised by k_triage (one store per tile) and grown by the propagate kernels (one atomicAdd per stored
    // set), so no pass over the sizes and no grid barrier is needed here.
    unsigned long long pre = 0;
    {
        const volatile unsigned long long *ts = b.tile_sum;
        const uint32_t tiles_before = (uint32_t)(lo / SG_TILE);
        for (uint32_t k = tid; k < tiles_before; k += SG_THREADS) pre += ts[k];
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) pre += __shfl_xor_sync(FULL, pre, d);
    if (lane == 0) s_red[warp] = pre;
    __syncthreads();
    unsigned long long run = 0;   // offset of the first item of the current step
#pragma unroll
    for (int k = 0; k < SG_THREADS / 32; ++k) run += s_red[k];
    if (blockIdx.x == gridDim.x - 1 && lo >= hi && tid == 0) {
        // degenerate tail (n == 0, or n a multiple of seg with spare CTAs): the last CTA still owes off[n]
        if (off32) ((uint32_t *)b.off)[n] = (uint32_t)(run + b.off_base); else ((uint64_t *)b.off)[n] = run + b.off_base;
        *b.total = run;
    }

    for (uint64_t cbase = lo; cbase < hi; cbase += SG_TILE) {
        if (tid == 0) s_nbig = 0;
        const uint64_t wbase = cbase + (uint64_t)warp * (32 * SG_ITEMS) + lane;
        const uint32_t nvalid = (uint32_t)min((unsigned long long)(32 * SG_ITEMS), (unsigned long long)(hi > wbase - lane ? hi - (wbase - lane) : 0));
        uint32_t len[SG_ITEMS];
#pragma unroll
        for (int j = 0; j < SG_ITEMS; ++j) len[j] = (uint32_t)(j * 32 + lane) < nvalid ? b.len[wbase + j * 32] : 0u;
        // sizes -> offsets inside the warp's 256 items. 32-bit shuffles unless some set of the step is huge
        // (then the row sums could leave 32 bits)
        uint32_t mx = 0;
#pragma unroll
        for (int j = 0; j < SG_ITEMS; ++j) mx = max(mx, len[j] & 0x7fffffffu);
        unsigned long long o[SG_ITEMS];
        unsigned long long rowbase = 0;
        if (!__any_sync(FULL, mx >= (1u << 20))) {
            uint32_t rb = 0;
#pragma unroll
            for (int j = 0; j < SG_ITEMS; ++j) {
                const uint32_t v = len[j] & 0x7fffffffu;
                if (__all_sync(FULL, v == 1u)) {   // a row of one-literal sets (self-implied probes): no scan needed
                    o[j] = rb + lane;
                    rb += 32;
                    continue;
                }
                uint32_t incl = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += t;
                }
                o[j] = rb + incl - v;
                rb += __shfl_sync(FULL, incl, 31);
            }
            rowbase = rb;
        } else {
#pragma unroll
            for (int j = 0; j < SG_ITEMS; ++j) {
                const unsigned long long v = len[j] & 0x7fffffffu;
                unsigned long long incl = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned long long t = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += t;
                }
                o[j] = rowbase + incl - v;
                rowbase += __shfl_sync(FULL, incl, 31);
            }
        }
        if (lane == 0) s_warp[warp] = rowbase;
        __syncthreads();
        unsigned long long before = 0, agg = 0;
#pragma unroll
        for (int k = 0; k < SG_THREADS / 32; ++k) { const unsigned long long x = s_warp[k]; agg += x; if (k < warp) before += x; }
        const unsigned long long wrun = run + before;
#pragma unroll
        for (int j = 0; j < SG_ITEMS; ++j) {
            o[j] += wrun;
            if ((uint32_t)(j * 32 + lane) < nvalid) {
                if (OFF32) ((uint32_t *)b.off)[wbase + j * 32] = (uint32_t)(o[j] + b.off_base); else ((uint64_t *)b.off)[wbase + j * 32] = o[j] + b.off_base;
            }
        }
        run += agg;
        if (cbase + SG_TILE >= hi && hi == n && tid == 0) {   // the step that holds the last item: off[n], total
            if (OFF32) ((uint32_t *)b.off)[n] = (uint32_t)(run + b.off_base); else ((uint64_t *)b.off)[n] = run + b.off_base;
            *b.total = run;
        }
        // literals
#pragma unroll
        for (int j = 0; j < SG_ITEMS; ++j) {
            const uint64_t item = wbase + j * 32;
            uint32_t l = len[j];
            if (l & LEN_SELF) { dst[o[j]] = item_self_lit(b, item); l = 0; }
            if (!__any_sync(FULL, l != 0)) continue;   // nothing stored for this row (the common case in a triage-only batch)
            uint64_t src = 0;
            if (l) src = b.raw_off[item];
            if (l && l <= GATHER_WARP) {
                for (uint32_t i = 0; i < l; ++i) dst[o[j] + i] = to_ext((uint32_t)b.out_lits[src + i]);
            }
            unsigned med = __ballot_sync(FULL, l > GATHER_WARP && l <= GATHER_BLOCK);
            while (med) {
                const int k = __ffs(med) - 1;
                med &= med - 1;
                const uint32_t ll = __shfl_sync(FULL, l, k);
                const uint64_t ss = __shfl_sync(FULL, src, k), dd = __shfl_sync(FULL, o[j], k);
                for (uint32_t i = lane; i < ll; i += 32) dst[dd + i] = to_ext((uint32_t)b.out_lits[ss + i]);
            }
            if (l > GATHER_BLOCK) {
                const uint32_t e = atomicAdd(&s_nbig, 1u);
                s_big[e] = (uint16_t)(warp * (32 * SG_ITEMS) + j * 32 + lane);
                s_bigdst[e] = o[j];
            }
        }
        __syncthreads();
        const uint32_t nbig = s_nbig;
        for (uint32_t e = 0; e < nbig; ++e) {
            const uint64_t item = cbase + s_big[e];
            const uint32_t l = b.len[item] & 0x7fffffffu;
            if (l > GATHER_GIANT) {
                const uint32_t nch = (l + GATHER_CHUNK - 1) / GATHER_CHUNK;
                if (tid == 0) s_pos = atomicAdd(b.chunk_count, nch);
                __syncthreads();
                const uint32_t pos = s_pos;
                for (uint32_t c = tid; c < nch; c += SG_THREADS) b.chunk_list[pos + c] = make_uint2((uint32_t)item, c);
                __syncthreads();
                continue;
            }
            const uint64_t ss = b.raw_off[item], dd = s_bigdst[e];
            for (uint32_t i = tid; i < l; i += SG_THREADS) dst[dd + i] = to_ext((uint32_t)b.out_lits[ss + i]);
        }
        if (nbig) __syncthreads();
    }
    // merge: by now the peers' flags have had a whole compaction to arrive; copy their slices
    if (merging) {
        // (measured and dropped: issuing the remote loads early and holding the values to the end -- +16 registers, one
        // CTA less per SM, no gain; letting only the last-dispatched CTAs pull -- slower as well)
        __syncthreads();
        if (warp == 0) { const bool ok = merge_wait_warp(mg); if (lane == 0) s_pos = ok; }
        __syncthreads();
        if (s_pos) merge_pull(mg, (uint64_t)blockIdx.x * SG_THREADS + tid, (uint64_t)gridDim.x * SG_THREADS);
    }
    // the last CTA to finish copies the counter block of this launch sequence into mapped host memory: the host
    // reads it after the stream synchronisation it does anyway (no copy-engine round trip, and no fence: the
    // kernel has completed by then)
    // (a merging kernel counts its CTAs in as well: the last one tells the peers that this rank is done reading their slices --
    // its own remote loads have returned, their values are stored, and so have those of every CTA counted before it)
    if (b.host_report || merging) {
        __syncthreads();   // this CTA's stores (offsets, total) are issued; warp 0 alone counts the CTA in, the rest leaves
        if (warp == 0) {
            unsigned last = 0;
            if (lane == 0) {
                __threadfence();
                last = atomicAdd(b.tile_ticket + 1, 1u) == gridDim.x - 1;
            }
            if (__shfl_sync(FULL, last, 0)) {
                if (merging && lane == 0) merge_done_signal(mg);
                if (b.host_report) {
                    __threadfence();
                    for (uint32_t i = lane; i < b.zero_words; i += 32) b.host_report[i] = ((volatile unsigned long long *)b.counters)[i];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_gather_long(const BatchView b)
{
    const uint32_t nchunks = *b.chunk_count;
    const bool off32 = (b.opts & OPT_OFF32) != 0;
    for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint2 e = b.chunk_list[c];
        const uint32_t len = b.len[e.x] & 0x7fffffffu;
        const uint64_t o = (off32 ? (uint64_t)((const uint32_t *)b.off)[e.x] : ((const uint64_t *)b.off)[e.x]) - (off32 ? (uint64_t)(uint32_t)b.off_base : b.off_base);
        const uint64_t src = b.raw_off[e.x] + (uint64_t)e.y * GATHER_CHUNK, d = o + (uint64_t)e.y * GATHER_CHUNK;
        const uint32_t m = min(GATHER_CHUNK, len - e.y * GATHER_CHUNK);
#pragma unroll
        for (uint32_t k = 0; k < GATHER_CHUNK / 256; ++k) {
            const uint32_t i = k * 256 + threadIdx.x;
            if (i < m) b.impl[d + i] = to_ext((uint32_t)b.out_lits[src + i]);
        }
    }
}

// ---------------------------------------------------------------- probing products (survey "next" row N4)
// From the implied sets of one all-literal pass (item 2k = +v, item 2k+1 = -v, both sorted by variable):
//   necessary literal  x in impl(+v) and x in impl(-v): x holds whichever way v is decided, so x is a unit;
//   equivalence        x in impl(+v) and not-x in impl(-v), x != v: v -> x and not-v -> not-x, so v == x.
// One warp per variable: the lanes walk the shorter set and binary-search the longer one.
template <bool OFF32>
__global__ void __launch_bounds__(256) k_products(const uint8_t *__restrict__ flag, const void *__restrict__ off_, const int32_t *__restrict__ impl,
                                                   const uint64_t npairs, const uint64_t var_base, uint32_t *__restrict__ necessary, int32_t *__restrict__ equiv,
                                                   const unsigned long long equiv_cap, unsigned long long *__restrict__ counts)
{
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * 256) >> 5;
    const int lane = lane_id();
    for (uint64_t v = warp; v < npairs; v += nwarps) {
        if (flag[2 * v] != FLAG_OK || flag[2 * v + 1] != FLAG_OK) continue;
        uint64_t o0, o1, o2;
        if (OFF32) { const uint32_t *o = (const uint32_t *)off_; o0 = o[2 * v]; o1 = o[2 * v + 1]; o2 = o[2 * v + 2]; }
        else { const uint64_t *o = (const uint64_t *)off_; o0 = o[2 * v]; o1 = o[2 * v + 1]; o2 = o[2 * v + 2]; }
        const int32_t *A = impl + o0, *B = impl + o1;
        uint64_t na = o1 - o0, nb = o2 - o1;
        const int32_t var = (int32_t)(var_base + v + 1);   // the variable this pair of probes belongs to
        const bool swap = na > nb;
        if (swap) { const int32_t *t = A; A = B; B = t; const uint64_t tn = na; na = nb; nb = tn; }
        for (uint64_t i = lane; i < na; i += 32) {
            const int32_t x = A[i];
            const int32_t ax = x < 0 ? -x : x;
            uint64_t lo = 0, hi = nb;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                const int32_t y = B[mid];
                if ((y < 0 ? -y : y) < ax) lo = mid + 1; else hi = mid;
            }
            if (lo >= nb) continue;
            const int32_t y = B[lo];
            if ((y < 0 ? -y : y) != ax) continue;
            if (y == x) {
                const uint32_t il = 2u * (uint32_t)ax + (x < 0);
                if (!(atomicOr(&necessary[il >> 5], 1u << (il & 31)) & (1u << (il & 31)))) atomicAdd(&counts[0], 1ull);
            } else {
                // opposite signs. x_pos = the literal implied by +v (A and B may have been swapped)
                const int32_t x_pos = swap ? y : x;
                if (ax == var) continue;                 // v itself: +v in impl(+v), -v in impl(-v)
                const unsigned long long e = atomicAdd(&counts[1], 1ull);
                if (e < equiv_cap) { equiv[2 * e] = var; equiv[2 * e + 1] = x_pos; }
            }
        }
    }
}

// ---------------------------------------------------------------- fixpoint checker (diagnostic)
// The static analogue of the reference's WLAssertNoMissedImplications (TopiWL.cc:330-442): for the items of the last
// batch that reached a fixpoint, walk EVERY row of the database under the item's implied set (+ level 0) and count the
// clauses that are not satisfied and have fewer than two unassigned literals -- a missed implication or a missed
// conflict. One CTA per item, the assignment as a dense 2-bit-per-variable scratch bitmap. O(database) per item: a test
// and debugging aid (bbcp_check_fixpoints), not part of any batch.
template <bool OFF32>
__global__ void __launch_bounds__(256) k_check_fixpoints(const DbView db, const uint8_t *__restrict__ flag, const void *__restrict__ off_, const int32_t *__restrict__ impl,
                                                        const uint64_t first, const uint64_t count, const uint64_t stride, uint32_t *__restrict__ scratch,
                                                        unsigned long long *__restrict__ out)
{
    uint32_t *bits = scratch + (size_t)blockIdx.x * db.nvars_words16;
    const uint32_t nl = 2u * ((uint32_t)db.max_var + 1u);
    auto val = [&](uint32_t lit) -> int {      // 0 unassigned, 1 true, 2 false
        const uint32_t var = lit >> 1;
        const uint32_t o = (bits[var >> 4] >> ((var & 15u) * 2u)) & 3u;
        if (o) return (o == 1u) == ((lit & 1u) == 0u) ? 1 : 2;
        const uint8_t li = db.litinfo[lit];
        return (li & LI_TRUE_L0) ? 1 : (li & LI_FALSE_L0) ? 2 : 0;
    };
    for (uint64_t k = blockIdx.x; k < count; k += gridDim.x) {
        const uint64_t item = first + k * stride;
        if (flag[item] != FLAG_OK) continue;
        uint64_t o0, o1;
        if (OFF32) { o0 = ((const uint32_t *)off_)[item]; o1 = ((const uint32_t *)off_)[item + 1]; }
        else { o0 = ((const uint64_t *)off_)[item]; o1 = ((const uint64_t *)off_)[item + 1]; }
        for (uint64_t i = o0 + threadIdx.x; i < o1; i += 256) {
            const int32_t l = impl[i];
            const uint32_t var = (uint32_t)(l < 0 ? -l : l);
            atomicOr(&bits[var >> 4], (l < 0 ? 2u : 1u) << ((var & 15u) * 2u));
        }
        __syncthreads();
        unsigned long long bad = 0, seen = 0;
        for (uint32_t l = 2 + threadIdx.x; l < nl; l += 256) {
            const uint4 hd = db.hdr[l];
            if (!(hd.y | hd.z | hd.w)) continue;
            const int vl = val(l);
            const uint2 *row = db.slots + hd.x;
            for (uint32_t j = 0; j < hd.y; ++j) {                         // binary clauses (l, o)
                const uint32_t o = (j & 1) ? row[j / 2].y : row[j / 2].x;
                const int vo = val(o);
                ++seen;
                bad += vl != 1 && vo != 1 && (vl == 2 || vo == 2);
            }
            row += (hd.y + 1) / 2;
            for (uint32_t j = 0; j < hd.z; ++j) {                         // ternary clauses (l, a, b)
                const int va = val(row[j].x), vb = val(row[j].y);
                ++seen;
                const int nfree = (vl == 0) + (va == 0) + (vb == 0);
                bad += vl != 1 && va != 1 && vb != 1 && nfree < 2;
            }
            row += hd.z;
            for (uint32_t j = 0; j < hd.w; ++j) {                         // long clauses from the arena
                const uint32_t *c = reinterpret_cast<const uint32_t *>(db.arena + row[j].y);
                const uint32_t size = c[0];
                int nfree = 0;
                bool sat = false;
                for (uint32_t q = 1; q <= size; ++q) { const int v = val(c[q]); sat |= v == 1; nfree += v == 0; }
                ++seen;
                bad += !sat && nfree < 2;
            }
        }
        if (bad) atomicAdd(&out[0], bad);
        atomicAdd(&out[1], seen);
        if (threadIdx.x == 0) atomicAdd(&out[2], 1ull);
        __syncthreads();
        for (uint64_t i = o0 + threadIdx.x; i < o1; i += 256) {
            const int32_t l = impl[i];
            const uint32_t var = (uint32_t)(l < 0 ? -l : l);
            atomicAnd(&bits[var >> 4], ~(3u << ((var & 15u) * 2u)));
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- unit feedback on the device (survey "next" row N2)
__global__ void __launch_bounds__(256) k_mark_level0(uint8_t *__restrict__ litinfo, const int32_t *__restrict__ lits, const uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const int32_t l = lits[i];
        const uint32_t il = 2u * (uint32_t)(l < 0 ? -l : l) + (l < 0);
        litinfo[il] = (litinfo[il] & ~(LI_TRUE_L0 | LI_FALSE_L0)) | LI_TRUE_L0;          // one byte per literal, one literal per variable: no races
        litinfo[il ^ 1u] = (litinfo[il ^ 1u] & ~(LI_TRUE_L0 | LI_FALSE_L0)) | LI_FALSE_L0;
    }
}

// ---------------------------------------------------------------- probe schedule for closure reuse, built on the device
// height(p) = length of the longest chain of binary implications that starts at p (p -> q when the clause (not-p or q)
// exists), capped at 255: round r of the relaxation below holds min(height, r), so an acyclic implication graph is done
// after depth+1 rounds and literals on or above a cycle (equivalent literals) end at the cap. The schedule is the
// literals counting-sorted by that height: children first wherever the graph is acyclic, which is all that closure reuse
// wants (the results never depend on the order). HostDb::build_order is the exact host form (Tarjan); at 50 M clauses it
// costs half a minute of one host core, this costs milliseconds.
__global__ void __launch_bounds__(256) k_height_round(const DbView db, const uint8_t *__restrict__ hin, uint8_t *__restrict__ hout, const uint32_t nl,
                                                      uint32_t *__restrict__ changed)
{
    bool any = false;
    for (uint32_t l = blockIdx.x * 256 + threadIdx.x; l < nl; l += gridDim.x * 256) {
        uint32_t h = 0;
        if (l >= 2) {
            const uint4 hd = __ldg(&db.hdr[l ^ 1u]);
            for (uint32_t j = 0; j < hd.y; j += 2) {
                const uint2 s = __ldg(&db.slots[hd.x + j / 2]);
                h = max(h, (uint32_t)hin[s.x] + 1u);
                if (j + 1 < hd.y) h = max(h, (uint32_t)hin[s.y] + 1u);
            }
            h = min(h, 255u);
        }
        any |= h != hin[l];
        hout[l] = (uint8_t)h;
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) *changed = 1u;
}

__global__ void __launch_bounds__(256) k_order_hist(const uint8_t *__restrict__ h, const uint32_t nl, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t l = 2 + blockIdx.x * 256 + threadIdx.x; l < nl; l += gridDim.x * 256) atomicAdd(&s_h[h[l]], 1u);
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_order_scan(uint32_t *hist)   // hist[256] -> exclusive prefix sums, in place
{
    __shared__ uint32_t s[256];
    s[threadIdx.x] = hist[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t run = 0; for (int i = 0; i < 256; ++i) { const uint32_t c = s[i]; s[i] = run; run += c; } }
    __syncthreads();
    hist[threadIdx.x] = s[threadIdx.x];
}

__global__ void __launch_bounds__(256) k_order_scatter(const uint8_t *__restrict__ h, const uint32_t nl, uint32_t *__restrict__ cursor, uint32_t *__restrict__ order)
{
    // every warp walks a contiguous run of literals, so a height class keeps (roughly) the literal order
    const uint32_t per_warp = ((nl + gridDim.x * 8 - 1) / (gridDim.x * 8) + 31) & ~31u;
    const uint32_t w = blockIdx.x * 8 + (threadIdx.x >> 5);
    const uint32_t lo = w * per_warp, hi = min(lo + per_warp, nl);
    for (uint32_t base = lo; base < hi; base += 32) {
        const uint32_t l = base + lane_id();
        const bool ok = l >= 2 && l < hi;
        const uint32_t key = ok ? h[l] : 0x100u + lane_id();
        const unsigned m = __match_any_sync(FULL, key);
        const int leader = __ffs(m) - 1;
        uint32_t pos = 0;
        if (ok && lane_id() == leader) pos = atomicAdd(&cursor[key], (uint32_t)__popc(m));
        pos = __shfl_sync(FULL, pos, leader) + __popc(m & lanemask_lt());
        if (ok) order[pos] = l;
    }
}

// ---------------------------------------------------------------- hyper-binary resolvents (survey "next" row N4)
// x is implied by probe v, but not through binary clauses alone: (not-v or x) is a hyper-binary resolvent -- a binary
// clause the database does not entail by its binary clauses, learnt for free from the probe. full = the implied sets of
// an all-literal pass, bin = the same pass restricted to the binary clauses (OPT_BINARY_ONLY); both sorted by variable,
// bin[i] a subset of full[i]. One warp per probe: the lanes take the literals of full[i] and binary-search bin[i].
__global__ void __launch_bounds__(256) k_hbr(const uint8_t *__restrict__ flag_full, const uint64_t *__restrict__ off_full, const int32_t *__restrict__ impl_full,
                                             const uint8_t *__restrict__ flag_bin, const uint64_t *__restrict__ off_bin, const int32_t *__restrict__ impl_bin,
                                             const uint64_t n, const uint64_t first, int32_t *__restrict__ pairs, const unsigned long long cap,
                                             unsigned long long *__restrict__ count)
{
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * 256) >> 5;
    const int lane = lane_id();
    for (uint64_t i = warp; i < n; i += nwarps) {
        if (flag_full[i] != FLAG_OK || flag_bin[i] != FLAG_OK) continue;
        const int32_t *A = impl_full + off_full[i], *B = impl_bin + off_bin[i];
        const uint64_t na = off_full[i + 1] - off_full[i], nb = off_bin[i + 1] - off_bin[i];
        if (na == nb) continue;
        const uint64_t idx = first + i;
        const int32_t pv = (int32_t)(idx / 2 + 1), probe = (idx & 1) ? -pv : pv;
        for (uint64_t k = lane; k < na; k += 32) {
            const int32_t x = A[k];
            const int32_t ax = x < 0 ? -x : x;
            uint64_t lo = 0, hi = nb;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                const int32_t y = B[mid];
                if ((y < 0 ? -y : y) < ax) lo = mid + 1; else hi = mid;
            }
            if (lo < nb && B[lo] == x) continue;
            const unsigned long long e = atomicAdd(count, 1ull);
            if (e < cap) { pairs[2 * e] = probe; pairs[2 * e + 1] = x; }
        }
    }
}

}  // namespace

void launch_hbr(const uint8_t *flag_full, const uint64_t *off_full, const int32_t *impl_full, const uint8_t *flag_bin, const uint64_t *off_bin, const int32_t *impl_bin,
                uint64_t n, uint64_t first, int32_t *pairs, unsigned long long cap, unsigned long long *count, int sm_count, cudaStream_t s)
{
    const int blocks = (int)std::min<uint64_t>(std::max<uint64_t>((n + 7) / 8, 1), (uint64_t)sm_count * 8);
    k_hbr<<<blocks, 256, 0, s>>>(flag_full, off_full, impl_full, flag_bin, off_bin, impl_bin, n, first, pairs, cap, count);
}

// d_height: 2 * nl bytes + 1028 bytes of scratch (ping-pong heights, changed flag, histogram); d_order: nl - 2 words
void build_order_device(const DbView &db, uint32_t nl, uint8_t *d_height, uint32_t *d_order, int sm_count, cudaStream_t s)
{
    uint8_t *h0 = d_height, *h1 = d_height + nl;
    uint32_t *changed = reinterpret_cast<uint32_t *>(d_height + 2 * (size_t)((nl + 3) & ~3u));
    uint32_t *hist = changed + 1;
    cudaMemsetAsync(d_height, 0, 2 * (size_t)((nl + 3) & ~3u) + 4 + 1024, s);
    const int blocks = sm_count * 8;
    uint32_t flag = 1;
    for (int round = 0; round < 256 && flag; round += 8) {
        cudaMemsetAsync(changed, 0, 4, s);
        for (int k = 0; k < 8; ++k) {
            k_height_round<<<blocks, 256, 0, s>>>(db, h0, h1, nl, changed);
            std::swap(h0, h1);
        }
        cudaMemcpyAsync(&flag, changed, 4, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
    }
    k_order_hist<<<blocks, 256, 0, s>>>(h0, nl, hist);
    k_order_scan<<<1, 256, 0, s>>>(hist);
    k_order_scatter<<<blocks, 256, 0, s>>>(h0, nl, hist, d_order);
}

static uint32_t ilog2(uint32_t x) { uint32_t l = 0; while ((1u << l) < x) ++l; return l; }

// first tier: (hash 2*cap + trail cap + slot indices cap/2) words per probe, SMALL_THREADS / W probes per CTA
static bool stage_rows_enabled() { static const bool on = getenv("BBCP_TMA_ROWS") != nullptr && atoi(getenv("BBCP_TMA_ROWS")) != 0; return on; }
size_t small_smem_bytes(uint32_t cap, int width)
{
    return (size_t)(SMALL_THREADS / width) * (2 * cap + cap + cap / 2 + (stage_rows_enabled() ? 2 * (STAGE_SLOTS + 2) + 4 : 0)) * 4;
}

template <int W> static int small_bps_w(uint32_t cap)
{
    int nb = 0;
    const size_t smem = small_smem_bytes(cap, W);
    if (cudaFuncSetAttribute(k_deep<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_deep<W>, SMALL_THREADS, smem);
    return nb;
}

int small_blocks_per_sm(uint32_t cap, int width) { return small_bps_w<32>(cap); }

void launch_triage(const BatchView &b, const DbView &db, int sm_count, cudaStream_t s)
{
    const int blocks = (int)std::min<uint64_t>((b.n + 256 * TRIAGE_ITEMS - 1) / (256 * TRIAGE_ITEMS), (uint64_t)sm_count * 8);
    if (b.cube_off) k_triage<TM_CUBES><<<blocks, 256, 0, s>>>(db, b);
    else if (b.cube_lits) k_triage<TM_LIST><<<blocks, 256, 0, s>>>(db, b);
    else k_triage<TM_UNIVERSE><<<blocks, 256, 0, s>>>(db, b);
}

// launch with programmatic stream serialization: the kernel may be scheduled before its predecessor in the stream has
// drained; it calls pdl_wait() before it touches anything the predecessor wrote
template <class... KArgs, class... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

void launch_small(const DbView &db, const BatchView &b, uint32_t cap, int width, int blocks, cudaStream_t s, bool pdl)
{
    const size_t smem = small_smem_bytes(cap, width);
    const uint32_t lh = ilog2(2 * cap);
    launch_pdl(k_deep<32>, dim3(blocks), dim3(SMALL_THREADS), smem, s, pdl, db, b, lh, cap, stage_rows_enabled() ? 1u : 0u);
}

size_t mid_slab_words_per_group(uint32_t cap) { return (size_t)2 * cap + cap; }

int mid_blocks_per_sm(int width)
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_deep2<32>, MID2_THREADS, 0);
    return nb;
}

void launch_mid(const DbView &db, const BatchView &b, uint32_t *slab, uint32_t cap, int width, int blocks, cudaStream_t s)
{
    const uint32_t lh = ilog2(2 * cap);
    k_deep2<32><<<blocks, MID2_THREADS, 0, s>>>(db, b, slab, lh, cap);
}

void launch_sort_sets(const BatchView &b, uint32_t cap, int sm_count, cudaStream_t s)
{
    const size_t smem = (size_t)std::max<uint32_t>(cap, 32) * 4;
    cudaFuncSetAttribute(k_sort_sets, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 64)));   // as many CTAs as the shared memory holds: barrier-bound
    k_sort_sets<<<sm_count * per_sm, SORT_THREADS, smem, s>>>(b, cap);
}

void launch_big(const DbView &db, const BatchView &b, const LargeState &ls, int blocks, bool from_huge, cudaStream_t s)
{
    k_big<<<blocks, BIG_THREADS, 0, s>>>(db, b, ls, from_huge ? 1 : 0);
}

// shapes of the CTA-per-probe tiers: the second tier runs several CTAs per SM on a 64 KB table, the third one CTA per SM
// on a 192 KB table (hash load <= 1/2 resp. 2/3 at the trail capacity, plus room for the entries every warp may add
// before the overflow is seen)
BlockTier block_tier_shape(int cls, uint32_t cap)
{
    BlockTier t;
    t.threads = cls ? 1024 : 256;
    t.cap = cap;
    t.hs = cls ? cap + cap / 2 : 2 * cap;
    t.hs = std::max<uint32_t>(t.hs, cap + 2048);
    return t;
}

template <int THREADS> static int block_bps(const BlockTier &t)
{
    int nb = 0;
    const size_t smem = (size_t)t.hs * 4;
    if (cudaFuncSetAttribute(k_block<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_block<THREADS>, THREADS, smem);
    return nb;
}

int block_tier_blocks_per_sm(const BlockTier &t) { return t.threads == 1024 ? block_bps<1024>(t) : block_bps<256>(t); }

void launch_block_tier(const DbView &db, const BatchView &b, const BlockTier &t, int cls, uint32_t *trail_slab, int blocks, cudaStream_t s)
{
    const size_t smem = (size_t)t.hs * 4;
    if (t.threads == 1024) k_block<1024><<<blocks, 1024, smem, s>>>(db, b, trail_slab, t.hs, t.cap, cls);
    else k_block<256><<<blocks, 256, smem, s>>>(db, b, trail_slab, t.hs, t.cap, cls);
}

int scan_gather_max_blocks(int sm_count)
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_scan_gather<false, false>, SG_THREADS, 0);
    return std::max(1, nb) * sm_count;
}

bool launch_scan_gather(const BatchView &b, int pass, int max_blocks, bool giant_sets_possible, const MergeArgs &mg, cudaStream_t s, bool pdl)
{
    // one contiguous segment of whole 2048-item tiles per CTA
    // one resident wave (max_blocks): measured 4 % faster than two waves of one-tile CTAs on one GPU, and with the merge
    // inside -- where every CTA ends with a wait + a remote-load round trip that a second wave would pay again -- 15 %
    const bool merge = mg.world > 1;
    const uint64_t limit = max_blocks;
    uint64_t blocks = std::min<uint64_t>(std::max<uint64_t>((b.n + SG_TILE - 1) / SG_TILE, 1), limit);
    uint64_t seg = ((b.n + blocks - 1) / blocks + SG_TILE - 1) / SG_TILE * SG_TILE;
    if (seg == 0) seg = SG_TILE;
    blocks = std::max<uint64_t>((b.n + seg - 1) / seg, 1);
    const dim3 g((unsigned)blocks), t(SG_THREADS);
    if (b.opts & OPT_OFF32) {
        if (merge) launch_pdl(k_scan_gather<true, true>, g, t, 0, s, pdl, b, pass, seg, mg);
        else launch_pdl(k_scan_gather<true, false>, g, t, 0, s, pdl, b, pass, seg, mg);
    } else {
        if (merge) launch_pdl(k_scan_gather<false, true>, g, t, 0, s, pdl, b, pass, seg, mg);
        else launch_pdl(k_scan_gather<false, false>, g, t, 0, s, pdl, b, pass, seg, mg);
    }
    if (giant_sets_possible) k_gather_long<<<max_blocks, 256, 0, s>>>(b);
    return cudaPeekAtLastError() == cudaSuccess;
}

void launch_check_fixpoints(const DbView &db, bool off32, const uint8_t *flag, const void *off, const int32_t *impl, uint64_t first, uint64_t count, uint64_t stride,
                            uint32_t *scratch, int blocks, unsigned long long *out, cudaStream_t s)
{
    if (off32) k_check_fixpoints<true><<<blocks, 256, 0, s>>>(db, flag, off, impl, first, count, stride, scratch, out);
    else k_check_fixpoints<false><<<blocks, 256, 0, s>>>(db, flag, off, impl, first, count, stride, scratch, out);
}

void launch_mark_level0(uint8_t *litinfo, const int32_t *lits, uint64_t n, cudaStream_t s)
{
    if (n) k_mark_level0<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 1024), 256, 0, s>>>(litinfo, lits, n);
}

void launch_merge_drain(const MergeArgs &m, cudaStream_t s) { k_merge_drain<<<1, 32, 0, s>>>(m); }

void launch_merge(const MergeArgs &m, int sm_count, cudaStream_t s)
{
    const uint64_t total = (m.wp / 4) * (uint64_t)std::max(m.world - 1, 0);
    const int blocks = (int)std::min<uint64_t>(std::max<uint64_t>((total + 255) / 256, 1), (uint64_t)sm_count * 2);
    k_merge<<<blocks, 256, 0, s>>>(m);
}

void launch_products(bool off32, const uint8_t *flag, const void *off, const int32_t *impl, uint64_t npairs, uint64_t var_base, uint32_t *necessary, int32_t *equiv,
                     unsigned long long equiv_cap, unsigned long long *counts, int sm_count, cudaStream_t s)
{
    const int blocks = (int)std::min<uint64_t>(std::max<uint64_t>((npairs + 7) / 8, 1), (uint64_t)sm_count * 8);
    if (off32) k_products<true><<<blocks, 256, 0, s>>>(flag, off, impl, npairs, var_base, necessary, equiv, equiv_cap, counts);
    else k_products<false><<<blocks, 256, 0, s>>>(flag, off, impl, npairs, var_base, necessary, equiv, equiv_cap, counts);
}

}  // namespace bbcp
