// bbcp_capi.cu -- the C ABI of include/bbcp.h: handle, device memory, launch orchestration, result fetch,
// and the Topor/IPASIR-style single-cube path. No propagation happens on the host.
#include "../../include/bbcp.h"
#include "bbcp_db.h"
#include "bbcp_device.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <unistd.h>
#include <chrono>
#include <string>
#include <vector>

using namespace bbcp;

namespace {

// Device buffers come from the device's stream-ordered memory pool (cudaMallocAsync), not from cudaMalloc: with peer
// GPUs attached (CUDA IPC mappings of the bitmap block enable peer access between the contexts) a cudaMalloc / cudaFree
// has to map / unmap the allocation in every peer context, which waits for the peers' running kernels -- and a peer
// may at that moment be spinning in a merge kernel for THIS rank's flag (a batch that needs its second-tier slabs
// for the first time, or a larger literal pool, allocates in the middle of the call). Pool memory is mapped on the
// owning device only, so growing a buffer never involves a peer. Only the bitmap block itself, which IS exported over
// IPC, is a cudaMalloc allocation (made in bbcp_finalize, before any peer can wait).
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaStream_t s = nullptr;   // the handle's stream (set by ensure_device); allocation, zeroing and release are ordered on it
    bool pooled = true;
    // grow-only; contents are not preserved. zero: the new allocation is cleared (on the stream the kernels run on)
    bool reserve(size_t need, bool zero = false)
    {
        if (need <= bytes && p) return true;
        release();
        size_t want = need + std::min<size_t>(need / 4, (size_t)1 << 30) + 256;   // head-room for the next batch, at most 1 GiB
        if (cudaMallocAsync(&p, want, s) != cudaSuccess) {
            cudaGetLastError();
            want = need;
            if (cudaMallocAsync(&p, want, s) != cudaSuccess) { cudaGetLastError(); p = nullptr; return false; }
        }
        bytes = want;
        if (zero) cudaMemsetAsync(p, 0, bytes, s);
        // the buffer is also used from the copy streams of bbcp_probe_lits_to_host: make it valid everywhere now
        // (allocations are rare: buffers only grow)
        if (cudaStreamSynchronize(s) != cudaSuccess) { cudaGetLastError(); return false; }
        return true;
    }
    void release()
    {
        if (p) { if (pooled) cudaFreeAsync(p, s); else cudaFree(p); }
        p = nullptr;
        bytes = 0;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

// misc device words (u64): [0..C_N) counters, cursor, a block of u32 words (U32_* below), total (a copy of off[n])
constexpr int MISC_CURSOR = C_N;
constexpr int MISC_U32 = C_N + 1;       // u32 words start at u32 index 2*MISC_U32
enum { U32_TICKET = 0 /* ..3: one dynamic scheduler per tier */, U32_MID_COUNT = 4, U32_BIG_COUNT = 5, U32_DEEP_COUNT = 6, U32_CHUNK_COUNT = 7,
       U32_TILE_TICKET = 8, U32_SG_FINISHED = 9 /* finished CTAs of k_scan_gather = "the report was written" */, U32_SORT_COUNT = 10,
       U32_HUGE_COUNT = 11, U32_N = 12 };
constexpr int MISC_TOTAL = MISC_U32 + U32_N / 2;
constexpr int MISC_WORDS64 = MISC_TOTAL + 1;
// two counter blocks used alternately: the triage kernel of one launch sequence zeroes the block of the next one

}  // namespace

struct bbcp_handle {
    HostDb db;
    int status = BBCP_OK;
    std::string err;
    int device = -1;
    bool prepared = false, finalized = false, device_ready = false;
    uint32_t db_version = 0;
    bool l0_dirty = false;   // level-0 facts were added on the device after the rows were built (bbcp_probe_to_fixpoint)
    int sm_count = 148;
    int sg_blocks = 148;   // co-resident CTAs of the cooperative compaction kernel

    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

    DevBuf d_hdr, d_slots, d_arena, d_litinfo, d_order;
    uint32_t order_n = 0;   // entries of the probe schedule on the device (0: none)
    // ONE allocation shared with peer GPUs (CUDA IPC): [failed: bm_stride words][unit: bm_stride words]
    // [256 control words: arrival epoch per rank, finished-CTA count, time-out latch, pull-done epoch per rank, empty-slice
    // hints], see merge_signal in bbcp_kernels.cu
    DevBuf d_bitmaps;
    uint64_t bm_stride = 0;
    uint32_t *failed_p() const { return d_bitmaps.as<uint32_t>(); }
    uint32_t *unit_p() const { return d_bitmaps.as<uint32_t>() + bm_stride; }
    int peer_rank = -1, peer_world = 0;
    std::vector<void *> peer_base;          // mapped base of every rank's d_bitmaps (own entry = own pointer)
    DevBuf d_peer_tab;                      // the same table on the device
    uint32_t merge_epoch = 0, drained_epoch = 0;
    bool bitmaps_dirty = false;             // some batch since the last clear found a failed literal
    uint32_t merge_timeout_ms = 20000;      // BBCP_MERGE_TIMEOUT_MS: how long a merging kernel waits for a peer (ranks may be seconds apart on skewed shards)
    cudaEvent_t ev_merge[2] = {nullptr, nullptr};
    uint64_t bitmap_words = 0;
    uint64_t device_db_bytes = 0;

    // results of a batch: two sets, so that the read-back of one chunk can overlap the next chunk's kernels
    // (bbcp_probe_lits_to_host); every other entry point uses set 0
    struct ResultSet { DevBuf flag, off, impl; } rs[2];
    int rb = 0;
    uint64_t off_base = 0;       // added to the offsets a batch writes (chunked calls)
    const int32_t *e2e_lits = nullptr;   // bbcp_probe_lits_to_host: the chunk's probe list / flag buffer in the caller's pinned memory
    uint8_t *e2e_flag = nullptr;         // (device addresses of mapped host memory; the triage kernel reads / writes them)
    uint32_t e2e_hint_flags = 0;         // options (+1) and stored literals of the last bbcp_probe_lits_to_host call: sizes the next call's chunks
    uint64_t e2e_hint_implied = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[16] = {}, ev_out[2] = {};
    DevBuf d_raw_off, d_len, d_out, d_cube_lits, d_cube_off, d_overflow, d_big, d_huge, d_deep, d_chunks, d_misc, d_lbits, d_ltrail, d_tiles;
    uint32_t misc_parity = 0;
    bool last_off32 = false;
    unsigned long long *h_misc = nullptr;   // pinned, mapped
    void *d_hmisc = nullptr;                // its device address
    int lean = 3;                           // bit 0: counters reach the host through mapped memory; bit 1: no per-kernel events
    uint64_t out_cap_hint = 0;   // stored literals the last batch needed (sizes the next batch's buffer)
    uint32_t big_slabs = 0;
    uint32_t slab_budget_per_sm = 8;   // resident k_big CTAs per SM the slabs are sized for (lowered when memory runs short)
    DevBuf d_mid_slab;           // second tier: per-warp hash + trail
    uint32_t mid_slab_cap = 0;   // trail capacity the slab is laid out (and zeroed) for
    uint64_t mid_slab_groups = 0;
    bool mid_slab_warp = false;  // the slab is laid out (zeroed hash tables) for the warp-per-probe second tier
    int32_t slab_max_var = -1;   // variable range the third-tier slabs are laid out for

    uint64_t last_n = 0, last_n_implied = 0, last_first = 0;
    bool last_universe = false, last_implied = false;
    DevBuf d_products;
    DevBuf d_flush;              // bbcp_l2_flush
    uint32_t flush_count = 0;
    bool have_batch = false;
    uint64_t total_props = 0;

    std::vector<int32_t> unit_log;   // failed-literal units in the order bbcp_probe_to_fixpoint derived them (each is RUP after the ones before)
    std::vector<int32_t> assumps;
    std::vector<int8_t> solve_val;   // per var after bbcp_solve: 0 unassigned, 1 true, -1 false (level > 0)
    bool solved = false;

    std::vector<DevBuf *> all_bufs()
    {
        return {&d_hdr, &d_slots, &d_arena, &d_litinfo, &d_order, &d_bitmaps, &d_peer_tab, &rs[0].flag, &rs[0].off, &rs[0].impl, &rs[1].flag, &rs[1].off,
                &rs[1].impl, &d_raw_off, &d_len, &d_out, &d_cube_lits, &d_cube_off, &d_overflow, &d_big, &d_huge, &d_deep, &d_chunks, &d_misc, &d_lbits, &d_ltrail,
                &d_tiles, &d_mid_slab, &d_products, &d_flush};
    }

    int fail(int code, const std::string &msg)
    {
        // resource / device failures are permanent (the reference's STATUS_FIRST_PERMANENT_ERROR class,
        // ToporExternalTypes.hpp:32-48); caller mistakes (ERR_STATE, ERR_ARG) only fail the call
        if (code == BBCP_ERR_ALLOC || code == BBCP_ERR_CUDA || code == BBCP_ERR_INDEX_TOO_NARROW || code == BBCP_CONTRADICTORY) status = code;
        err = msg;
        return code;
    }
    unsigned long long *h_dbg = nullptr;    // mapped: record of a failed check of a diagnostic build (make checks)
    bool cuda_ok(cudaError_t e, const char *what)
    {
        if (e == cudaSuccess) return true;
        fail(BBCP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
        if (h_dbg && h_dbg[0])
            err += " [check failed at bbcp_propagate.cuh/bbcp_kernels.cu line " + std::to_string(h_dbg[0]) + ", CTA " + std::to_string(h_dbg[1]) + ", thread " +
                   std::to_string(h_dbg[2]) + ", values " + std::to_string(h_dbg[3]) + " " + std::to_string(h_dbg[4]) + "]";
        cudaGetLastError();
        return false;
    }
};

namespace {

bool ensure_device(bbcp_handle *h)
{
    if (h->device_ready) return true;
    int ndev = 0;
    if (!h->cuda_ok(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount")) return false;
    if (ndev == 0) { h->fail(BBCP_ERR_CUDA, "no CUDA device: this engine has no CPU path"); return false; }
    if (h->device < 0) { if (!h->cuda_ok(cudaGetDevice(&h->device), "cudaGetDevice")) return false; }
    if (!h->cuda_ok(cudaSetDevice(h->device), "cudaSetDevice")) return false;
    cudaDeviceProp prop;
    if (!h->cuda_ok(cudaGetDeviceProperties(&prop, h->device), "cudaGetDeviceProperties")) return false;
    h->sm_count = prop.multiProcessorCount;
    h->sg_blocks = scan_gather_max_blocks(h->sm_count);
    if (!h->cuda_ok(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
    for (auto &e : h->ev) if (!h->cuda_ok(cudaEventCreate(&e), "cudaEventCreate")) return false;
    for (DevBuf *b : h->all_bufs()) b->s = h->stream;
    h->d_bitmaps.pooled = false;
    {
        // keep freed pool memory cached in the pool (the default gives it back to the driver at every synchronisation)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, h->device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if (!h->cuda_ok(cudaHostAlloc(&h->h_misc, MISC_WORDS64 * 8, cudaHostAllocMapped), "cudaHostAlloc")) return false;
    if (!h->cuda_ok(cudaHostGetDevicePointer(&h->d_hmisc, h->h_misc, 0), "cudaHostGetDevicePointer")) return false;
    if (cudaHostAlloc(&h->h_dbg, 65536, cudaHostAllocMapped) == cudaSuccess) {
        memset(h->h_dbg, 0, 65536);
        void *d = nullptr;
        if (cudaHostGetDevicePointer(&d, h->h_dbg, 0) == cudaSuccess) set_debug_buffer(static_cast<unsigned long long *>(d));
    }
    h->lean = getenv("BBCP_LEAN") ? atoi(getenv("BBCP_LEAN")) : 3;
    if (const char *e = getenv("BBCP_MERGE_TIMEOUT_MS")) h->merge_timeout_ms = (uint32_t)std::max(1, atoi(e));
    if (!h->d_misc.reserve(2 * MISC_WORDS64 * 8, true)) { h->fail(BBCP_ERR_ALLOC, "device allocation failed (counters)"); return false; }
    if (!h->cuda_ok(cudaDeviceSynchronize(), "counter init")) return false;
    h->device_ready = true;
    return true;
}

bool upload(bbcp_handle *h, DevBuf &b, const void *src, size_t bytes)
{
    if (!b.reserve(bytes ? bytes : 16)) { h->fail(BBCP_ERR_ALLOC, "device allocation failed (database)"); return false; }
    if (bytes && !h->cuda_ok(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream), "upload")) return false;
    return true;
}

bool drain_peers(bbcp_handle *h);
DbView db_view(const bbcp_handle *h);

bool upload_db(bbcp_handle *h, const std::vector<uint8_t> &varinfo, bool with_order)
{
    HostDb &db = h->db;
    // per-literal facts for the triage: level-0 value and mappedness of the variable, and what the row of
    // the literal's NEGATION holds (that row is the one scanned when the literal is assigned)
    std::vector<uint8_t> litinfo(db.hdr.size() + 16, 0);   // padded: the triage reads it in aligned 32-bit words
    for (size_t l = 2; l < db.hdr.size(); ++l) {
        const uint8_t vi = varinfo[l >> 1];
        uint8_t li = 0;
        if (vi & 4) li |= LI_MAPPED;
        if (vi & 3) li |= (((vi & 3) == 1) == ((l & 1) == 0)) ? LI_TRUE_L0 : LI_FALSE_L0;
        const RowHdr &r = db.hdr[l ^ 1];
        if (r.nbin) li |= LI_HAS_BIN;
        if (r.nbin | r.ntern | r.nlong) li |= LI_NONEMPTY;
        litinfo[l] = li;
    }
    if (!upload(h, h->d_hdr, db.hdr.data(), db.hdr.size() * sizeof(RowHdr))) return false;
    if (!upload(h, h->d_slots, db.slots.data(), db.slots.size() * sizeof(Slot))) return false;
    if (!upload(h, h->d_arena, db.arena.data(), db.arena.size() * 4)) return false;
    if (!upload(h, h->d_litinfo, litinfo.data(), litinfo.size())) return false;
    // probe schedule for closure reuse (only databases with binary clauses have one): built on the device from the rows
    // just uploaded (BBCP_ORDER=host: the exact host form, HostDb::build_order; BBCP_NO_ORDER: none, variable order)
    static const bool no_order = getenv("BBCP_NO_ORDER") != nullptr;
    static const bool host_order = getenv("BBCP_ORDER") && !strcmp(getenv("BBCP_ORDER"), "host");
    h->order_n = 0;
    if (with_order && !no_order && db.n_bin && db.hdr.size() > 2) {
        const uint32_t nl = (uint32_t)db.hdr.size();
        if (host_order) {
            db.build_order();
            if (!upload(h, h->d_order, db.order.data(), db.order.size() * 4)) return false;
        } else {
            if (!h->d_order.reserve((size_t)(nl - 2) * 4) || !h->d_products.reserve(2 * (size_t)((nl + 3) & ~3u) + 4 + 1024)) { h->fail(BBCP_ERR_ALLOC, "device allocation failed (probe schedule)"); return false; }
            DbView v = db_view(h);
            build_order_device(v, nl, h->d_products.as<uint8_t>(), h->d_order.as<uint32_t>(), h->sm_count, h->stream);
        }
        h->order_n = nl - 2;
    }
    h->bitmap_words = (2 * ((uint64_t)db.max_var + 1) + 31) / 32;
    const uint64_t stride = (h->bitmap_words + 3) & ~3ull;
    if (stride != h->bm_stride || !h->d_bitmaps.p) {
        if (h->peer_world) { h->fail(BBCP_ERR_STATE, "the bitmap size changed while peers are attached: bbcp_peer_detach first"); return false; }
        h->d_bitmaps.release();
        // exact size, own cudaMalloc: the IPC handle of this allocation is what peers map
        if (cudaMalloc(&h->d_bitmaps.p, (2 * stride + MERGE_CTRL_WORDS) * 4) != cudaSuccess) { cudaGetLastError(); h->d_bitmaps.p = nullptr; h->fail(BBCP_ERR_ALLOC, "device allocation failed (bitmaps)"); return false; }
        h->d_bitmaps.bytes = (2 * stride + MERGE_CTRL_WORDS) * 4;
        h->bm_stride = stride;
        cudaMemsetAsync(h->d_bitmaps.p, 0, h->d_bitmaps.bytes, h->stream);
    } else {
        drain_peers(h);   // a peer may still be reading the old bitmaps (last merge)
        cudaMemsetAsync(h->d_bitmaps.p, 0, 2 * stride * 4, h->stream);   // keep the arrival flags: they are monotonic epochs
    }
    // L2 persistence window (measured in round 2, OFF by default: BBCP_L2_WINDOW=hdr|slots|litinfo): keep one array of the
    // database resident in the set-aside part of the L2 for every kernel of this stream
    if (const char *w = getenv("BBCP_L2_WINDOW")) {
        DevBuf *buf = !strcmp(w, "hdr") ? &h->d_hdr : !strcmp(w, "slots") ? &h->d_slots : !strcmp(w, "litinfo") ? &h->d_litinfo : nullptr;
        int max_win = 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
        if (buf && buf->p && max_win > 0 && max_persist > 0) {
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            const size_t bytes = std::min<size_t>(buf->bytes, (size_t)max_win);
            attr.accessPolicyWindow.base_ptr = buf->p;
            attr.accessPolicyWindow.num_bytes = bytes;
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)max_persist / (double)bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            cudaGetLastError();
        }
    }
    if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "upload sync")) return false;
    h->l0_dirty = false;
    h->bitmaps_dirty = false;
    h->device_db_bytes = db.hdr.size() * sizeof(RowHdr) + db.slots.size() * sizeof(Slot) + db.arena.size() * 4 + litinfo.size() + 2 * h->bitmap_words * 4 + (size_t)h->order_n * 4;
    return true;
}

// the time-out latch of the merging kernels: read it and clear it, so that one late peer does not fail every later merge
bool merge_timed_out(bbcp_handle *h)
{
    uint32_t *latch = h->d_bitmaps.as<uint32_t>() + 2 * h->bm_stride + MERGE_CTRL_ERROR;
    uint32_t err = 0;
    cudaMemcpyAsync(&err, latch, 4, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    if (err) { cudaMemsetAsync(latch, 0, 4, h->stream); cudaStreamSynchronize(h->stream); }
    return err != 0;
}

// Before this rank's bitmaps are cleared, rebuilt or unmapped: wait until every peer has finished reading them in the
// last merge (the merge itself only waits for the peers' slices to ARRIVE, not for the peers to be done pulling ours).
bool drain_peers(bbcp_handle *h)
{
    if (h->peer_world <= 1 || h->merge_epoch == 0 || h->drained_epoch == h->merge_epoch) return true;
    launch_merge_drain(MergeArgs{static_cast<uint32_t *const *>(h->d_peer_tab.p), h->peer_rank, h->peer_world, h->bm_stride, 0, h->merge_epoch, h->merge_timeout_ms, 1u}, h->stream);
    if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "waiting for the peers to finish reading the bitmaps")) return false;
    h->drained_epoch = h->merge_epoch;
    if (merge_timed_out(h)) { h->fail(BBCP_ERR_STATE, "a peer did not finish the last bitmap merge within the time-out"); return false; }
    return true;
}

// BBCP_DEBUG_SYNC=1: synchronise after every kernel of a batch and name the one that faulted
bool dbg_check(bbcp_handle *h, const char *what)
{
    static const bool on = getenv("BBCP_DEBUG_SYNC") != nullptr;
    if (!on) return true;
    static const bool verbose = atoi(getenv("BBCP_DEBUG_SYNC")) >= 2;
    if (verbose) {
        fprintf(stderr, "bbcp: waiting for %s\n", what);
        fflush(stderr);
        // diagnostic builds: give up on a kernel that does not finish within 10 s and show where its first CTAs stand
        for (int i = 0; i < 1000 && cudaStreamQuery(h->stream) == cudaErrorNotReady; ++i) { struct timespec ts = {0, 10000000}; nanosleep(&ts, nullptr); }
        if (cudaStreamQuery(h->stream) == cudaErrorNotReady && h->h_dbg) {
            fprintf(stderr, "bbcp: %s still running after 10 s; per-warp marks (value:line) of the 256-thread CTAs that have not finished:\n", what);
            int shown = 0;
            for (int c = 0; c < 480 && shown < 12; ++c) {
                bool any = false, dead = true;
                for (int w = 0; w < 8; ++w) { const unsigned long long v = h->h_dbg[256 + c * 8 + w]; any |= v != 0; dead &= (v >> 32) == 0xDEAD; }
                if (!any || dead) continue;
                ++shown;
                fprintf(stderr, "  CTA %d:", c);
                for (int w = 0; w < 8; ++w) { const unsigned long long v = h->h_dbg[256 + c * 8 + w]; fprintf(stderr, " w%d=%llu:%llu", w, v >> 32, v & 0xffffffffull); }
                fprintf(stderr, "\n");
            }
            fprintf(stderr, "  last long hash probe (lit:steps) %llu:%llu\n", h->h_dbg[8] >> 32, h->h_dbg[8] & 0xffffffffull);
            fprintf(stderr, "  failed check: line %llu CTA %llu thread %llu values %llu %llu\n", h->h_dbg[0], h->h_dbg[1], h->h_dbg[2], h->h_dbg[3], h->h_dbg[4]);
            fflush(stderr);
            _exit(3);
        }
    }
    const cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) return true;
    h->fail(BBCP_ERR_CUDA, std::string("[debug] fault in ") + what + ": " + cudaGetErrorString(e));
    fprintf(stderr, "bbcp: %s\n", h->err.c_str());
    return false;
}

DbView db_view(const bbcp_handle *h)
{
    DbView v;
    v.hdr = h->d_hdr.as<uint4>();
    v.slots = h->d_slots.as<uint2>();
    v.arena = h->d_arena.as<uint4>();
    v.litinfo = h->d_litinfo.as<uint8_t>();
    v.max_var = h->db.max_var;
    v.n_slots = (uint32_t)h->db.slots.size();
    v.n_arena_units = (uint32_t)(h->db.arena.size() / 4);
    v.nvars_words16 = (uint32_t)((((uint64_t)h->db.max_var + 1 + 15) / 16 + 3) & ~3ull);   // whole 16-byte vectors: the last tier sweeps with 128-bit loads
    v.l0_dirty = h->l0_dirty ? 1u : 0u;
    return v;
}

uint32_t pow2_at_least(uint32_t x) { uint32_t p = 32; while (p < x) p <<= 1; return p; }

// third-tier slabs: one dense assignment + trail per resident CTA, sized for the whole variable range
bool ensure_slabs(bbcp_handle *h, const DbView &dbv)
{
    if (h->d_lbits.p && h->slab_max_var == dbv.max_var) return true;   // keyed on what sizes the slabs, not on the database version
    const size_t per_slab = (size_t)dbv.nvars_words16 * 4 + (size_t)big_trail_words((uint32_t)dbv.max_var) * 4;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    free_b += h->d_lbits.bytes + h->d_ltrail.bytes;
    h->d_lbits.release();
    h->d_ltrail.release();
    // one slab per resident CTA of k_big; more CTAs = more giant probes in flight (each is a serial chain of
    // breadth-first levels), bounded by a quarter of the free memory
    uint64_t per_sm = h->slab_budget_per_sm;
    if (const char *e = getenv("BBCP_BIG_SLABS_PER_SM")) per_sm = std::max(1, atoi(e));
    uint64_t slabs = std::min<uint64_t>((uint64_t)h->sm_count * per_sm, (free_b / 4) / std::max<size_t>(per_slab, 1));
    slabs = std::max<uint64_t>(slabs, 1);
    h->slab_max_var = -1;
    if (!h->d_lbits.reserve(slabs * dbv.nvars_words16 * 4, true) || !h->d_ltrail.reserve(slabs * (size_t)big_trail_words((uint32_t)dbv.max_var) * 4)) return false;
    h->big_slabs = (uint32_t)slabs;
    h->slab_max_var = dbv.max_var;
    return true;
}

// The whole device-side batch. Input arrays already on the device: (d_lits, d_off) cubes in CSR form,
// (d_lits, nullptr) a list of probe literals, (nullptr, nullptr) the probe universe from `first`.
// One stream, no memset nodes, one host synchronisation in the common case:
//     k_triage -> k_deep -> k_scan_gather -> counters through mapped host memory -> sync
// Only if some probe outgrew the warp-private shared-memory state (the read-back says so; k_scan_gather saw it
// too and returned at once) a second sequence runs: k_deep2 -> k_big -> k_sort_sets -> k_scan_gather
// (-> k_gather_long) -> read-back -> sync (BBCP_TIER2=block: k_block x 2 -> k_big instead). Each tier hands what it
// cannot hold to the next through a device-side list.
int run_batch(bbcp_handle *h, const int32_t *d_lits, const uint64_t *d_off, uint64_t first, uint64_t n,
              const bbcp_opts *uopts, bbcp_batch_info *info)
{
    bbcp_opts opts;
    memset(&opts, 0, sizeof(opts));
    opts.flags = BBCP_OPT_IMPLIED;
    if (uopts) opts = *uopts;
    if (n > 0x7ffffff0ull) return h->fail(BBCP_ERR_ARG, "batch too large (split it)");
    bbcp_batch_info bi;
    memset(&bi, 0, sizeof(bi));
    bi.n = n;
    h->have_batch = false;

    const size_t n1 = (size_t)n + 1;
    if (!h->rs[h->rb].flag.reserve(n1) || !h->d_raw_off.reserve(n1 * 8) || !h->d_len.reserve(n1 * 4 + 64) || !h->rs[h->rb].off.reserve(n1 * 8) ||
        !h->d_overflow.reserve(n1 * 4) || !h->d_big.reserve(n1 * 4) || !h->d_huge.reserve(n1 * 4) || !h->d_deep.reserve(n1 * 4) || !h->d_tiles.reserve((n / SG_TILE + 4) * 8))
        return h->fail(BBCP_ERR_ALLOC, "device allocation failed (batch buffers)");

    // lanes per probe (bbcp_propagate.cuh): a full warp. Narrower groups were measured and rejected, see DESIGN.md
    const int width = 32;
    uint32_t cap = opts.small_trail ? pow2_at_least(opts.small_trail) : 512u;
    if (!opts.small_trail) if (const char *e = getenv("BBCP_SMALL_TRAIL")) cap = pow2_at_least((uint32_t)atoi(e));
    if (cap > 8192) cap = 8192;
    while (small_smem_bytes(cap, width) > 200 * 1024 && cap > 32) cap >>= 1;
    uint32_t cap_mid = opts.mid_trail ? pow2_at_least(opts.mid_trail) : 8192;
    if (!opts.mid_trail) if (const char *e = getenv("BBCP_MID_TRAIL")) cap_mid = pow2_at_least((uint32_t)atoi(e));
    if (cap_mid > 16384) cap_mid = 16384;   // the second tier's set lives (and is sorted) in shared memory
    // third tier: one CTA per SM on a 192 KB table; with a test-sized second tier, four times its capacity
    uint32_t cap_big = opts.mid_trail ? std::min<uint32_t>(4 * cap_mid, 32768) : 32768;
    if (const char *e = getenv("BBCP_BIG_TRAIL")) cap_big = std::min<uint32_t>(pow2_at_least((uint32_t)atoi(e)), 32768);
    const bool implied = (opts.flags & BBCP_OPT_IMPLIED) != 0;
    const bool off32 = (opts.flags & BBCP_OPT_OFF32) != 0;
    const bool want_kernel_times = (opts.flags & BBCP_OPT_KERNEL_TIMES) != 0 || !(h->lean & 2);
    // BBCP_OPT_STORE_SETS: the implied sets are kept in the pool for closure reuse (a probe adopts the finished sets of the
    // literals it reaches) although the caller does not want them: flags + bitmaps come back, nothing is compacted
    const bool store = implied || (opts.flags & BBCP_OPT_STORE_SETS) != 0;
    uint64_t out_cap = store ? (opts.out_capacity ? opts.out_capacity : std::max<uint64_t>({(uint64_t)1 << 20, 4 * n, h->out_cap_hint + h->out_cap_hint / 4})) : 0;

    DbView dbv = db_view(h);
    LargeState ls{nullptr, nullptr};

    BatchView b;
    memset(&b, 0, sizeof(b));
    b.cube_lits = d_lits;
    b.cube_off = d_off;
    b.host_lits = d_off ? nullptr : h->e2e_lits;
    b.host_flag = d_off ? nullptr : h->e2e_flag;
    b.first = first;
    b.n = n;
    b.flag = h->rs[h->rb].flag.as<uint8_t>();
    b.raw_off = h->d_raw_off.as<uint64_t>();
    b.len = h->d_len.as<uint32_t>();
    b.failed = h->failed_p();
    b.unit = h->unit_p();
    b.mid_list = h->d_overflow.as<uint32_t>();
    b.big_list = h->d_big.as<uint32_t>();
    b.huge_list = h->d_huge.as<uint32_t>();
    b.deep_list = h->d_deep.as<uint32_t>();
    b.off = h->rs[h->rb].off.p;
    b.off_base = h->off_base;
    b.tile_sum = h->d_tiles.as<unsigned long long>();
    b.zero_words = MISC_WORDS64;
    b.opts = opts.flags | (store ? (uint32_t)BBCP_OPT_IMPLIED : 0u);
    // rows older than the level-0 facts: a ternary clause may have shrunk to a binary one, so the triage cannot tell from the
    // row shape which probes imply only themselves
    if (h->l0_dirty) b.opts |= BBCP_OPT_NO_LEN_PRUNE;
    // closure reuse: all-literal batches only (item i is the probe of internal literal first + i + 2)
    static const bool no_reuse = getenv("BBCP_NO_REUSE") != nullptr;
    if (!d_lits && !d_off && !no_reuse && !(opts.flags & BBCP_OPT_NO_REUSE)) {
        b.reuse = 1;
        if (h->order_n) { b.order = h->d_order.as<uint32_t>(); b.order_n = h->order_n; }
        // with closure reuse a probe beyond ~1000 literals is mostly adoption of finished sets, which the CTA-per-probe last
        // tier does with all its warps: the warp-per-probe second tier keeps only what is just above the first tier's capacity
        // (full-size C3, second-tier capacity 8192 / 4096 / 2048 / 1024 / 512: 1191 / 1034 / 1004 / 986 / 1001 ms per pass)
        if (!opts.mid_trail && !getenv("BBCP_MID_TRAIL")) cap_mid = 1024;
    }

    const int bps = small_blocks_per_sm(cap, width);
    if (bps <= 0) return h->fail(BBCP_ERR_CUDA, "propagate kernel does not fit (shared memory)");
    const uint32_t *h_u32 = reinterpret_cast<uint32_t *>(h->h_misc) + 2 * MISC_U32;
    float triage_ms = 0, deep_ms = 0, block_ms = 0, mid_ms = 0, compact_ms = 0, merge_ms = 0;

    // multi-GPU options (bbcp_probe_range only): device-side rendezvous with all peers before the batch, and the
    // all-gather of this rank's bitmap slice as the last kernel of the batch
    const bool peers = h->peer_world > 1 && !d_lits && !d_off;
    uint32_t *const *ptab = static_cast<uint32_t *const *>(h->d_peer_tab.p);
    if (peers && (opts.flags & BBCP_OPT_RENDEZVOUS))
        launch_merge(MergeArgs{ptab, h->peer_rank, h->peer_world, h->bm_stride, 0, ++h->merge_epoch, h->merge_timeout_ms, 1u}, h->sm_count, h->stream);
    const bool do_merge = peers && (opts.flags & BBCP_OPT_MERGE);
    if (do_merge && (opts.merge_words_per_rank == 0 || (opts.merge_words_per_rank & 3)))
        return h->fail(BBCP_ERR_ARG, "BBCP_OPT_MERGE: opts.merge_words_per_rank must be a non-zero multiple of 4");
    const uint32_t mepoch = do_merge ? ++h->merge_epoch : 0;
    // with implied sets the merge rides inside k_scan_gather (stores at its start, handshake at its end)
    const MergeArgs no_merge{nullptr, 0, 0, 0, 0, 0, 0, 0};
    const MergeArgs in_batch = do_merge ? MergeArgs{ptab, h->peer_rank, h->peer_world, h->bm_stride, opts.merge_words_per_rank, mepoch, h->merge_timeout_ms, h->bitmaps_dirty ? 1u : 0u} : no_merge;

    cudaEventRecord(h->ev[0], h->stream);
    for (int attempt = 0;; ++attempt) {
        // room for every self-implying probe (n) plus out_cap stored literals
        auto pool_ok = [&]() { return h->d_out.reserve(std::max<uint64_t>(out_cap, 4) * 4) && (!implied || h->rs[h->rb].impl.reserve((n + out_cap + 4) * 4)); };
        if (!pool_ok() && h->d_lbits.p) {
            // the third tier's slabs were sized when memory was plentiful: give most of them back and try again
            h->d_lbits.release();
            h->d_ltrail.release();
            h->slab_max_var = -1;
            h->slab_budget_per_sm = 1;
        }
        if (!pool_ok() ||
            !h->d_chunks.reserve((out_cap / 2048 + out_cap / 16384 + 64) * 8))   // sets above 16384 literals, 2048 per chunk, rounded up per set
            return h->fail(BBCP_ERR_ALLOC, "device allocation failed (implied literals)");
        b.out_lits = h->d_out.as<int32_t>();
        b.impl = h->rs[h->rb].impl.as<int32_t>();
        b.chunk_list = h->d_chunks.as<uint2>();
        b.out_cap = out_cap;
        // counter block of this launch sequence (zeroed by the previous sequence's triage, or at creation)
        unsigned long long *misc = h->d_misc.as<unsigned long long>() + (size_t)h->misc_parity * MISC_WORDS64;
        b.zero_next = h->d_misc.as<unsigned long long>() + (size_t)(h->misc_parity ^ 1u) * MISC_WORDS64;
        h->misc_parity ^= 1u;
        b.counters = misc;
        b.out_cursor = misc + MISC_CURSOR;
        b.total = misc + MISC_TOTAL;
        uint32_t *u32 = reinterpret_cast<uint32_t *>(misc) + 2 * MISC_U32;
        b.ticket = u32 + U32_TICKET;
        b.mid_count = u32 + U32_MID_COUNT;
        b.big_count = u32 + U32_BIG_COUNT;
        b.huge_count = u32 + U32_HUGE_COUNT;
        b.deep_count = u32 + U32_DEEP_COUNT;
        b.chunk_count = u32 + U32_CHUNK_COUNT;
        b.tile_ticket = u32 + U32_TILE_TICKET;
        b.sort_count = u32 + U32_SORT_COUNT;
        b.sort_list = h->d_deep.as<uint32_t>();   // the triage queue is spent by the time the second tier runs
        const bool zero_copy = implied && n && (h->lean & 1);
        const bool kevents = want_kernel_times;
        b.host_report = zero_copy ? static_cast<unsigned long long *>(h->d_hmisc) : nullptr;
        bi.launches = 0;
        bool second = false;
        reinterpret_cast<uint32_t *>(h->h_misc)[2 * MISC_U32 + U32_SG_FINISHED] = 0;   // "report written" mark of the mapped copy
        if (n) {
            if (kevents) cudaEventRecord(h->ev[3], h->stream);
            launch_triage(b, dbv, h->sm_count, h->stream);
            if (!dbg_check(h, "k_triage")) return BBCP_ERR_CUDA;
            if (kevents) cudaEventRecord(h->ev[4], h->stream);
            // (programmatic dependent launch only when no event sits between the kernels, and switchable for diagnosis)
            static const bool pdl_on = getenv("BBCP_NO_PDL") == nullptr;
            const bool pdl = pdl_on && !kevents;
            launch_small(dbv, b, cap, width, bps * h->sm_count, h->stream, pdl);
            if (!dbg_check(h, "k_deep")) return BBCP_ERR_CUDA;
            if (kevents) cudaEventRecord(h->ev[5], h->stream);
            bi.launches += 2;
            if (implied) {
                if (!launch_scan_gather(b, 0, h->sg_blocks, false, in_batch, h->stream, pdl)) return h->fail(BBCP_ERR_CUDA, "launch of the compaction kernel failed");
                if (!dbg_check(h, "k_scan_gather (first sequence)")) return BBCP_ERR_CUDA;
                ++bi.launches;
            }
        } else {
            // empty batch: nothing ran, but the next sequence still needs a zeroed counter block
            cudaMemsetAsync(b.zero_next, 0, MISC_WORDS64 * 8, h->stream);
            cudaMemsetAsync(h->rs[h->rb].off.p, 0, 8, h->stream);
        }
        if (!zero_copy) cudaMemcpyAsync(h->h_misc, misc, MISC_WORDS64 * 8, cudaMemcpyDeviceToHost, h->stream);
        cudaEventRecord(h->ev[2], h->stream);
        if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "propagate kernels")) return BBCP_ERR_CUDA;
        if (zero_copy && h_u32[U32_SG_FINISHED] == 0) {
            // k_scan_gather returned early (probes wait for the CTA-per-probe tiers): only then the report is missing
            cudaMemcpyAsync(h->h_misc, misc, MISC_WORDS64 * 8, cudaMemcpyDeviceToHost, h->stream);
            if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "counter read-back")) return BBCP_ERR_CUDA;
        }
        if (n && h_u32[U32_MID_COUNT] != 0) {
            // some probes outgrew the first tier: CTA-per-probe tiers, then the compaction for real
            second = true;
            // second tier: warp per probe, growing hash + trail in a global slab (k_deep2), then the dense last tier.
            // BBCP_TIER2=block selects the alternative that was built and measured in round 2 -- two CTA-per-probe tiers with
            // the hash in shared memory (k_block): no DRAM traffic for the state, but a CTA pays a barrier round per
            // breadth-first level and the Tseitin probes are hundreds of levels deep: 3.0 s against 1.95 s per C3 pass
            static const bool tier2_warp = !(getenv("BBCP_TIER2") && !strcmp(getenv("BBCP_TIER2"), "block"));
            if (!ensure_slabs(h, dbv)) return h->fail(BBCP_ERR_ALLOC, "device allocation failed (last-tier state)");
            ls.bits = h->d_lbits.as<uint32_t>();
            ls.trail = h->d_ltrail.as<uint32_t>();
            if (tier2_warp) {
                const int bps_mid = mid_blocks_per_sm(width);
                if (bps_mid <= 0) return h->fail(BBCP_ERR_CUDA, "second-tier kernel does not fit");
                const uint64_t gpc = MID2_THREADS / width;   // probes in flight per CTA
                const uint64_t mid_blocks = std::min<uint64_t>((uint64_t)bps_mid * h->sm_count, ((uint64_t)h_u32[U32_MID_COUNT] + gpc - 1) / gpc);
                const uint64_t mid_groups = (uint64_t)bps_mid * h->sm_count * gpc;
                if (h->mid_slab_cap != cap_mid || h->mid_slab_groups < mid_groups || !h->d_mid_slab.p || !h->mid_slab_warp) {
                    // (re)lay out the per-group slabs; the hash tables must start from zero
                    const size_t bytes = mid_groups * mid_slab_words_per_group(cap_mid) * 4;
                    if (!h->d_mid_slab.reserve(bytes)) return h->fail(BBCP_ERR_ALLOC, "device allocation failed (second-tier state)");
                    cudaMemsetAsync(h->d_mid_slab.p, 0, bytes, h->stream);
                    h->mid_slab_cap = cap_mid;
                    h->mid_slab_groups = mid_groups;
                    h->mid_slab_warp = true;
                }
                cudaEventRecord(h->ev[1], h->stream);
                launch_mid(dbv, b, h->d_mid_slab.as<uint32_t>(), cap_mid, width, (int)mid_blocks, h->stream);
                if (!dbg_check(h, "k_deep2")) return BBCP_ERR_CUDA;
                cudaEventRecord(h->ev[7], h->stream);
                // (the second tier's sets are sorted BEFORE the last tier runs: it adopts them, and its bulk adoption looks
                // literals up in a set by binary search)
                if (store) { launch_sort_sets(b, cap_mid, h->sm_count, h->stream); ++bi.launches; }
                if (!dbg_check(h, "k_sort_sets")) return BBCP_ERR_CUDA;
                launch_big(dbv, b, ls, (int)h->big_slabs, false, h->stream);
                if (!dbg_check(h, "k_big")) return BBCP_ERR_CUDA;
            } else {
                // second tier: 256-thread CTAs, several per SM; third tier: 1024-thread CTAs on (nearly) all of an SM's
                // shared memory; last tier: dense state, any size. Each hands what it cannot hold to the next through a list.
                const BlockTier t2 = block_tier_shape(0, cap_mid), t3 = block_tier_shape(1, cap_big);
                const int bps2 = block_tier_blocks_per_sm(t2), bps3 = block_tier_blocks_per_sm(t3);
                if (bps2 <= 0 || bps3 <= 0) return h->fail(BBCP_ERR_CUDA, "CTA-per-probe tier does not fit (shared memory)");
                const uint64_t blocks2 = (uint64_t)bps2 * h->sm_count, blocks3 = (uint64_t)bps3 * h->sm_count;
                if (!h->d_mid_slab.reserve((blocks2 * t2.cap + blocks3 * t3.cap) * 4)) return h->fail(BBCP_ERR_ALLOC, "device allocation failed (CTA-per-probe tiers)");
                h->mid_slab_warp = false;
                cudaEventRecord(h->ev[1], h->stream);
                launch_block_tier(dbv, b, t2, 0, h->d_mid_slab.as<uint32_t>(), (int)std::min<uint64_t>(blocks2, h_u32[U32_MID_COUNT]), h->stream);
                if (!dbg_check(h, "k_block (second tier)")) return BBCP_ERR_CUDA;
                cudaEventRecord(h->ev[7], h->stream);
                launch_block_tier(dbv, b, t3, 1, h->d_mid_slab.as<uint32_t>() + blocks2 * t2.cap, (int)blocks3, h->stream);
                if (!dbg_check(h, "k_block (third tier)")) return BBCP_ERR_CUDA;
                launch_big(dbv, b, ls, (int)h->big_slabs, true, h->stream);
                if (!dbg_check(h, "k_big")) return BBCP_ERR_CUDA;
                ++bi.launches;
            }
            cudaEventRecord(h->ev[6], h->stream);
            bi.launches += 2;
            if (implied) {
                if (!launch_scan_gather(b, 1, h->sg_blocks, true, in_batch, h->stream, false)) return h->fail(BBCP_ERR_CUDA, "launch of the compaction kernel failed");
                if (!dbg_check(h, "k_scan_gather / k_gather_long (second sequence)")) return BBCP_ERR_CUDA;
                bi.launches += 2;
            }
            cudaMemcpyAsync(h->h_misc, misc, MISC_WORDS64 * 8, cudaMemcpyDeviceToHost, h->stream);
            cudaEventRecord(h->ev[2], h->stream);
            if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "propagate kernels (CTA-per-probe tiers)")) return BBCP_ERR_CUDA;
        }
        if (n && kevents) {
            cudaEventElapsedTime(&triage_ms, h->ev[3], h->ev[4]);
            cudaEventElapsedTime(&deep_ms, h->ev[4], h->ev[5]);
            cudaEventElapsedTime(&compact_ms, second ? h->ev[6] : h->ev[5], h->ev[2]);
            if (second) {
                cudaEventElapsedTime(&block_ms, h->ev[1], h->ev[6]);
                cudaEventElapsedTime(&mid_ms, h->ev[1], h->ev[7]);
            }
        }
        const uint64_t cursor = h->h_misc[MISC_CURSOR];
        if (cursor <= out_cap) { bi.n_implied = implied ? h->h_misc[MISC_TOTAL] : 0; if (!opts.out_capacity) h->out_cap_hint = cursor; break; }
        // implied-literal pool too small: redo the batch with room for everything that was asked for
        out_cap = std::max<uint64_t>(cursor + cursor / 8, out_cap * 2);
        if (attempt == 7) return h->fail(BBCP_ERR_ALLOC, "implied-literal buffer keeps overflowing");
    }
    if (!implied || n == 0) {
        if (!implied) cudaMemsetAsync(h->rs[h->rb].off.p, 0, n1 * 8, h->stream);
        if (do_merge) {
            // flags-only or empty batch: no compaction kernel to ride in, the merge is its own launch
            cudaEventRecord(h->ev_merge[0], h->stream);
            launch_merge(in_batch, h->sm_count, h->stream);
            cudaEventRecord(h->ev[2], h->stream);
            ++bi.launches;
        }
        if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "batch")) return BBCP_ERR_CUDA;
        if (do_merge) cudaEventElapsedTime(&merge_ms, h->ev_merge[0], h->ev[2]);
    }
    if ((do_merge || (peers && (opts.flags & BBCP_OPT_RENDEZVOUS))) && merge_timed_out(h))
        return h->fail(BBCP_ERR_STATE, "bitmap merge: a peer did not arrive within the time-out (BBCP_MERGE_TIMEOUT_MS)");
    if (off32 && bi.n_implied > 0xffffffffull) return h->fail(BBCP_ERR_ARG, "BBCP_OPT_OFF32: more than 2^32-1 implied literals in one batch (split it)");
    h->last_off32 = off32 && implied;

    bi.props_all = h->h_misc[C_PROPS_ALL];
    bi.props_ref = h->h_misc[C_PROPS_REF];
    bi.slots = h->h_misc[C_SLOTS];
    bi.clause_fetches = h->h_misc[C_CLAUSE_FETCH];
    bi.n_large = h->h_misc[C_LARGE];
    bi.n_mid = h->h_misc[C_MID];
    bi.n_ok = h->h_misc[C_OK];
    bi.n_conflict = h->h_misc[C_CONFLICT];
    bi.n_skipped = h->h_misc[C_SKIPPED];
    bi.n_deep = h_u32[U32_DEEP_COUNT];
    bi.adopted = h->h_misc[C_ADOPTED];
    bi.adopt_lits = h->h_misc[C_ADOPT_LITS];
    bi.inherited = h->h_misc[C_INHERITED];
    bi.assignments = h->h_misc[C_ASSIGN];
    bi.assignments_ok = h->h_misc[C_ASSIGN_OK];
    bi.props_ref_ok = h->h_misc[C_PROPS_REF_OK];
    bi.bcps = h->h_misc[C_BCPS];
    bi.hot_rows = h->h_misc[C_HOT_ROWS];
    // bytes model: triage = 1 B of literal facts + 1 B flag + 4 B size per item (+ 4 B literal for
    // probe lists, + 16 B offsets more for cubes); propagation = row headers, row slots, clause units, stored
    // implied literals (kernel counters) + the per-item record (flag, size, position) of the queued items;
    // compaction = 4 B size read + offset written per item + 8 B per implied literal (read + write; 4 B for the
    // regenerated self literals)
    bi.triage_bytes = n * (6 + (d_off ? 20 : d_lits ? 4 : 0));
    bi.kernel_bytes = bi.triage_bytes + h->h_misc[C_BYTES] + 17ull * bi.n_deep;
    cudaEventElapsedTime(&bi.total_ms, h->ev[0], h->ev[2]);
    bi.triage_ms = triage_ms;
    bi.deep_ms = deep_ms;
    bi.block_ms = block_ms;
    bi.mid_ms = mid_ms;
    bi.compact_ms = implied ? compact_ms : 0.f;
    bi.merge_ms = merge_ms;
    bi.tier1_ms = triage_ms + deep_ms;
    bi.kernel_ms = triage_ms + deep_ms + block_ms;
    h->last_n = n;
    h->last_first = first;
    h->last_universe = !d_lits && !d_off;
    h->last_implied = implied;
    h->last_n_implied = bi.n_implied;
    h->have_batch = true;
    h->total_props += bi.props_ref;
    if (bi.n_conflict) h->bitmaps_dirty = true;
    if (info) *info = bi;
    return BBCP_OK;
}

int need_final(bbcp_handle *h)
{
    if (!h) return BBCP_ERR_ARG;
    if (h->status < 0) return h->status;
    if (!h->finalized) return h->fail(BBCP_ERR_STATE, "bbcp_finalize has not been called");
    if (h->status == BBCP_CONTRADICTORY) return BBCP_CONTRADICTORY;
    return BBCP_OK;
}

}  // namespace

extern "C" {

bbcp_handle *bbcp_create(int32_t nvars_hint)
{
    bbcp_handle *h = new (std::nothrow) bbcp_handle();
    if (h && nvars_hint > 0) h->db.raw.reserve((size_t)nvars_hint * 4);
    return h;
}

void bbcp_release(bbcp_handle *h)
{
    if (!h) return;
    if (h->device_ready) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        bbcp_peer_detach(h);
        for (auto &e : h->ev_merge) if (e) cudaEventDestroy(e);
        for (DevBuf *b : h->all_bufs()) b->release();
        cudaStreamSynchronize(h->stream);
        if (h->h_misc) cudaFreeHost(h->h_misc);
        if (h->h_dbg) cudaFreeHost(h->h_dbg);
        for (auto &e : h->ev_in) if (e) cudaEventDestroy(e);
        for (auto &e : h->ev_out) if (e) cudaEventDestroy(e);
        if (h->s_in) cudaStreamDestroy(h->s_in);
        if (h->s_out) cudaStreamDestroy(h->s_out);
        for (auto &e : h->ev) if (e) cudaEventDestroy(e);
        if (h->stream) cudaStreamDestroy(h->stream);
    }
    delete h;
}

int bbcp_set_device(bbcp_handle *h, int device)
{
    if (!h) return BBCP_ERR_ARG;
    if (h->device_ready) return h->fail(BBCP_ERR_STATE, "device already in use by this handle");
    h->device = device;
    return BBCP_OK;
}

static int intake_guard(bbcp_handle *h)
{
    if (!h) return BBCP_ERR_ARG;
    if (h->status < 0) return h->status;
    // clauses added after a finalize make the device snapshot stale: the next finalize rebuilds it
    h->finalized = false;
    h->prepared = false;
    return BBCP_OK;
}

// internal literals are 2*var+neg in 30 bits (the two bits above are the trail-entry mode, bbcp_device.cuh): variables stop
// below 2^29. The reference reports STATUS_INDEX_TOO_NARROW for what its index type cannot hold (TopiConflictAnalysis.cc:67-71)
static inline bool lit_in_range(int32_t l) { return l > -(1 << 29) && l < (1 << 29); }
static const char *const LIT_RANGE_MSG = "literal out of range (|literal| must be below 2^29)";

int bbcp_add(bbcp_handle *h, int32_t lit)
{
    int rc = intake_guard(h);
    if (rc) return rc;
    if (!lit_in_range(lit)) return h->fail(BBCP_ERR_INDEX_TOO_NARROW, LIT_RANGE_MSG);
    try {
    h->db.pending.push_back(lit);
    if (lit == 0) {
        h->db.add_words(h->db.pending.data(), h->db.pending.size());
        h->db.pending.clear();
    }
    } catch (const std::bad_alloc &) { return h->fail(BBCP_ERR_ALLOC, "host allocation failed (clause intake)"); }
    return BBCP_OK;
}

int bbcp_add_clause(bbcp_handle *h, const int32_t *lits, size_t n)
{
    int rc = intake_guard(h);
    if (rc) return rc;
    size_t m = 0;
    while (m < n && lits[m] != 0) {
        if (!lit_in_range(lits[m])) return h->fail(BBCP_ERR_INDEX_TOO_NARROW, LIT_RANGE_MSG);
        ++m;
    }
    try {
        h->db.add_words(lits, m);
        const int32_t z = 0;
        h->db.add_words(&z, 1);
    } catch (const std::bad_alloc &) { return h->fail(BBCP_ERR_ALLOC, "host allocation failed (clause intake)"); }
    return BBCP_OK;
}

int bbcp_add_clauses(bbcp_handle *h, const int32_t *words, uint64_t nwords)
{
    int rc = intake_guard(h);
    if (rc) return rc;
    if (nwords && words[nwords - 1] != 0) return h->fail(BBCP_ERR_ARG, "clause stream must end with 0");
    for (uint64_t i = 0; i < nwords; ++i) if (!lit_in_range(words[i])) return h->fail(BBCP_ERR_INDEX_TOO_NARROW, LIT_RANGE_MSG);
    try { h->db.add_words(words, nwords); }
    catch (const std::bad_alloc &) { return h->fail(BBCP_ERR_ALLOC, "host allocation failed (clause intake)"); }
    return BBCP_OK;
}

static int host_prepare_impl(bbcp_handle *h);
static int finalize_impl(bbcp_handle *h);

// no exception crosses the C ABI: a host allocation failure becomes BBCP_ERR_ALLOC (the reference's STATUS_ALLOC_FAILED)
int bbcp_host_prepare(bbcp_handle *h)
{
    try { return host_prepare_impl(h); }
    catch (const std::bad_alloc &) { return h->fail(BBCP_ERR_ALLOC, "host allocation failed (database build)"); }
}

int bbcp_finalize(bbcp_handle *h)
{
    try { return finalize_impl(h); }
    catch (const std::bad_alloc &) { return h->fail(BBCP_ERR_ALLOC, "host allocation failed (database build)"); }
}

static int host_prepare_impl(bbcp_handle *h)
{
    if (!h) return BBCP_ERR_ARG;
    if (h->status < 0) return h->status;
    // BBCP_HOST_TRACE: seconds of the two host phases on stderr
    const bool trace = getenv("BBCP_HOST_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    h->db.normalise();
    const auto t1 = std::chrono::steady_clock::now();
    if (h->db.contradictory) { h->prepared = true; return h->fail(BBCP_CONTRADICTORY, "AddClause: the clause set is contradictory at intake"); }
    const std::string e = h->db.build_rows();
    if (trace)
        fprintf(stderr, "bbcp host prepare: normalise %.3f s (%s), build_rows %.3f s (%zu words in, %zu slots)\n", std::chrono::duration<double>(t1 - t0).count(),
                h->db.intake_by_chunk ? "by chunk" : "stream order", std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count(), h->db.raw.size(),
                h->db.slots.size());
    if (!e.empty()) return h->fail(BBCP_ERR_INDEX_TOO_NARROW, e);
    h->prepared = true;
    if (h->status == BBCP_CONTRADICTORY) h->status = BBCP_OK;
    return BBCP_OK;
}

static int finalize_impl(bbcp_handle *h)
{
    if (!h) return BBCP_ERR_ARG;
    if (h->status < 0) return h->status;
    h->status = BBCP_OK;
    h->finalized = false;
    h->have_batch = false;
    h->solved = false;
    if (!ensure_device(h)) return h->status;
    cudaSetDevice(h->device);
    int rc = host_prepare_impl(h);
    if (rc == BBCP_CONTRADICTORY) { h->finalized = true; ++h->db_version; return rc; }
    if (rc) return rc;
    HostDb &db = h->db;
    if (!db.units.empty()) {
        // level-0 unit propagation on the device: the unit clauses form one cube over the un-simplified
        // database (Topi.cc:1137-1160 runs BCP() at level 0 before anything else)
        std::vector<uint8_t> mapped_only(db.varinfo);
        for (auto &v : mapped_only) v &= 4;
        if (!upload_db(h, mapped_only, false)) return h->status;
        std::vector<int32_t> cube(db.units.size());
        for (size_t i = 0; i < cube.size(); ++i) cube[i] = (db.units[i] & 1) ? -(int32_t)(db.units[i] >> 1) : (int32_t)(db.units[i] >> 1);
        const uint64_t off[2] = {0, cube.size()};
        h->finalized = true;  // the snapshot is usable for this internal batch
        bbcp_batch_info bi;
        rc = bbcp_propagate_batch(h, cube.data(), off, 1, nullptr, &bi);
        h->finalized = false;
        if (rc) return rc;
        uint8_t flag = 0;
        if (bbcp_fetch_flags(h, &flag)) return h->status;
        if (flag == BBCP_FLAG_CONFLICT) { h->finalized = true; ++h->db_version; return h->fail(BBCP_CONTRADICTORY, "level-0 unit propagation derives a conflict"); }
        std::vector<uint64_t> ioff(2);
        std::vector<int32_t> lits(std::max<uint64_t>(bi.n_implied, 1));
        if (bbcp_fetch_implied(h, ioff.data(), lits.data())) return h->status;
        for (uint64_t i = 0; i < bi.n_implied; ++i) {
            const int32_t l = lits[i];
            const uint32_t v = (uint32_t)(l < 0 ? -l : l);
            db.varinfo[v] = (db.varinfo[v] & ~3) | (l < 0 ? 2 : 1);
        }
        if (!db.simplify_level0()) return h->fail(BBCP_ERR_STATE, "internal: level-0 fixpoint left a unit clause");
        const std::string e = db.build_rows();
        if (!e.empty()) return h->fail(BBCP_ERR_INDEX_TOO_NARROW, e);
        h->total_props = 0;
    }
    if (!upload_db(h, db.varinfo, true)) return h->status;
    h->finalized = true;
    h->have_batch = false;
    ++h->db_version;
    return BBCP_OK;
}

int bbcp_get_db_info(const bbcp_handle *h, bbcp_db_info *out)
{
    if (!h || !out) return BBCP_ERR_ARG;
    memset(out, 0, sizeof(*out));
    out->max_var = h->db.max_var;
    out->n_mapped_vars = h->db.n_mapped();
    out->n_level0 = h->db.n_level0();
    out->n_bin = h->db.n_bin;
    out->n_tern = h->db.n_tern;
    out->n_long = h->db.n_long;
    out->n_long_lits = h->db.n_long_lits;
    out->row_slots = h->db.slots.size();
    out->arena_words = h->db.arena.size();
    out->device_bytes = h->device_db_bytes;
    out->db_version = h->db_version;
    return BBCP_OK;
}

uint64_t bbcp_level0_lits(const bbcp_handle *h, int32_t *out, uint64_t cap)
{
    if (!h) return 0;
    uint64_t n = 0;
    for (size_t v = 1; v < h->db.varinfo.size(); ++v) {
        const uint8_t val = h->db.varinfo[v] & 3;
        if (!val) continue;
        if (out && n < cap) out[n] = val == 1 ? (int32_t)v : -(int32_t)v;
        ++n;
    }
    return n;
}

uint64_t bbcp_unit_log(const bbcp_handle *h, int32_t *out, uint64_t cap)
{
    if (!h) return 0;
    for (size_t i = 0; i < h->unit_log.size() && i < cap; ++i) out[i] = h->unit_log[i];
    return h->unit_log.size();
}

uint64_t bbcp_host_row(const bbcp_handle *h, uint32_t ilit, uint32_t *out, uint64_t cap)
{
    if (!h || !h->prepared || ilit >= h->db.hdr.size()) return 0;
    const RowHdr &r = h->db.hdr[ilit];
    std::vector<uint32_t> w = {r.nbin, r.ntern, r.nlong};
    const Slot *s = h->db.slots.data() + r.start;
    for (uint32_t i = 0; i < r.nbin; ++i) w.push_back((i & 1) ? s[i / 2].y : s[i / 2].x);
    s += (r.nbin + 1) / 2;
    for (uint32_t i = 0; i < r.ntern; ++i) { w.push_back(s[i].x); w.push_back(s[i].y); }
    s += r.ntern;
    for (uint32_t i = 0; i < r.nlong; ++i) {
        w.push_back(s[i].x);
        const uint32_t *c = h->db.arena.data() + (size_t)s[i].y * 4;
        for (uint32_t k = 0; k <= c[0]; ++k) w.push_back(c[k]);
    }
    for (size_t i = 0; i < w.size() && i < cap; ++i) out[i] = w[i];
    return w.size();
}

uint64_t bbcp_host_order(bbcp_handle *h, uint32_t *out, uint64_t cap)
{
    if (!h || !h->prepared) return 0;
    try { h->db.build_order(); } catch (const std::bad_alloc &) { h->fail(BBCP_ERR_ALLOC, "host allocation failed (probe schedule)"); return 0; }
    for (size_t i = 0; i < h->db.order.size() && i < cap; ++i) out[i] = h->db.order[i];
    return h->db.order.size();
}

int bbcp_status(const bbcp_handle *h) { return h ? h->status : BBCP_ERR_ARG; }
const char *bbcp_error_string(const bbcp_handle *h) { return h ? h->err.c_str() : "null handle"; }

int bbcp_propagate_batch_device(bbcp_handle *h, const int32_t *d_lits, const uint64_t *d_off, uint64_t ncubes, uint64_t nlits,
                                const bbcp_opts *opts, bbcp_batch_info *info)
{
    int rc = need_final(h);
    if (rc) return rc;
    (void)nlits;
    cudaSetDevice(h->device);
    return run_batch(h, d_lits, d_off, 0, ncubes, opts, info);
}

int bbcp_propagate_batch(bbcp_handle *h, const int32_t *lits, const uint64_t *off, uint64_t ncubes, const bbcp_opts *opts,
                         bbcp_batch_info *info)
{
    int rc = need_final(h);
    if (rc) return rc;
    if (!off) return h->fail(BBCP_ERR_ARG, "cube offsets missing");
    cudaSetDevice(h->device);
    const uint64_t nlits = off[ncubes];
    if (!h->d_cube_lits.reserve(std::max<uint64_t>(nlits, 4) * 4) || !h->d_cube_off.reserve((ncubes + 1) * 8))
        return h->fail(BBCP_ERR_ALLOC, "device allocation failed (cubes)");
    if (nlits) cudaMemcpyAsync(h->d_cube_lits.p, lits, nlits * 4, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_cube_off.p, off, (ncubes + 1) * 8, cudaMemcpyHostToDevice, h->stream);
    return run_batch(h, h->d_cube_lits.as<int32_t>(), h->d_cube_off.as<uint64_t>(), 0, ncubes, opts, info);
}

int bbcp_probe_lits(bbcp_handle *h, const int32_t *lits, uint64_t n, const bbcp_opts *opts, bbcp_batch_info *info)
{
    int rc = need_final(h);
    if (rc) return rc;
    if (!lits && n) return h->fail(BBCP_ERR_ARG, "probe list missing");
    cudaSetDevice(h->device);
    if (!h->d_cube_lits.reserve(std::max<uint64_t>(n, 4) * 4)) return h->fail(BBCP_ERR_ALLOC, "device allocation failed (probe list)");
    if (n) cudaMemcpyAsync(h->d_cube_lits.p, lits, n * 4, cudaMemcpyHostToDevice, h->stream);
    return run_batch(h, h->d_cube_lits.as<int32_t>(), nullptr, 0, n, opts, info);
}

int bbcp_probe_lits_to_host(bbcp_handle *h, const int32_t *lits, uint64_t n, const bbcp_opts *uopts, uint8_t *flag_out, void *off_out,
                            int32_t *impl_out, uint64_t impl_cap, bbcp_batch_info *info)
{
    int rc = need_final(h);
    if (rc) return rc;
    if ((!lits && n) || !flag_out || !off_out || (!impl_out && impl_cap)) return h->fail(BBCP_ERR_ARG, "bbcp_probe_lits_to_host: null buffer");
    cudaSetDevice(h->device);
    bbcp_opts opts;
    memset(&opts, 0, sizeof(opts));
    opts.flags = BBCP_OPT_IMPLIED;
    if (uopts) opts = *uopts;
    opts.flags |= BBCP_OPT_IMPLIED;
    opts.flags &= ~(uint32_t)(BBCP_OPT_MERGE | BBCP_OPT_RENDEZVOUS);
    const bool off32 = (opts.flags & BBCP_OPT_OFF32) != 0;
    const size_t osz = off32 ? 4 : 8;
    // How the probe list reaches the device and the flags reach the host (BBCP_E2E=pull|mirror|dma, default pull):
    //   pull    caller buffers in pinned memory are used in place: the triage kernel reads the list over PCIe with 16-byte loads
    //           (and leaves a device copy for the kernels after it) and writes the flags it decides into the caller's flag
    //           buffer, so input, compute and flag read-back of a batch overlap inside one kernel
    //   mirror  the list is uploaded by the copy engine (chunks on their own stream), the flags are written as in pull
    //   dma     uploads and read-backs are all copy-engine transfers (the round-1 path; also what pageable buffers get)
    enum { E2E_DMA, E2E_MIRROR, E2E_PULL };
    int mode = E2E_PULL;
    if (const char *e = getenv("BBCP_E2E")) mode = !strcmp(e, "dma") ? E2E_DMA : !strcmp(e, "mirror") ? E2E_MIRROR : E2E_PULL;
    auto mapped = [](const void *p) -> void * {
        cudaPointerAttributes a;
        if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
    };
    const int32_t *m_lits = mode == E2E_PULL && n ? static_cast<const int32_t *>(mapped(lits)) : nullptr;
    uint8_t *m_flag = mode != E2E_DMA && n && (reinterpret_cast<uintptr_t>(flag_out) & 7u) == 0 ? static_cast<uint8_t *>(mapped(flag_out)) : nullptr;
    const bool pull = m_lits != nullptr;
    // chunks: whole compaction tiles, at most 16 of them, not smaller than 1 M probes (every chunk pays the fixed ~35 us of
    // a batch, about twice that while an upload is in flight). Pulling, a call that brings back little more than flags
    // (as its predecessor with the same options did) is one batch: there is no transfer to hide behind another chunk.
    uint64_t nchunks = std::min<uint64_t>(16, std::max<uint64_t>(1, n / 1000000));
    if (pull && h->e2e_hint_flags == opts.flags + 1u && h->e2e_hint_implied * 4 < n) nchunks = 1;
    if (const char *e = getenv("BBCP_CHUNKS")) nchunks = std::min<uint64_t>(16, std::max<uint64_t>(1, (uint64_t)atoi(e)));
    uint64_t csize = ((n + nchunks - 1) / nchunks + SG_TILE - 1) / SG_TILE * SG_TILE;
    if (csize == 0) csize = SG_TILE;
    nchunks = std::max<uint64_t>(1, (n + csize - 1) / csize);
    if (!h->s_in) {
        if (!h->cuda_ok(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking), "cudaStreamCreate") ||
            !h->cuda_ok(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking), "cudaStreamCreate")) return BBCP_ERR_CUDA;
        for (auto &e : h->ev_in) if (!h->cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate")) return BBCP_ERR_CUDA;
        for (auto &e : h->ev_out) if (!h->cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate")) return BBCP_ERR_CUDA;
    }
    if (!h->d_cube_lits.reserve(std::max<uint64_t>(n, 4) * 4)) return h->fail(BBCP_ERR_ALLOC, "device allocation failed (probe list)");
    // BBCP_E2E_TRACE: host time line of the call on stderr (us since entry: uploads queued, every chunk's batch done, read-back done)
    const bool trace = getenv("BBCP_E2E_TRACE") != nullptr;
    const auto t_enter = std::chrono::steady_clock::now();
    auto us_now = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_enter).count(); };
    double t_up = 0, t_chunk[16] = {};
    // all input chunks go out at once on their own stream; each batch waits for its chunk only
    for (uint64_t k = 0; k < nchunks && !pull; ++k) {
        const uint64_t lo = k * csize, cnt = std::min(csize, n - lo);
        cudaMemcpyAsync(h->d_cube_lits.as<int32_t>() + lo, lits + lo, cnt * 4, cudaMemcpyHostToDevice, h->s_in);
        cudaEventRecord(h->ev_in[k], h->s_in);
    }
    if (trace) t_up = us_now();
    bbcp_batch_info tot;
    memset(&tot, 0, sizeof(tot));
    uint64_t base = 0;
    int ret = BBCP_OK;
    bool any_stored = false;
    for (uint64_t k = 0; k < nchunks && ret == BBCP_OK; ++k) {
        const uint64_t lo = k * csize, cnt = std::min(csize, n - lo);
        h->rb = (int)(k & 1);
        h->off_base = base;
        if (k >= 2) cudaStreamWaitEvent(h->stream, h->ev_out[k & 1], 0);   // this result set is still being read back
        if (!pull) cudaStreamWaitEvent(h->stream, h->ev_in[k], 0);
        bbcp_batch_info bi;
        h->e2e_lits = pull ? m_lits + lo : nullptr;
        h->e2e_flag = m_flag ? m_flag + lo : nullptr;
        ret = run_batch(h, h->d_cube_lits.as<int32_t>() + lo, nullptr, 0, cnt, &opts, &bi);
        h->e2e_lits = nullptr;
        h->e2e_flag = nullptr;
        if (ret != BBCP_OK) break;
        if (base + bi.n_implied > impl_cap) { ret = h->fail(BBCP_ERR_ARG, "bbcp_probe_lits_to_host: implied-literal buffer too small"); break; }
        // the batch is complete (run_batch synchronises): its results leave on the output stream while the next
        // chunk computes into the other result set
        // (flags the triage wrote into the caller's buffer are final when it queued nothing for the propagation tiers)
        if (!(m_flag && bi.n_deep == 0)) cudaMemcpyAsync(flag_out + lo, h->rs[h->rb].flag.p, cnt, cudaMemcpyDeviceToHost, h->s_out);
        // compact read-back: while no chunk has stored a set, every offset is zero and nothing but the flags travels
        const bool compact = (opts.flags & BBCP_OPT_COMPACT_SELF) != 0;
        if (compact && bi.n_implied && !any_stored && lo) memset(off_out, 0, lo * osz);   // the chunks skipped so far
        any_stored |= bi.n_implied != 0;
        if (!compact || any_stored) cudaMemcpyAsync(static_cast<char *>(off_out) + lo * osz, h->rs[h->rb].off.p, cnt * osz, cudaMemcpyDeviceToHost, h->s_out);
        if (bi.n_implied) cudaMemcpyAsync(impl_out + base, h->rs[h->rb].impl.p, bi.n_implied * 4, cudaMemcpyDeviceToHost, h->s_out);
        cudaEventRecord(h->ev_out[k & 1], h->s_out);
        if (trace) t_chunk[k] = us_now();
        base += bi.n_implied;
        tot.n += bi.n; tot.n_implied += bi.n_implied; tot.n_ok += bi.n_ok; tot.n_conflict += bi.n_conflict; tot.n_skipped += bi.n_skipped;
        tot.n_large += bi.n_large; tot.n_mid += bi.n_mid; tot.props_all += bi.props_all; tot.props_ref += bi.props_ref; tot.slots += bi.slots;
        tot.clause_fetches += bi.clause_fetches; tot.kernel_bytes += bi.kernel_bytes; tot.triage_bytes += bi.triage_bytes; tot.n_deep += bi.n_deep;
        tot.adopted += bi.adopted; tot.adopt_lits += bi.adopt_lits; tot.inherited += bi.inherited; tot.assignments += bi.assignments; tot.bcps += bi.bcps;
        tot.assignments_ok += bi.assignments_ok; tot.props_ref_ok += bi.props_ref_ok; tot.hot_rows += bi.hot_rows;
        tot.kernel_ms += bi.kernel_ms; tot.total_ms += bi.total_ms; tot.launches += bi.launches; tot.triage_ms += bi.triage_ms;
        tot.deep_ms += bi.deep_ms; tot.block_ms += bi.block_ms; tot.mid_ms += bi.mid_ms; tot.compact_ms += bi.compact_ms; tot.tier1_ms += bi.tier1_ms;
    }
    h->rb = 0;
    h->off_base = 0;
    h->have_batch = false;   // the results are in the caller's buffers, not in one device-side batch
    h->e2e_hint_flags = opts.flags + 1u;
    h->e2e_hint_implied = base;
    const bool ok = h->cuda_ok(cudaStreamSynchronize(h->s_out), "result read-back") && h->cuda_ok(cudaStreamSynchronize(h->s_in), "probe upload");
    if (trace) {
        fprintf(stderr, "bbcp_probe_lits_to_host (%s%s): n %llu, %llu chunks of %llu; uploads queued %.1f us; batches done", pull ? "pull" : "upload",
                m_flag ? ", flags in place" : "", (unsigned long long)n, (unsigned long long)nchunks, (unsigned long long)csize, t_up);
        for (uint64_t k = 0; k < nchunks; ++k) fprintf(stderr, " %.1f", t_chunk[k]);
        fprintf(stderr, "; read-back done %.1f us; kernels %.1f us\n", us_now(), 1e3 * tot.total_ms);
    }
    if (ret != BBCP_OK) return ret;
    if (!ok) return BBCP_ERR_CUDA;
    if (off32) {
        if (base > 0xffffffffull) return h->fail(BBCP_ERR_ARG, "BBCP_OPT_OFF32: more than 2^32-1 implied literals");
        static_cast<uint32_t *>(off_out)[n] = (uint32_t)base;
    } else static_cast<uint64_t *>(off_out)[n] = base;
    (void)any_stored;
    if (info) *info = tot;
    return BBCP_OK;
}

int bbcp_probe_range(bbcp_handle *h, uint64_t first, uint64_t last, const bbcp_opts *opts, bbcp_batch_info *info)
{
    int rc = need_final(h);
    if (rc) return rc;
    const uint64_t universe = 2ull * (uint64_t)h->db.max_var;
    if (first > last || last > universe) return h->fail(BBCP_ERR_ARG, "probe range outside the universe");
    cudaSetDevice(h->device);
    return run_batch(h, nullptr, nullptr, first, last - first, opts, info);
}

int bbcp_probe_all(bbcp_handle *h, const bbcp_opts *opts, bbcp_batch_info *info)
{
    if (!h) return BBCP_ERR_ARG;
    return bbcp_probe_range(h, 0, 2ull * (uint64_t)h->db.max_var, opts, info);
}

// Failed-literal probing to a fixpoint: probe every literal, add the negation of every failed literal as a unit
// clause, re-finalize (level-0 propagation + simplification of the database), repeat until a pass finds no new
// failed literal. What the units would do inside the reference: each is an Assign at level 0 followed by BCP
// (Topi.cc:1311-1333), and the satisfied clauses / false literals go away in SimplifyIfRequired
// (TopiCompression.cc:509-640).
int bbcp_probe_to_fixpoint(bbcp_handle *h, uint32_t max_rounds, uint32_t *rounds_out, uint64_t *units_out)
{
    if (rounds_out) *rounds_out = 0;
    if (units_out) *units_out = 0;
    if (!h) return BBCP_ERR_ARG;
    if (h->status == BBCP_OK && !h->finalized) bbcp_finalize(h);
    int rc = need_final(h);
    if (rc) return rc;
    // BBCP_FIXPOINT=rebuild: the round-1 loop (re-finalize = host rebuild of the database after every round)
    static const bool rebuild = getenv("BBCP_FIXPOINT") && !strcmp(getenv("BBCP_FIXPOINT"), "rebuild");
    uint64_t total_units = 0;
    uint32_t round = 0;
    std::vector<uint32_t> unit;
    std::vector<int32_t> cube, lits;
    for (; max_rounds == 0 || round < max_rounds; ++round) {
        bbcp_opts o;
        // flags and bitmaps only. (BBCP_OPT_STORE_SETS -- keeping the sets on the device so that probes adopt each other's
        // results -- was measured here and is NOT used: on the Tseitin shape a pass writes 64 GB of sets and the first pass of
        // a handle runs twice to size the pool: 8.6 s per round against 3.4 s with conflict / overflow inheritance alone)
        memset(&o, 0, sizeof(o));
        if ((rc = bbcp_clear_bitmaps(h)) != BBCP_OK) return rc;
        if ((rc = bbcp_probe_all(h, &o, nullptr)) != BBCP_OK) return rc;
        unit.assign(bbcp_bitmap_words(h), 0);
        if ((rc = bbcp_fetch_bitmaps(h, nullptr, unit.data())) != BBCP_OK) return rc;
        cube.clear();
        for (size_t w = 0; w < unit.size(); ++w)
            for (uint32_t m = unit[w]; m; m &= m - 1) {
                const uint32_t il = (uint32_t)(w * 32) + (uint32_t)__builtin_ctz(m);
                cube.push_back((il & 1) ? -(int32_t)(il >> 1) : (int32_t)(il >> 1));
            }
        if (rounds_out) *rounds_out = round + 1;
        if (cube.empty()) break;
        total_units += cube.size();
        if (units_out) *units_out = total_units;
        try { h->unit_log.insert(h->unit_log.end(), cube.begin(), cube.end()); }
        catch (const std::bad_alloc &) { return h->fail(BBCP_ERR_ALLOC, "host allocation failed (unit log)"); }
        // the units join the clause stream either way, so that a later bbcp_finalize rebuilds the simplified database
        try {
            for (int32_t l : cube) { const int32_t c[2] = {l, 0}; h->db.add_words(c, 2); }
        } catch (const std::bad_alloc &) { return h->fail(BBCP_ERR_ALLOC, "host allocation failed (unit clauses)"); }
        if (rebuild) {
            h->finalized = false;
            h->prepared = false;
            rc = bbcp_finalize(h);
            if (rc != BBCP_OK) return rc;   // BBCP_CONTRADICTORY: x and not-x both failed, or the units propagate to a conflict
            continue;
        }
        // ON THE DEVICE: the new units are one cube; its fixpoint is what level 0 gains (Topi.cc:1311-1333: Assign at level
        // 0 + BCP). The rows are NOT rebuilt: the gained literals are marked in litinfo and every later look-up of an
        // unassigned literal consults it (value_of), which is what dropping the satisfied clauses and the false literals
        // (SimplifyIfRequired, TopiCompression.cc:509-640) would achieve. bbcp_finalize re-simplifies for real.
        const uint64_t off[2] = {0, cube.size()};
        bbcp_batch_info bi;
        if ((rc = bbcp_propagate_batch(h, cube.data(), off, 1, nullptr, &bi)) != BBCP_OK) return rc;
        uint8_t flag = 0;
        if (bbcp_fetch_flags(h, &flag)) return h->status;
        if (flag == BBCP_FLAG_CONFLICT || flag == BBCP_FLAG_FALSE_L0) { ++h->db_version; return h->fail(BBCP_CONTRADICTORY, "the failed-literal units refute the instance at level 0"); }
        if (bi.n_implied) {
            uint64_t ioff[2];
            lits.resize(bi.n_implied);
            if (bbcp_fetch_implied(h, ioff, lits.data())) return h->status;
            for (int32_t l : lits) {
                const uint32_t v = (uint32_t)(l < 0 ? -l : l);
                h->db.varinfo[v] = (h->db.varinfo[v] & ~3) | (l < 0 ? 2 : 1);
            }
            launch_mark_level0(h->d_litinfo.as<uint8_t>(), h->rs[h->rb].impl.as<int32_t>(), bi.n_implied, h->stream);
            if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "level-0 marking")) return BBCP_ERR_CUDA;
            h->l0_dirty = true;
            h->have_batch = false;
        }
    }
    return BBCP_OK;
}

// Products of the last all-literal pass (survey "next" row N4), computed on the device from its implied sets.
int bbcp_probe_products(bbcp_handle *h, uint32_t *necessary, uint64_t *n_necessary, int32_t *equiv, uint64_t equiv_cap, uint64_t *n_equiv)
{
    if (n_necessary) *n_necessary = 0;
    if (n_equiv) *n_equiv = 0;
    if (!h || !h->have_batch) return h ? h->fail(BBCP_ERR_STATE, "no batch results") : BBCP_ERR_ARG;
    if (!h->last_universe || !h->last_implied || (h->last_first & 1) || (h->last_n & 1))
        return h->fail(BBCP_ERR_STATE, "bbcp_probe_products needs the results of bbcp_probe_all / bbcp_probe_range (even bounds) with implied sets");
    if (equiv_cap && !equiv) return h->fail(BBCP_ERR_ARG, "bbcp_probe_products: null pair buffer");
    cudaSetDevice(h->device);
    const uint64_t words = h->bitmap_words;
    const size_t bytes = 16 + words * 4 + equiv_cap * 8;
    if (!h->d_products.reserve(bytes)) return h->fail(BBCP_ERR_ALLOC, "device allocation failed (probing products)");
    cudaMemsetAsync(h->d_products.p, 0, 16 + words * 4, h->stream);
    unsigned long long *counts = h->d_products.as<unsigned long long>();
    uint32_t *d_nec = reinterpret_cast<uint32_t *>(counts + 2);
    int32_t *d_eq = reinterpret_cast<int32_t *>(d_nec + words);
    launch_products(h->last_off32, h->rs[h->rb].flag.as<uint8_t>(), h->rs[h->rb].off.p, h->rs[h->rb].impl.as<int32_t>(), h->last_n / 2, h->last_first / 2,
                    d_nec, d_eq, equiv_cap, counts, h->sm_count, h->stream);
    unsigned long long hc[2] = {0, 0};
    cudaMemcpyAsync(hc, counts, 16, cudaMemcpyDeviceToHost, h->stream);
    if (necessary) cudaMemcpyAsync(necessary, d_nec, words * 4, cudaMemcpyDeviceToHost, h->stream);
    if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "probing products")) return BBCP_ERR_CUDA;
    const uint64_t ne = std::min<uint64_t>(hc[1], equiv_cap);
    if (ne) {
        cudaMemcpyAsync(equiv, d_eq, ne * 8, cudaMemcpyDeviceToHost, h->stream);
        if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "probing products")) return BBCP_ERR_CUDA;
    }
    if (n_necessary) *n_necessary = hc[0];
    if (n_equiv) *n_equiv = hc[1];   // may exceed equiv_cap: call again with a larger buffer
    return BBCP_OK;
}

// The static no-missed-implication check (TopiWL.cc:330-442) over the last batch's fixpoints, on the device.
int bbcp_check_fixpoints(bbcp_handle *h, uint64_t first_item, uint64_t count, uint64_t stride, uint64_t *violations, uint64_t *clauses_checked, uint64_t *items_checked)
{
    if (violations) *violations = 0;
    if (clauses_checked) *clauses_checked = 0;
    if (items_checked) *items_checked = 0;
    if (!h || !h->have_batch) return h ? h->fail(BBCP_ERR_STATE, "no batch results") : BBCP_ERR_ARG;
    if (!h->last_implied) return h->fail(BBCP_ERR_STATE, "bbcp_check_fixpoints needs a batch run with implied sets");
    if (stride == 0) stride = 1;
    if (first_item >= h->last_n || count == 0) return BBCP_OK;
    count = std::min<uint64_t>(count, (h->last_n - first_item + stride - 1) / stride);
    cudaSetDevice(h->device);
    const DbView dbv = db_view(h);
    const int blocks = (int)std::min<uint64_t>(count, (uint64_t)h->sm_count * 4);
    // scratch: one dense assignment per CTA + three counters (the products buffer doubles as scratch)
    const size_t words = (size_t)blocks * dbv.nvars_words16;
    if (!h->d_products.reserve(words * 4 + 32)) return h->fail(BBCP_ERR_ALLOC, "device allocation failed (fixpoint checker)");
    cudaMemsetAsync(h->d_products.p, 0, words * 4 + 32, h->stream);
    unsigned long long *out = reinterpret_cast<unsigned long long *>(h->d_products.as<uint32_t>() + ((words + 1) & ~(size_t)1));
    launch_check_fixpoints(dbv, h->last_off32, h->rs[h->rb].flag.as<uint8_t>(), h->rs[h->rb].off.p, h->rs[h->rb].impl.as<int32_t>(), first_item, count, stride,
                           h->d_products.as<uint32_t>(), blocks, out, h->stream);
    unsigned long long res[3] = {0, 0, 0};
    cudaMemcpyAsync(res, out, 24, cudaMemcpyDeviceToHost, h->stream);
    if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "fixpoint checker")) return BBCP_ERR_CUDA;
    if (violations) *violations = res[0];
    if (clauses_checked) *clauses_checked = res[1];
    if (items_checked) *items_checked = res[2];
    return BBCP_OK;
}

// Hyper-binary resolvents of an all-literal pass (survey "next" row N4; not in the reference: defined on the implied sets
// the reference would produce). Two passes on the device -- the full one and one through the binary clauses only -- and the
// per-probe set difference.
int bbcp_probe_hbr(bbcp_handle *h, int32_t *pairs, uint64_t pair_cap, uint64_t *n_pairs)
{
    if (n_pairs) *n_pairs = 0;
    int rc = need_final(h);
    if (rc) return rc;
    if (pair_cap && !pairs) return h->fail(BBCP_ERR_ARG, "bbcp_probe_hbr: null pair buffer");
    if (h->l0_dirty) return h->fail(BBCP_ERR_STATE, "bbcp_probe_hbr: level-0 facts are newer than the rows (bbcp_finalize first)");
    cudaSetDevice(h->device);
    const uint64_t n = 2ull * (uint64_t)h->db.max_var;
    bbcp_opts o;
    memset(&o, 0, sizeof(o));
    o.flags = BBCP_OPT_IMPLIED;
    h->rb = 1;                                  // the full pass keeps its results in the second result set
    rc = run_batch(h, nullptr, nullptr, 0, n, &o, nullptr);
    h->rb = 0;
    if (rc != BBCP_OK) return rc;
    o.flags = BBCP_OPT_IMPLIED | BBCP_OPT_BINARY_ONLY;
    rc = run_batch(h, nullptr, nullptr, 0, n, &o, nullptr);
    if (rc != BBCP_OK) return rc;
    if (!h->d_products.reserve(16 + pair_cap * 8)) return h->fail(BBCP_ERR_ALLOC, "device allocation failed (hyper-binary resolvents)");
    cudaMemsetAsync(h->d_products.p, 0, 16, h->stream);
    unsigned long long *count = h->d_products.as<unsigned long long>();
    int32_t *d_pairs = reinterpret_cast<int32_t *>(count + 2);
    launch_hbr(h->rs[1].flag.as<uint8_t>(), h->rs[1].off.as<uint64_t>(), h->rs[1].impl.as<int32_t>(), h->rs[0].flag.as<uint8_t>(), h->rs[0].off.as<uint64_t>(),
               h->rs[0].impl.as<int32_t>(), n, 0, d_pairs, pair_cap, count, h->sm_count, h->stream);
    unsigned long long hc = 0;
    cudaMemcpyAsync(&hc, count, 8, cudaMemcpyDeviceToHost, h->stream);
    if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "hyper-binary resolvents")) return BBCP_ERR_CUDA;
    const uint64_t ne = std::min<uint64_t>(hc, pair_cap);
    if (ne) {
        cudaMemcpyAsync(pairs, d_pairs, ne * 8, cudaMemcpyDeviceToHost, h->stream);
        if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "hyper-binary resolvents")) return BBCP_ERR_CUDA;
    }
    if (n_pairs) *n_pairs = hc;     // may exceed pair_cap: call again with a larger buffer
    h->have_batch = false;          // (the device holds the binary-only pass: not something to fetch as "the last batch")
    return BBCP_OK;
}

int bbcp_fetch_flags(bbcp_handle *h, uint8_t *flag)
{
    if (!h || !h->have_batch) return h ? h->fail(BBCP_ERR_STATE, "no batch results") : BBCP_ERR_ARG;
    cudaSetDevice(h->device);
    if (h->last_n && !h->cuda_ok(cudaMemcpyAsync(flag, h->rs[h->rb].flag.p, h->last_n, cudaMemcpyDeviceToHost, h->stream), "fetch flags")) return BBCP_ERR_CUDA;
    return h->cuda_ok(cudaStreamSynchronize(h->stream), "fetch flags") ? BBCP_OK : BBCP_ERR_CUDA;
}

int bbcp_fetch_implied32(bbcp_handle *h, uint32_t *impl_off, int32_t *impl_lits)
{
    if (!h || !h->have_batch) return h ? h->fail(BBCP_ERR_STATE, "no batch results") : BBCP_ERR_ARG;
    if (!h->last_off32) return h->fail(BBCP_ERR_STATE, "the last batch was not run with BBCP_OPT_OFF32");
    cudaSetDevice(h->device);
    if (!h->cuda_ok(cudaMemcpyAsync(impl_off, h->rs[h->rb].off.p, (h->last_n + 1) * 4, cudaMemcpyDeviceToHost, h->stream), "fetch offsets")) return BBCP_ERR_CUDA;
    if (h->last_n_implied && !h->cuda_ok(cudaMemcpyAsync(impl_lits, h->rs[h->rb].impl.p, h->last_n_implied * 4, cudaMemcpyDeviceToHost, h->stream), "fetch implied")) return BBCP_ERR_CUDA;
    return h->cuda_ok(cudaStreamSynchronize(h->stream), "fetch implied") ? BBCP_OK : BBCP_ERR_CUDA;
}

int bbcp_fetch_implied(bbcp_handle *h, uint64_t *impl_off, int32_t *impl_lits)
{
    if (!h || !h->have_batch) return h ? h->fail(BBCP_ERR_STATE, "no batch results") : BBCP_ERR_ARG;
    if (h->last_off32) return h->fail(BBCP_ERR_STATE, "the last batch was run with BBCP_OPT_OFF32: use bbcp_fetch_implied32");
    cudaSetDevice(h->device);
    if (!h->cuda_ok(cudaMemcpyAsync(impl_off, h->rs[h->rb].off.p, (h->last_n + 1) * 8, cudaMemcpyDeviceToHost, h->stream), "fetch offsets")) return BBCP_ERR_CUDA;
    if (h->last_n_implied && !h->cuda_ok(cudaMemcpyAsync(impl_lits, h->rs[h->rb].impl.p, h->last_n_implied * 4, cudaMemcpyDeviceToHost, h->stream), "fetch implied")) return BBCP_ERR_CUDA;
    return h->cuda_ok(cudaStreamSynchronize(h->stream), "fetch implied") ? BBCP_OK : BBCP_ERR_CUDA;
}

uint64_t bbcp_bitmap_words(const bbcp_handle *h) { return h ? (2 * ((uint64_t)h->db.max_var + 1) + 31) / 32 : 0; }

int bbcp_fetch_bitmaps(bbcp_handle *h, uint32_t *failed, uint32_t *unit)
{
    int rc = need_final(h);
    if (rc) return rc;
    cudaSetDevice(h->device);
    if (failed) cudaMemcpyAsync(failed, h->failed_p(), h->bitmap_words * 4, cudaMemcpyDeviceToHost, h->stream);
    if (unit) cudaMemcpyAsync(unit, h->unit_p(), h->bitmap_words * 4, cudaMemcpyDeviceToHost, h->stream);
    return h->cuda_ok(cudaStreamSynchronize(h->stream), "fetch bitmaps") ? BBCP_OK : BBCP_ERR_CUDA;
}

int bbcp_clear_bitmaps(bbcp_handle *h)
{
    int rc = need_final(h);
    if (rc) return rc;
    cudaSetDevice(h->device);
    if (!drain_peers(h)) return h->status < 0 ? h->status : BBCP_ERR_STATE;
    h->bitmaps_dirty = false;
    cudaMemsetAsync(h->d_bitmaps.p, 0, 2 * h->bm_stride * 4, h->stream);
    return h->cuda_ok(cudaStreamSynchronize(h->stream), "clear bitmaps") ? BBCP_OK : BBCP_ERR_CUDA;
}

void bbcp_debug_phase_ticks(bbcp_handle *h, uint64_t *out8)
{
    if (!h || !out8) return;
    if (h->device_ready) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); }
    unsigned long long t[8];
    read_phase_ticks(t);
    for (int i = 0; i < 8; ++i) out8[i] = t[i];
}

int bbcp_l2_flush(bbcp_handle *h, uint64_t bytes)
{
    if (!h) return BBCP_ERR_ARG;
    if (!ensure_device(h)) return h->status;
    cudaSetDevice(h->device);
    if (bytes == 0) bytes = 256ull << 20;
    if (!h->d_flush.reserve(bytes)) return h->fail(BBCP_ERR_ALLOC, "device allocation failed (flush buffer)");
    return h->cuda_ok(cudaMemsetAsync(h->d_flush.p, (int)(++h->flush_count & 0xff), bytes, h->stream), "l2 flush") ? BBCP_OK : BBCP_ERR_CUDA;
}

void *bbcp_device_failed_bitmap(bbcp_handle *h) { return h ? h->failed_p() : nullptr; }
void *bbcp_device_unit_bitmap(bbcp_handle *h) { return h ? h->unit_p() : nullptr; }
void *bbcp_device_flags(bbcp_handle *h) { return h && h->have_batch ? h->rs[h->rb].flag.p : nullptr; }
void *bbcp_device_impl_off(bbcp_handle *h) { return h && h->have_batch ? h->rs[h->rb].off.p : nullptr; }
void *bbcp_device_impl_lits(bbcp_handle *h) { return h && h->have_batch ? h->rs[h->rb].impl.p : nullptr; }
void *bbcp_stream(bbcp_handle *h) { return h ? (void *)h->stream : nullptr; }
int bbcp_sync(bbcp_handle *h)
{
    if (!h || !h->device_ready) return BBCP_ERR_STATE;
    cudaSetDevice(h->device);
    return h->cuda_ok(cudaStreamSynchronize(h->stream), "sync") ? BBCP_OK : BBCP_ERR_CUDA;
}

// ---------------------------------------------------------------- multi-GPU merge over NVLink peer memory
int bbcp_peer_export(bbcp_handle *h, void *out)
{
    int rc = need_final(h);
    if (rc) return rc;
    if (!out) return h->fail(BBCP_ERR_ARG, "bbcp_peer_export: null buffer");
    cudaSetDevice(h->device);
    static_assert(sizeof(cudaIpcMemHandle_t) <= BBCP_PEER_HANDLE_BYTES - 16, "IPC handle size");
    memset(out, 0, BBCP_PEER_HANDLE_BYTES);
    cudaIpcMemHandle_t ipc;
    if (!h->cuda_ok(cudaIpcGetMemHandle(&ipc, h->d_bitmaps.p), "cudaIpcGetMemHandle")) return BBCP_ERR_CUDA;
    memcpy(out, &ipc, sizeof(ipc));
    memcpy(static_cast<char *>(out) + BBCP_PEER_HANDLE_BYTES - 16, &h->bm_stride, 8);
    return BBCP_OK;
}

int bbcp_peer_detach(bbcp_handle *h)
{
    if (!h) return BBCP_ERR_ARG;
    if (h->device_ready) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); drain_peers(h); }
    for (int r = 0; r < (int)h->peer_base.size(); ++r)
        if (r != h->peer_rank && h->peer_base[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
    cudaGetLastError();
    h->peer_base.clear();
    h->peer_rank = -1;
    h->peer_world = 0;
    return BBCP_OK;
}

int bbcp_peer_attach(bbcp_handle *h, int rank, int world, const void *all)
{
    int rc = need_final(h);
    if (rc) return rc;
    if (!all || world < 1 || world > 64 || rank < 0 || rank >= world) return h->fail(BBCP_ERR_ARG, "bbcp_peer_attach: bad rank / world (at most 64 ranks)");
    cudaSetDevice(h->device);
    bbcp_peer_detach(h);
    h->peer_base.assign(world, nullptr);
    for (int r = 0; r < world; ++r) {
        const char *e = static_cast<const char *>(all) + (size_t)r * BBCP_PEER_HANDLE_BYTES;
        uint64_t stride = 0;
        memcpy(&stride, e + BBCP_PEER_HANDLE_BYTES - 16, 8);
        if (stride != h->bm_stride) { bbcp_peer_detach(h); return h->fail(BBCP_ERR_ARG, "bbcp_peer_attach: rank " + std::to_string(r) + " holds a different database (bitmap size differs)"); }
        if (r == rank) { h->peer_base[r] = h->d_bitmaps.p; continue; }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, e, sizeof(ipc));
        void *p = nullptr;
        if (!h->cuda_ok(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (peer bitmaps)")) { h->peer_rank = rank; bbcp_peer_detach(h); return BBCP_ERR_CUDA; }
        h->peer_base[r] = p;
    }
    h->peer_rank = rank;
    h->peer_world = world;
    h->drained_epoch = h->merge_epoch;   // nothing outstanding with the new peers
    if (!h->d_peer_tab.reserve(64 * sizeof(void *))) return h->fail(BBCP_ERR_ALLOC, "device allocation failed (peer table)");
    cudaMemcpyAsync(h->d_peer_tab.p, h->peer_base.data(), world * sizeof(void *), cudaMemcpyHostToDevice, h->stream);
    for (auto &e : h->ev_merge) if (!e && !h->cuda_ok(cudaEventCreate(&e), "cudaEventCreate")) return BBCP_ERR_CUDA;
    return h->cuda_ok(cudaStreamSynchronize(h->stream), "peer table upload") ? BBCP_OK : BBCP_ERR_CUDA;
}

int bbcp_merge_bitmaps(bbcp_handle *h, uint64_t words_per_rank, float *ms)
{
    int rc = need_final(h);
    if (rc) return rc;
    if (h->peer_world < 1) return h->fail(BBCP_ERR_STATE, "bbcp_merge_bitmaps: no peers attached");
    if (words_per_rank & 3) return h->fail(BBCP_ERR_ARG, "bbcp_merge_bitmaps: words_per_rank must be a multiple of 4");
    cudaSetDevice(h->device);
    ++h->merge_epoch;
    cudaEventRecord(h->ev_merge[0], h->stream);
    launch_merge(MergeArgs{static_cast<uint32_t *const *>(h->d_peer_tab.p), h->peer_rank, h->peer_world, h->bm_stride, words_per_rank, h->merge_epoch, h->merge_timeout_ms, 1u},
                 h->sm_count, h->stream);
    cudaEventRecord(h->ev_merge[1], h->stream);
    if (!h->cuda_ok(cudaStreamSynchronize(h->stream), "bitmap merge")) return BBCP_ERR_CUDA;
    if (merge_timed_out(h)) return h->fail(BBCP_ERR_STATE, "bbcp_merge_bitmaps: a peer did not arrive within the time-out (BBCP_MERGE_TIMEOUT_MS)");
    if (ms) cudaEventElapsedTime(ms, h->ev_merge[0], h->ev_merge[1]);
    return BBCP_OK;
}

int bbcp_assume(bbcp_handle *h, int32_t lit)
{
    if (!h) return BBCP_ERR_ARG;
    if (lit == 0 || lit == INT32_MIN) return h->fail(BBCP_ERR_ARG, "bad assumption literal");
    h->assumps.push_back(lit);
    return BBCP_OK;
}

int bbcp_solve(bbcp_handle *h)
{
    if (!h) return 0;
    std::vector<int32_t> cube;
    cube.swap(h->assumps);
    if (h->status == BBCP_OK && !h->finalized) bbcp_finalize(h);
    if (h->status == BBCP_CONTRADICTORY) return 20;
    if (need_final(h)) return 0;
    h->solve_val.assign((size_t)h->db.max_var + 1, 0);
    h->solved = false;
    // assumptions on variables that occur in no clause are free variables (the reference creates them in
    // Solve, Topi.cc:1117-1125): they cannot propagate, only contradict each other
    std::vector<int32_t> dev_cube;
    std::vector<std::pair<int32_t, int8_t>> free_vals;
    for (int32_t l : cube) {
        const int32_t v = l < 0 ? -l : l;
        if (v <= h->db.max_var && (h->db.varinfo[v] & 4)) { dev_cube.push_back(l); continue; }
        const int8_t want = l < 0 ? -1 : 1;
        for (auto &fv : free_vals) if (fv.first == v && fv.second != want) return 20;
        free_vals.emplace_back(v, want);
    }
    const uint64_t off[2] = {0, dev_cube.size()};
    bbcp_batch_info bi;
    if (bbcp_propagate_batch(h, dev_cube.data(), off, 1, nullptr, &bi)) return 0;
    uint8_t flag = 0;
    if (bbcp_fetch_flags(h, &flag)) return 0;
    if (flag == BBCP_FLAG_CONFLICT || flag == BBCP_FLAG_FALSE_L0) return 20;
    std::vector<uint64_t> ioff(2);
    std::vector<int32_t> lits(std::max<uint64_t>(bi.n_implied, 1));
    if (bbcp_fetch_implied(h, ioff.data(), lits.data())) return 0;
    for (uint64_t i = 0; i < bi.n_implied; ++i) h->solve_val[lits[i] < 0 ? -lits[i] : lits[i]] = lits[i] < 0 ? -1 : 1;
    for (auto &fv : free_vals) if (fv.first <= h->db.max_var) h->solve_val[fv.first] = fv.second;
    h->solved = true;
    return bi.n_implied + h->db.n_level0() == h->db.n_mapped() ? 10 : 0;
}

int bbcp_val(bbcp_handle *h, int32_t lit)
{
    if (!h || lit == 0 || lit == INT32_MIN) return 0;
    const int32_t v = lit < 0 ? -lit : lit;
    const int sign = lit < 0 ? -1 : 1;
    if (v > h->db.max_var || h->db.varinfo.size() <= (size_t)v) return 2;
    const uint8_t vi = h->db.varinfo[v];
    if (vi & 3) return ((vi & 3) == 1 ? 1 : -1) * sign;
    if (h->solved && h->solve_val[v]) return h->solve_val[v] * sign;
    return (vi & 4) ? 0 : 2;
}

int bbcp_declevel(bbcp_handle *h, int32_t lit)
{
    if (!h || lit == 0 || lit == INT32_MIN) return -1;
    const int32_t v = lit < 0 ? -lit : lit;
    if (v > h->db.max_var || h->db.varinfo.size() <= (size_t)v) return -1;
    if (h->db.varinfo[v] & 3) return 0;
    if (h->solved && h->solve_val[v]) return 1;
    return -1;
}

int bbcp_backtrack(bbcp_handle *h, int32_t level)
{
    if (!h) return BBCP_ERR_ARG;
    if (level <= 0) { h->solved = false; }
    return BBCP_OK;
}

uint64_t bbcp_propagations(const bbcp_handle *h) { return h ? h->total_props : 0; }

}  // extern "C"
