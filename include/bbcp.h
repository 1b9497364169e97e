/*
 * bbcp.h -- C ABI of the B200 batched Boolean-constraint-propagation engine (libbbcp.so).
 *
 * This is the drop-in boundary for ONE path of the reference solver (IntelSAT / "Topor"): propagating
 * many independent probes or cubes to their unit-propagation fixpoint over one shared, read-only clause
 * database. Every entry point names the reference interface it stands in for (file:line relative to the
 * reference tree). Conventions follow the reference's own C ABI (ToporIpasir.cc:118-178): opaque handle,
 * int32 DIMACS literals, integer return codes, no exceptions, sticky status, caller-owned buffers.
 *
 * There is no CPU propagation path behind this ABI: every propagate/probe/solve call runs the CUDA
 * kernels (sm_100a) and fails with BBCP_ERR_CUDA when no device is usable.
 */
#ifndef BBCP_H
#define BBCP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bbcp_handle bbcp_handle;

/* return / status codes. Sticky ones mirror TToporStatus (ToporExternalTypes.hpp:13-49):
 * CONTRADICTORY ~ STATUS_CONTRADICTORY, ERR_ALLOC ~ STATUS_ALLOC_FAILED,
 * ERR_INDEX_TOO_NARROW ~ STATUS_INDEX_TOO_NARROW (TopiConflictAnalysis.cc:67-71). */
enum {
    BBCP_OK = 0,
    BBCP_CONTRADICTORY = 1,         /* the clause database is unsatisfiable by unit propagation at level 0 */
    BBCP_ERR_ALLOC = -1,
    BBCP_ERR_CUDA = -2,             /* no device / CUDA failure (message in bbcp_error_string) */
    BBCP_ERR_INDEX_TOO_NARROW = -3, /* database exceeds the 32-bit slot / clause indices */
    BBCP_ERR_STATE = -4,            /* call not valid in this state (e.g. propagate before finalize) */
    BBCP_ERR_ARG = -5
};

/* per probe / per cube result flag */
enum {
    BBCP_FLAG_OK = 0,          /* fixpoint reached without conflict; implied set valid */
    BBCP_FLAG_CONFLICT = 1,    /* unit propagation derives a conflict (a failed literal for a depth-1 probe) */
    BBCP_FLAG_TRIVIAL_L0 = 2,  /* every literal of the cube is already true at level 0 (Topi.cc:851-865) */
    BBCP_FLAG_FALSE_L0 = 3,    /* some literal of the cube is false at level 0 (Topi.cc:853-860) */
    BBCP_FLAG_UNMAPPED = 4,    /* the cube names a variable that occurs in no kept clause; not propagated */
    BBCP_FLAG_SELF = 0x80      /* only with BBCP_OPT_COMPACT_SELF: fixpoint reached and the implied set is exactly the probe's
                                  own literal; that one-literal set is NOT stored (its range in impl_off is empty) */
};

/* option bits */
enum {
    BBCP_OPT_IMPLIED = 1u,     /* produce the per-probe implied-literal sets (else flags + bitmaps only) */
    BBCP_OPT_NO_LEN_PRUNE = 2u,/* diagnostic: scan ternary/long rows even while only one literal is assigned */
    BBCP_OPT_KERNEL_TIMES = 8u,/* record CUDA events between the kernels and fill triage_ms / deep_ms / block_ms (a few
                                  microseconds of stream time per batch; total_ms is always measured) */
    BBCP_OPT_MERGE = 16u,      /* bbcp_probe_range with peers attached: the all-gather of this range's slice of the failed /
                                  unit bitmaps (bbcp_merge_bitmaps) runs as the last kernel of the batch, no host round trip */
    BBCP_OPT_RENDEZVOUS = 32u, /* bbcp_probe_range with peers attached: wait on the device for every peer before the batch
                                  starts (aligns the ranks' device clocks; every rank must pass the same flags) */
    BBCP_OPT_NO_REUSE = 64u,   /* diagnostic: all-literal batches propagate every probe from scratch (no adoption of the
                                  finished probes' results, DESIGN.md "closure reuse"); results are identical */
    BBCP_OPT_COMPACT_SELF = 128u, /* compact read-back for depth-1 probes: a probe that implies nothing but itself is reported as
                                  BBCP_FLAG_SELF and its set is left out of impl_off / impl_lits (9 of the 13 bytes a probe
                                  costs on the way back are that one literal and its offset). bbcp_probe_lits_to_host then
                                  copies offsets back only from the first chunk on that stores a set: when info.n_implied
                                  == 0 the offsets are all zero and are not written at all */
    BBCP_OPT_BINARY_ONLY = 256u, /* propagate through the binary clauses only: the closure of the binary implication graph
                                  (what bbcp_probe_hbr compares the full pass with) */
    BBCP_OPT_STORE_SETS = 512u, /* without BBCP_OPT_IMPLIED: keep the implied sets on the device for the duration of the batch all the
                                  same, so that probes adopt each other's finished results (closure reuse); only flags and bitmaps
                                  come back. Costs the pool memory, saves most of the propagation on deep instances */
    BBCP_OPT_OFF32 = 4u        /* produce impl_off as uint32 (fetch with bbcp_fetch_implied32): halves the offset
                                  read-back; the batch fails with BBCP_ERR_ARG if it implies >= 2^32 literals */
};

typedef struct {
    uint32_t flags;            /* BBCP_OPT_* ; default BBCP_OPT_IMPLIED */
    uint32_t small_trail;      /* trail capacity of the first tier, shared memory (0 = 512) */
    uint64_t out_capacity;     /* device capacity for implied literals per launch, in literals (0 = auto) */
    uint32_t mid_trail;        /* trail capacity of the second tier, hash + trail in a global slab (0 = 1024 for all-literal
                                  batches, where finished probes' sets are adopted and the last tier does that faster; 8192 otherwise) */
    uint32_t merge_words_per_rank; /* BBCP_OPT_MERGE: bitmap words owned by each rank (as in bbcp_merge_bitmaps) */
} bbcp_opts;

typedef struct {
    uint64_t n;                /* probes / cubes in the batch */
    uint64_t n_implied;        /* total implied literals (length of impl_lits) */
    uint64_t n_ok, n_conflict, n_skipped; /* flag 0 / flag 1 / flags 2,3,4 */
    uint64_t n_large;          /* probes that ran in the third tier (one CTA per probe, dense state in a global slab) */
    uint64_t props_all;        /* literals dequeued and propagated (every trail literal of every probe) */
    uint64_t props_ref;        /* of those, literals whose negation has a non-empty row: the reference's
                                  m_Implications convention (TopiBcp.cc:193-198) */
    uint64_t slots;            /* 8-byte row slots examined */
    uint64_t clause_fetches;   /* long clauses fetched from the arena */
    uint64_t kernel_bytes;     /* algorithmic bytes moved by triage + propagation ("bytes model") */
    uint64_t triage_bytes;     /* of those, the triage kernel's */
    uint64_t n_deep;           /* items the triage queued for propagation */
    float kernel_ms;           /* triage_ms + deep_ms + block_ms (per-kernel times need BBCP_OPT_KERNEL_TIMES) */
    float total_ms;            /* device time of the whole call incl. result compaction (CUDA events) */
    uint32_t launches;         /* kernels of this engine launched by the call (library scan kernels not counted) */
    float tier1_ms;            /* triage_ms + deep_ms */
    float triage_ms;           /* device time of k_triage (last attempt) */
    float deep_ms;             /* device time of k_deep, the first tier: warp per probe, state in shared memory (last attempt) */
    float block_ms;            /* device time of the second and third tiers + the sorting of second-tier sets (last attempt) */
    float mid_ms;              /* of which the second tier (k_deep2) */
    uint64_t n_mid;            /* probes that ran in the second tier (warp per probe, hash + trail in a global slab) */
    float compact_ms;          /* device time of k_scan_gather (+ k_gather_long), last sequence */
    float merge_ms;            /* device time of the in-batch bitmap all-gather (BBCP_OPT_MERGE), incl. waiting for peers */
    /* closure reuse (all-literal batches): finished probes whose implied set was adopted by a later probe, the literals
     * read back that way, and probes decided by a finished probe's conflict / tier overflow */
    uint64_t adopted, adopt_lits, inherited;
    /* the reference's other two counters on this path: Assign calls (m_Assignments, TopiAsg.cc:49: every literal put on
     * a trail, decisions included) and BCP() calls (m_BCPs, TopiBcp.cc:176: one per decision level opened, i.e. per
     * probe, and per cube literal that was still unassigned when its turn came -- see DESIGN.md on what is exact) */
    uint64_t assignments, bcps;
    /* the order-independent parts, comparable with the reference run on the same items: literals put on the trails of the
     * items that reached a fixpoint (= n_implied of a batch of probes), and props_ref restricted to those items */
    uint64_t assignments_ok, props_ref_ok;
    uint64_t hot_rows;         /* rows left unwalked because they were several times longer than everything the probe had walked so far
                                  (the earlier members were re-walked instead; DESIGN.md "hot rows") */
} bbcp_batch_info;

typedef struct {
    int32_t max_var;           /* probe universe = +1,-1,...,+max_var,-max_var */
    uint64_t n_mapped_vars, n_level0;
    uint64_t n_bin, n_tern, n_long, n_long_lits;   /* clauses kept after level-0 simplification */
    uint64_t row_slots, arena_words;
    uint64_t device_bytes;
    uint32_t db_version;
    uint32_t reserved;
} bbcp_db_info;

/* ---- lifetime: CTopor ctor/dtor (Topor.hpp:31, Topor.cc:12,29-33); ipasir_init/release (ToporIpasir.cc:118-131) */
bbcp_handle *bbcp_create(int32_t nvars_hint);
void bbcp_release(bbcp_handle *h);
/* choose the CUDA device before finalize (default: the current device of the calling thread) */
int bbcp_set_device(bbcp_handle *h, int device);

/* ---- clause intake: CTopor::AddClause (Topor.hpp:33-38) -> CTopi::AddUserClause (Topi.cc:377-653);
 *      ipasir_add (ToporIpasir.cc:14-25). 0 terminates a clause; reading stops at the first 0. */
int bbcp_add(bbcp_handle *h, int32_t lit_or_zero);
int bbcp_add_clause(bbcp_handle *h, const int32_t *lits, size_t n);
int bbcp_add_clauses(bbcp_handle *h, const int32_t *words, uint64_t nwords); /* stream of 0-terminated clauses */

/* ---- finalize: what Solve does before search (Topi.cc:1127-1160: ReserveVarAndLitData + level-0 BCP),
 *      here: normalise (A12), level-0 unit propagation ON THE DEVICE, level-0 simplification, build the
 *      CSR rows + clause arena, upload. Returns BBCP_OK or BBCP_CONTRADICTORY (sticky) or an error. */
int bbcp_finalize(bbcp_handle *h);
/* host-only half of finalize (normalise + build rows, no level-0 propagation, no device): lets CPU-only
 * tests inspect the database layout. */
int bbcp_host_prepare(bbcp_handle *h);
int bbcp_get_db_info(const bbcp_handle *h, bbcp_db_info *out);
/* external literals assigned at level 0, ascending |var|; returns the count */
uint64_t bbcp_level0_lits(const bbcp_handle *h, int32_t *out, uint64_t cap);
/* row of internal literal ilit (= 2*var + neg) as built on the host: copies up to cap words of
 * [nbin, ntern, nlong, bins..., tern pairs..., long (blocker, clause lits...)...]; returns words needed */
uint64_t bbcp_host_row(const bbcp_handle *h, uint32_t ilit, uint32_t *out, uint64_t cap);

/* the probe schedule of the prepared database (DESIGN.md "closure reuse"): every internal literal 2..2*max_var+1,
 * sorted so that q comes before p whenever the binary clause (not-p or q) exists and p, q are not equivalent; empty (0)
 * for a database without binary clauses. Copies up to cap entries, returns the count. Host only. */
uint64_t bbcp_host_order(bbcp_handle *h, uint32_t *out, uint64_t cap);

/* ---- status: CTopor::IsError / GetStatusExplanation (Topor.hpp:111-114) */
int bbcp_status(const bbcp_handle *h);
const char *bbcp_error_string(const bbcp_handle *h);

/* ---- the hot path: batched BCP to fixpoint. Stands in for one
 *      NewDecLevel/Assign/BCP()/Backtrack(0) round of the reference per probe (TopiBcp.cc:103-410,
 *      TopiAsg.cc:46-139, TopiBacktrack.cc:18-53) and for HandleAssumptions/AssignAssumptions per cube
 *      (Topi.cc:797-938, 715-769). Host buffers in; results stay on the device until fetched. */
int bbcp_propagate_batch(bbcp_handle *h, const int32_t *lits, const uint64_t *off, uint64_t ncubes,
                         const bbcp_opts *opts, bbcp_batch_info *info);
/* probe universe indices [first, last): index i is literal (i/2+1) * (i odd ? -1 : +1) */
int bbcp_probe_range(bbcp_handle *h, uint64_t first, uint64_t last, const bbcp_opts *opts, bbcp_batch_info *info);
int bbcp_probe_all(bbcp_handle *h, const bbcp_opts *opts, bbcp_batch_info *info);
/* explicit list of depth-1 probes: item i is the single literal lits[i] (4 bytes per probe instead of a CSR
 * cube; a 0 entry is reported BBCP_FLAG_UNMAPPED). Same results as one-literal cubes. */
int bbcp_probe_lits(bbcp_handle *h, const int32_t *lits, uint64_t n, const bbcp_opts *opts, bbcp_batch_info *info);
/* The same, results straight into HOST buffers.
 * PINNED caller buffers (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) are used IN PLACE: the first kernel of
 * the batch reads `lits` out of the caller's memory over PCIe and writes the flags it decides into `flag_out`, so no
 * staging copy runs for either (flag_out is copied only when a probe needed the propagation tiers); offsets and
 * literals travel back on a copy stream while the next chunk computes. The buffers must stay untouched until the call
 * returns (it returns after the last byte has arrived).
 * Pageable buffers (or BBCP_E2E=dma in the environment) take the copy-engine path: the list is cut into chunks, all
 * chunks are uploaded on one stream, each chunk's kernels start when its part has arrived, and its flags / offsets /
 * literals travel back on a third stream while the next chunk computes.
 * flag_out[n]; off_out[n+1] uint64, or uint32 with BBCP_OPT_OFF32 (offsets run through the whole list);
 * impl_out[impl_cap] (BBCP_ERR_ARG if the list implies more). Bitmaps accumulate as usual. */
int bbcp_probe_lits_to_host(bbcp_handle *h, const int32_t *lits, uint64_t n, const bbcp_opts *opts, uint8_t *flag_out, void *off_out,
                            int32_t *impl_out, uint64_t impl_cap, bbcp_batch_info *info);
/* same as bbcp_propagate_batch with the cube arrays already resident in device memory (d_off == NULL: d_lits is
 * a probe list of ncubes literals) */
int bbcp_propagate_batch_device(bbcp_handle *h, const int32_t *d_lits, const uint64_t *d_off, uint64_t ncubes,
                                uint64_t nlits, const bbcp_opts *opts, bbcp_batch_info *info);

/* failed-literal probing to a fixpoint ("next" row N2 of the survey): probe all literals, add the negations of the
 * failed literals as unit clauses, re-finalize (level-0 propagation + database simplification), repeat until a
 * pass finds none (max_rounds == 0: no limit). Returns BBCP_OK, or BBCP_CONTRADICTORY if the units refute the
 * instance. *rounds = passes run, *units = failed literals turned into units; bbcp_level0_lits has the result. */
int bbcp_probe_to_fixpoint(bbcp_handle *h, uint32_t max_rounds, uint32_t *rounds, uint64_t *units);

/* the units bbcp_probe_to_fixpoint has derived so far (negations of failed literals), in derivation order: each one follows
 * by unit propagation (RUP) from the clauses plus the units before it, which is the order a DRAT trace needs
 * (TopiConflictAnalysis.cc:1117-1184 logs learnt units the same way). Copies up to cap, returns the count. */
uint64_t bbcp_unit_log(const bbcp_handle *h, int32_t *out, uint64_t cap);

/* products of the last all-literal pass with implied sets (bbcp_probe_all, or bbcp_probe_range with even bounds),
 * computed on the device ("next" row N4; not in the reference, defined on the per-probe implied sets the reference
 * would produce): necessary literals -- x implied by +v and by -v, so x is a unit -- as a bitmap over internal
 * literals (bbcp_bitmap_words() words, may be NULL), and equivalences -- +v implies x and -v implies not-x, so
 * v == x -- as pairs (v, x) of external literals in no particular order. *n_equiv is the number found, which may
 * exceed equiv_cap (then only equiv_cap pairs were stored). */
int bbcp_probe_products(bbcp_handle *h, uint32_t *necessary, uint64_t *n_necessary, int32_t *equiv /* 2 * equiv_cap */, uint64_t equiv_cap,
                        uint64_t *n_equiv);

/* diagnostic: the static analogue of the reference's WLAssertNoMissedImplications (TopiWL.cc:330-442) over the fixpoints of
 * the last batch (run with implied sets): for items first_item, first_item + stride, ... (count of them) that reached a
 * fixpoint, every clause of the database is evaluated under the item's implied set (+ level 0); *violations counts the
 * clause occurrences that are unsatisfied with fewer than two unassigned literals (a missed implication or conflict).
 * O(database) per item, on the device. */
int bbcp_check_fixpoints(bbcp_handle *h, uint64_t first_item, uint64_t count, uint64_t stride, uint64_t *violations, uint64_t *clauses_checked,
                         uint64_t *items_checked);

/* hyper-binary resolvents ("next" row N4; not in the reference, defined on the implied sets the reference would produce): for
 * every literal v whose probe reaches a fixpoint, every x implied by v but not through binary clauses alone, as pairs
 * (v, x) of external literals in no particular order -- the binary clause (not-v or x) follows from the database and is new
 * to its binary implication graph. Runs an all-literal pass and a binary-clauses-only pass on the device. *n_pairs may exceed
 * pair_cap (then only pair_cap pairs were stored). */
int bbcp_probe_hbr(bbcp_handle *h, int32_t *pairs /* 2 * pair_cap */, uint64_t pair_cap, uint64_t *n_pairs);

/* ---- results of the last batch, device -> caller buffers */
int bbcp_fetch_flags(bbcp_handle *h, uint8_t *flag /* n */);
int bbcp_fetch_implied(bbcp_handle *h, uint64_t *impl_off /* n+1 */, int32_t *impl_lits /* n_implied */);
/* the same after a batch run with BBCP_OPT_OFF32 */
int bbcp_fetch_implied32(bbcp_handle *h, uint32_t *impl_off /* n+1 */, int32_t *impl_lits /* n_implied */);
/* failed-literal bitmap (bit 2*var+neg set <=> that literal is a failed literal found by a depth-1 probe so
 * far) and unit bitmap (the negations). nwords32 = bbcp_bitmap_words(). Bitmaps accumulate over calls until
 * bbcp_clear_bitmaps. */
uint64_t bbcp_bitmap_words(const bbcp_handle *h);
int bbcp_fetch_bitmaps(bbcp_handle *h, uint32_t *failed, uint32_t *unit);
int bbcp_clear_bitmaps(bbcp_handle *h);
/* device addresses of the two bitmaps, for an on-stream NCCL all-gather / all-reduce by the caller */
void *bbcp_device_failed_bitmap(bbcp_handle *h);
void *bbcp_device_unit_bitmap(bbcp_handle *h);
/* device addresses of the last batch's results (flag u8[n]; impl_off u64[n+1]; impl_lits i32[n_implied]) */
void *bbcp_device_flags(bbcp_handle *h);
void *bbcp_device_impl_off(bbcp_handle *h);
void *bbcp_device_impl_lits(bbcp_handle *h);
/* the CUDA stream (cudaStream_t) every kernel of this handle is launched on */
void *bbcp_stream(bbcp_handle *h);
/* blocks until the handle's stream is idle */
int bbcp_sync(bbcp_handle *h);
/* diagnostic builds only (make checks; zeros otherwise): clock ticks the CTA tiers spent per phase since the last call, summed
 * over CTAs: [0] narrow-frontier mode, [1] closure look-up + adoption, [2] row walks, [3] load, [4] emission, [5] clean-up +
 * publication + ticket */
void bbcp_debug_phase_ticks(bbcp_handle *h, uint64_t *out8);
/* measurement aid: enqueue, on the handle's stream, a write of `bytes` (0 = 256 MiB, twice the L2) that evicts the L2
 * before the next batch; returns at once. Because the next call's kernels queue up behind it, the host is done launching
 * them before the device starts on them: the batch's device time does not depend on how fast the host issues launches. */
int bbcp_l2_flush(bbcp_handle *h, uint64_t bytes);

/* ---- multi-GPU: one handle per GPU (one process per GPU), the database replicated, the probe universe cut into
 *      contiguous ranges aligned to bitmap words: rank r probes the literals whose bits live in bitmap words
 *      [r*wpr, (r+1)*wpr), wpr = words_per_rank, a multiple of 4. bbcp_merge_bitmaps is the all-gather of those
 *      slices of the failed-literal and unit bitmaps, done by ONE kernel over NVLink peer memory: every rank
 *      raises a flag in every peer, waits for the flags of all peers and copies the peers' slices out of their
 *      (CUDA-IPC-mapped) bitmaps with remote loads. words_per_rank == 0: the handshake alone (a device-side
 *      barrier across the GPUs). Set-up: every rank calls bbcp_peer_export after bbcp_finalize, the
 *      BBCP_PEER_HANDLE_BYTES blobs are exchanged by the host program (e.g. a torch.distributed / MPI all-gather),
 *      every rank calls bbcp_peer_attach with all of them in rank order. A re-finalize that changes the variable
 *      range needs bbcp_peer_detach + a new exchange. Every rank must make the same sequence of merging calls.
 *      A merge OVERWRITES this rank's copy of the peers' slices (of both bitmaps) with what the peers hold -- it does
 *      not OR into it. A rank tells its peers when it has finished reading their slices; bbcp_clear_bitmaps,
 *      bbcp_finalize and bbcp_peer_detach wait for that (on the device) before they touch the bitmaps, so the caller
 *      needs no barrier between a merge and a clear. A peer that does not show up within BBCP_MERGE_TIMEOUT_MS
 *      (environment, default 20000) fails the call with BBCP_ERR_STATE; the time-out is not sticky.
 *      *ms (optional) = device time incl. the wait. */
#define BBCP_PEER_HANDLE_BYTES 128
int bbcp_peer_export(bbcp_handle *h, void *out /* BBCP_PEER_HANDLE_BYTES */);
int bbcp_peer_attach(bbcp_handle *h, int rank, int world, const void *all /* world * BBCP_PEER_HANDLE_BYTES */);
int bbcp_peer_detach(bbcp_handle *h);
int bbcp_merge_bitmaps(bbcp_handle *h, uint64_t words_per_rank, float *ms);

/* ---- single-cube convenience path with the Topor / IPASIR conventions:
 *      ipasir_assume / ipasir_solve / ipasir_val (ToporIpasir.cc:27-68), CTopor::Solve(assumps)
 *      (Topor.hpp:48), GetLitValue (Topor.hpp:80), GetLitDecLevel (Topor.hpp:88), Backtrack (Topor.hpp:68).
 *      bbcp_solve returns 20 when unit propagation of the assumptions derives a conflict, 10 when every
 *      mapped variable got assigned without conflict, else 0 (fixpoint reached; the search the reference
 *      would start here is out of scope). Assumptions are cleared by bbcp_solve as in IPASIR. */
int bbcp_assume(bbcp_handle *h, int32_t lit);
int bbcp_solve(bbcp_handle *h);
/* value of lit after the last bbcp_solve: 1 satisfied, -1 unsatisfied, 0 unassigned, 2 don't-care (variable
 * occurs in no clause): TToporLitVal (ToporExternalTypes.hpp:81-87) */
int bbcp_val(bbcp_handle *h, int32_t lit);
/* 0 for level-0 literals, 1 for literals assigned by the last solve's cube, -1 if unassigned */
int bbcp_declevel(bbcp_handle *h, int32_t lit);
int bbcp_backtrack(bbcp_handle *h, int32_t level);

/* ---- counters: CTopor::GetPropagations (Topor.hpp:107-108, Topi.cc:1427-1431): props_ref summed over all calls */
uint64_t bbcp_propagations(const bbcp_handle *h);

#ifdef __cplusplus
}
#endif
#endif
