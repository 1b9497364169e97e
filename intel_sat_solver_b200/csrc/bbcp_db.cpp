// bbcp_db.cpp -- see bbcp_db.h. Host-only code (no CUDA): unit-testable on a CPU box.
#include "bbcp_db.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace bbcp {

namespace {
unsigned host_threads(size_t work_words);
template <class F> void parallel(unsigned T, F f);
}  // namespace

void HostDb::add_words(const int32_t *w, uint64_t n)
{
    raw.insert(raw.end(), w, w + n);
    for (uint64_t i = 0; i < n; ++i) {
        // INT32_MIN cannot be negated; it is rejected by the C ABI before it gets here
        const int32_t a = w[i] < 0 ? -w[i] : w[i];
        if (a > max_var) max_var = a;
    }
}

// One pass over the user clauses in the order they were added, with the observable behaviour of
// AddUserClause (Topi.cc:377-653):
//   - an empty clause makes the instance contradictory (451-456);
//   - literals false under the units added SO FAR are dropped before anything else (487-491);
//   - duplicate literals are dropped, a tautology drops the clause and un-maps the variables that were
//     first seen in it (493-510, 434-443);
//   - all literals dropped => contradictory (552-557); one literal left => level-0 assignment, propagated
//     later (590-619); otherwise the clause is kept (621).
// A variable is "mapped" iff some kept (or unit) clause mentions it.
void HostDb::normalise()
{
    intake_by_chunk = normalise_parallel();
    if (!intake_by_chunk) normalise_serial();
}

void HostDb::normalise_serial()
{
    const size_t nv = (size_t)max_var + 1;
    varinfo.assign(nv, 0);
    cls.clear();
    units.clear();
    ncls = 0;
    contradictory = false;
    std::vector<int64_t> stamp(nv, 0);
    std::vector<uint32_t> cur, fresh;
    int64_t counter = 0;
    const size_t n = raw.size();
    size_t s = 0;
    while (s < n && !contradictory) {
        size_t e = s;
        while (e < n && raw[e] != 0) ++e;
        if (e == s) { contradictory = true; break; }
        ++counter;
        cur.clear();
        fresh.clear();
        bool taut = false;
        for (size_t i = s; i < e; ++i) {
            const int32_t l = raw[i];
            const uint32_t v = (uint32_t)(l < 0 ? -l : l);
            const uint32_t il = 2u * v + (l < 0);
            if (!(varinfo[v] & 4)) { varinfo[v] |= 4; fresh.push_back(v); }
            const uint8_t val = varinfo[v] & 3;
            if (val && ((val == 1) == (l < 0))) continue;  // false at level 0
            const int64_t me = (l < 0) ? -counter : counter;
            if (stamp[v] == me) continue;
            if (stamp[v] == -me) { taut = true; break; }
            stamp[v] = me;
            cur.push_back(il);
        }
        s = e + 1;
        if (taut) {
            for (uint32_t v : fresh) { varinfo[v] &= ~4; stamp[v] = 0; }
            continue;
        }
        if (cur.empty()) { contradictory = true; break; }
        if (cur.size() == 1) {
            const uint32_t il = cur[0], v = il >> 1;
            if ((varinfo[v] & 3) == 0) {
                varinfo[v] |= (il & 1) ? 2 : 1;
                units.push_back(il);
            }
            continue;
        }
        cls.push_back((uint32_t)cur.size());
        cls.insert(cls.end(), cur.begin(), cur.end());
        ++ncls;
    }
}

bool HostDb::simplify_level0()
{
    size_t r = 0, w = 0;
    uint64_t kept = 0;
    const size_t n = cls.size();
    while (r < n) {
        const uint32_t len = cls[r];
        const size_t hdr_pos = w++;
        uint32_t out = 0;
        bool sat = false;
        for (uint32_t i = 0; i < len; ++i) {
            const uint32_t il = cls[r + 1 + i];
            const uint8_t val = varinfo[il >> 1] & 3;
            if (val == 0) { cls[w++] = il; ++out; }
            else if ((val == 1) == ((il & 1) == 0)) sat = true;
        }
        r += 1 + len;
        if (sat) { w = hdr_pos; continue; }
        if (out < 2) return false;
        cls[hdr_pos] = out;
        ++kept;
    }
    cls.resize(w);
    ncls = kept;
    return true;
}

namespace {

unsigned host_threads(size_t work_words)
{
    unsigned t = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("BBCP_HOST_THREADS")) t = (unsigned)std::atoi(e);   // (explicit: also for small inputs)
    else if (work_words < (1u << 20)) return 1;   // small databases: not worth the thread start-up
    t = std::min(std::max(t, 1u), 32u);
    while (t & (t - 1)) t &= t - 1;   // a power of two: ownership below is a mask test
    return t;
}

template <class F>
void parallel(unsigned T, F f)
{
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(f, t);
    f(0u);
    for (auto &x : th) x.join();
}

}  // namespace

// The order-free case of normalise(), by clause chunk. Without unit clauses in the stream nothing a clause goes through
// depends on the clauses before it: no literal is ever false at level 0, a clause is kept iff it is no tautology, its
// literals stay in first-occurrence order, and a variable ends up mapped iff some non-tautological clause mentions it
// (the serial pass un-maps exactly the variables first seen in a tautology, and any later kept clause maps them again).
// Every thread normalises one contiguous chunk of the stream into its own buffer; the buffers are concatenated in chunk
// order, so `cls` is byte-identical to the serial result. The first unit clause (or empty clause) anywhere abandons the
// attempt: the caller then runs the serial pass, whose "units so far" semantics is order-dependent.
bool HostDb::normalise_parallel()
{
    const size_t n = raw.size();
    const unsigned T = host_threads(n);
    if (T < 2 || n < 2) return false;
    const size_t nv = (size_t)max_var + 1;
    // chunk boundaries right after a terminating 0
    std::vector<size_t> lo(T + 1, n);
    lo[0] = 0;
    for (unsigned t = 1; t < T; ++t) {
        size_t p = std::max(lo[t - 1], n / T * t);
        while (p < n && (p == 0 || raw[p - 1] != 0)) ++p;
        lo[t] = p;
    }
    varinfo.assign(nv, 0);
    std::vector<std::vector<uint32_t>> out(T);
    std::vector<uint64_t> kept(T, 0);
    std::vector<uint8_t> bail(T, 0);
    uint8_t *vi = varinfo.data();
    parallel(T, [&](unsigned t) {
        std::vector<uint32_t> &o = out[t];
        o.reserve((lo[t + 1] - lo[t]) + (lo[t + 1] - lo[t]) / 8 + 16);
        std::vector<uint32_t> table;   // long clauses: open-addressed set of the literals seen
        size_t s = lo[t];
        const size_t end = lo[t + 1];
        while (s < end) {
            size_t e = s;
            while (e < end && raw[e] != 0) ++e;
            if (e == s || e == end) { bail[t] = 1; return; }   // empty clause (contradictory) / unterminated tail: the serial pass decides
            const size_t len = e - s, hdr_pos = o.size();
            o.push_back(0);
            bool taut = false;
            if (len <= 16) {
                for (size_t i = s; i < e && !taut; ++i) {
                    const int32_t l = raw[i];
                    const uint32_t il = 2u * (uint32_t)(l < 0 ? -l : l) + (l < 0);
                    bool dup = false;
                    for (size_t k = hdr_pos + 1; k < o.size(); ++k) {
                        if (o[k] == il) { dup = true; break; }
                        if (o[k] == (il ^ 1u)) { taut = true; break; }
                    }
                    if (!dup && !taut) o.push_back(il);
                }
            } else {
                size_t cap = 32;
                while (cap < 2 * len) cap <<= 1;
                table.assign(cap, 0);
                for (size_t i = s; i < e && !taut; ++i) {
                    const int32_t l = raw[i];
                    const uint32_t il = 2u * (uint32_t)(l < 0 ? -l : l) + (l < 0);
                    bool dup = false;
                    // one table for both polarities, keyed by variable: the entry holds the literal seen
                    size_t h = ((il >> 1) * 0x9E3779B1u) & (cap - 1);
                    for (;; h = (h + 1) & (cap - 1)) {
                        const uint32_t x = table[h];
                        if (x == 0) { table[h] = il; break; }
                        if (x == il) { dup = true; break; }
                        if (x == (il ^ 1u)) { taut = true; break; }
                    }
                    if (!dup && !taut) o.push_back(il);
                }
            }
            s = e + 1;
            if (taut) { o.resize(hdr_pos); continue; }
            const size_t cnt = o.size() - hdr_pos - 1;
            if (cnt < 2) { bail[t] = 1; return; }   // a unit clause: later clauses depend on it
            o[hdr_pos] = (uint32_t)cnt;
            ++kept[t];
            for (size_t k = hdr_pos + 1; k < o.size(); ++k) {
                uint8_t *b = vi + (o[k] >> 1);
                if (!(__atomic_load_n(b, __ATOMIC_RELAXED) & 4)) __atomic_fetch_or(b, (uint8_t)4, __ATOMIC_RELAXED);
            }
        }
    });
    for (unsigned t = 0; t < T; ++t) if (bail[t]) return false;
    std::vector<size_t> at(T + 1, 0);
    for (unsigned t = 0; t < T; ++t) at[t + 1] = at[t] + out[t].size();
    cls.resize(at[T]);
    parallel(T, [&](unsigned t) { if (!out[t].empty()) std::memcpy(cls.data() + at[t], out[t].data(), out[t].size() * 4); });
    units.clear();
    ncls = 0;
    for (unsigned t = 0; t < T; ++t) ncls += kept[t];
    contradictory = false;
    return true;
}

// Flattening, multi-threaded by OWNER-COMPUTES: thread t owns the literals l with (l / 256) % T == t (blocks of 256
// literals dealt round-robin, so that the hot low-numbered variables of a power-law instance are spread). Every thread
// streams the whole clause list (sequential reads, cheap) and touches only the counters, cursors and rows of its own
// literals (the random accesses, a T-th of them each, with no atomics -- locked increments were measured to be slower
// than one thread, they serialise the cache misses). Entries land in every row in clause order, exactly as with one
// thread, so the database does not depend on the thread count. Only the arena is filled by clause chunk.
std::string HostDb::build_rows()
{
    const size_t nl = 2 * ((size_t)max_var + 1);
    hdr.assign(nl, RowHdr{0, 0, 0, 0});
    n_bin = n_tern = n_long = n_long_lits = 0;
    const unsigned T = host_threads(cls.size());
    const uint32_t own_mask = T - 1;
    auto owns = [own_mask](unsigned t, uint32_t l) { return ((l >> 8) & own_mask) == t; };

    // pass 1: occurrences per literal (owner-computes) + the global clause counts (thread 0)
    std::vector<uint32_t> bcount(nl + 1, 0);
    uint64_t arena_words = 4;  // unit 0 is never a valid clause reference
    parallel(T, [&](unsigned t) {
        for (size_t r = 0; r < cls.size();) {
            const uint32_t len = cls[r];
            const uint32_t *c = &cls[r + 1];
            if (len == 2) {
                if (owns(t, c[0])) ++bcount[c[0]];
                if (owns(t, c[1])) ++bcount[c[1]];
            } else if (len == 3) {
                for (int i = 0; i < 3; ++i) if (owns(t, c[i])) ++hdr[c[i]].ntern;
                if (t == 0) ++n_tern;
            } else {
                for (uint32_t i = 0; i < len; ++i) if (owns(t, c[i])) ++hdr[c[i]].nlong;
                if (t == 0) { ++n_long; n_long_lits += len; arena_words += (1 + (uint64_t)len + 3) & ~3ull; }
            }
            r += 1 + len;
        }
    });
    if (arena_words / 4 >= 0xffffffffull) return "clause arena exceeds the 32-bit clause index";

    // binary rows: gather, sort, unique -- a duplicate binary clause is kept once
    // (TopiConflictAnalysis.cc:25-41 with the default /bcp/existing_bin_wl_start = 1)
    std::vector<uint64_t> boff(nl + 1, 0);
    for (size_t l = 0; l < nl; ++l) boff[l + 1] = boff[l] + bcount[l];
    std::vector<uint32_t> bins(boff[nl]);
    std::vector<uint64_t> c_bins(T, 0);
    std::vector<uint64_t> pos(boff.begin(), boff.end() - 1);   // shared, every entry written by its owner only
    parallel(T, [&](unsigned t) {
        for (size_t r = 0; r < cls.size();) {
            const uint32_t len = cls[r];
            if (len == 2) {
                const uint32_t a = cls[r + 1], b = cls[r + 2];
                if (owns(t, a)) bins[pos[a]++] = b;
                if (owns(t, b)) bins[pos[b]++] = a;
            }
            r += 1 + len;
        }
        uint64_t tot = 0;
        for (size_t l = 0; l < nl; ++l) {
            if (!owns(t, (uint32_t)l)) { l |= 255; continue; }   // skip the rest of a block that is not mine
            uint32_t *b = bins.data() + boff[l], *e = bins.data() + boff[l + 1];
            if (e - b > 1) {
                std::sort(b, e);
                e = std::unique(b, e);
            }
            hdr[l].nbin = (uint32_t)(e - b);
            tot += hdr[l].nbin;
        }
        c_bins[t] = tot;
    });
    uint64_t total_bins = 0;
    for (unsigned t = 0; t < T; ++t) total_bins += c_bins[t];
    n_bin = total_bins / 2;

    // row extents
    uint64_t nslots = 0;
    for (size_t l = 0; l < nl; ++l) {
        if (nslots > 0xffffffffull) return "rows exceed the 32-bit slot index";
        hdr[l].start = (uint32_t)nslots;
        nslots += (uint64_t)((hdr[l].nbin + 1) / 2) + hdr[l].ntern + hdr[l].nlong;
    }
    if (nslots > 0xffffffffull) return "rows exceed the 32-bit slot index";
    slots.assign(nslots, Slot{0, 0});
    arena.assign(arena_words, 0);

    // clause references: the arena position of every long clause follows from the clause order alone
    // fill (owner-computes): binary part, then the ternary and long slots of the owned rows; thread 0 also copies the
    // long clauses into the arena
    std::vector<uint32_t> tpos(nl), lpos(nl);   // shared, every entry written by its owner only
    parallel(T, [&](unsigned t) {
        for (size_t l = 0; l < nl; ++l) {
            if (!owns(t, (uint32_t)l)) { l |= 255; continue; }
            const uint32_t *b = bins.data() + boff[l];
            Slot *row = slots.data() + hdr[l].start;
            for (uint32_t i = 0; i < hdr[l].nbin; ++i) {
                if (i & 1) row[i / 2].y = b[i]; else row[i / 2].x = b[i];
            }
            tpos[l] = hdr[l].start + (hdr[l].nbin + 1) / 2;
            lpos[l] = tpos[l] + hdr[l].ntern;
        }
        uint64_t aw = 4;
        for (size_t r = 0; r < cls.size();) {
            const uint32_t len = cls[r];
            const uint32_t *c = &cls[r + 1];
            if (len == 3) {
                if (owns(t, c[0])) slots[tpos[c[0]]++] = Slot{c[1], c[2]};
                if (owns(t, c[1])) slots[tpos[c[1]]++] = Slot{c[0], c[2]};
                if (owns(t, c[2])) slots[tpos[c[2]]++] = Slot{c[0], c[1]};
            } else if (len > 3) {
                const uint32_t cref = (uint32_t)(aw / 4);
                if (t == 0) {
                    arena[aw] = len;
                    std::memcpy(&arena[aw + 1], c, (size_t)len * 4);
                }
                for (uint32_t i = 0; i < len; ++i)
                    if (owns(t, c[i])) slots[lpos[c[i]]++] = Slot{c[(i + 1) % len], cref};
                aw += (1 + (uint64_t)len + 3) & ~3ull;
            }
            r += 1 + len;
        }
    });
    return "";
}

void HostDb::build_order()
{
    order.clear();
    max_height = 0;
    if (n_bin == 0 || hdr.size() < 4) return;
    const size_t nl = hdr.size();
    // successors of p = the binary partners in the row scanned when p becomes true (the row of not-p)
    auto nsucc = [&](uint32_t p) { return hdr[p ^ 1u].nbin; };
    auto succ = [&](uint32_t p, uint32_t j) {
        const Slot &s = slots[hdr[p ^ 1u].start + j / 2];
        return (j & 1) ? s.y : s.x;
    };
    std::vector<uint32_t> index(nl, 0), low(nl, 0), it(nl, 0), height(nl, 0);
    std::vector<uint8_t> onstack(nl, 0);
    std::vector<uint32_t> stack, call;
    uint32_t counter = 0;
    for (uint32_t root = 2; root < nl; ++root) {
        if (index[root] || nsucc(root) == 0) continue;   // literals without successors are complete at height 0
        index[root] = low[root] = ++counter;
        stack.push_back(root);
        onstack[root] = 1;
        call.push_back(root);
        while (!call.empty()) {
            const uint32_t p = call.back();
            if (it[p] < nsucc(p)) {
                const uint32_t q = succ(p, it[p]++);
                if (nsucc(q) == 0) continue;
                if (!index[q]) {
                    index[q] = low[q] = ++counter;
                    stack.push_back(q);
                    onstack[q] = 1;
                    call.push_back(q);
                } else if (onstack[q]) low[p] = std::min(low[p], index[q]);
                continue;
            }
            call.pop_back();
            if (!call.empty()) low[call.back()] = std::min(low[call.back()], low[p]);
            if (low[p] != index[p]) continue;
            // p closes a component: its members are the stack entries from p upwards; every successor outside of it
            // is complete, so the component's height is known now
            size_t base = stack.size();
            while (stack[--base] != p) {}
            uint32_t h = 0;
            for (size_t k = base; k < stack.size(); ++k) {
                const uint32_t m = stack[k];
                for (uint32_t j = 0, n = nsucc(m); j < n; ++j) {
                    const uint32_t q = succ(m, j);
                    if (!onstack[q]) h = std::max(h, height[q] + 1);
                }
            }
            for (size_t k = base; k < stack.size(); ++k) { height[stack[k]] = h; onstack[stack[k]] = 0; }
            stack.resize(base);
            max_height = std::max(max_height, h);
        }
    }
    std::vector<uint64_t> start((size_t)max_height + 2, 0);
    for (size_t l = 2; l < nl; ++l) ++start[height[l] + 1];
    for (size_t h = 0; h <= max_height; ++h) start[h + 1] += start[h];
    order.resize(nl - 2);
    for (size_t l = 2; l < nl; ++l) order[start[height[l]]++] = (uint32_t)l;
}

uint64_t HostDb::n_mapped() const
{
    uint64_t n = 0;
    for (uint8_t v : varinfo) n += (v >> 2) & 1;
    return n;
}

uint64_t HostDb::n_level0() const
{
    uint64_t n = 0;
    for (uint8_t v : varinfo) n += (v & 3) != 0;
    return n;
}

}  // namespace bbcp
