"""CPU tests of the host side: the C-ABI library loads and exports what include/bbcp.h declares; the
clause normaliser and row builder agree with an independent Python model; the engine refuses to run
without a CUDA device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from intel_sat_solver_b200 import BatchedTopor, BbcpError, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bbcp.h")).read()
    declared = sorted(set(re.findall(r"\b(bbcp_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == _lib.exported_names()


def py_normalise(words):
    """Independent model of AddUserClause's observable result (Topi.cc:377-653)."""
    unit_val, mapped, kept, contradictory = {}, set(), [], False
    cur = []
    for w in words.tolist():
        if w != 0:
            cur.append(w)
            continue
        clause, cur = cur, []
        if contradictory:
            continue
        if not clause:
            contradictory = True
            continue
        out, taut = [], False
        fresh = [abs(l) for l in clause if abs(l) not in mapped]
        seen_until = 0
        for l in clause:
            seen_until += 1
            if unit_val.get(abs(l)) == -l // abs(l) * 1 and False:
                pass
            if abs(l) in unit_val and unit_val[abs(l)] != (l > 0):
                continue
            if l in out:
                continue
            if -l in out:
                taut = True
                break
            out.append(l)
        touched = {abs(l) for l in clause[:seen_until]}
        if taut:
            continue
        mapped |= touched
        if not out:
            contradictory = True
        elif len(out) == 1:
            if abs(out[0]) not in unit_val:
                unit_val[abs(out[0])] = out[0] > 0
        else:
            kept.append(out)
    return kept, mapped, unit_val, contradictory


def ilit(l):
    return 2 * abs(l) + (l < 0)


def check_rows(t, kept, max_var):
    occ = {}
    for c in kept:
        for i, l in enumerate(c):
            occ.setdefault(ilit(l), []).append((c, i))
    for il in range(2, 2 * max_var + 2):
        row = t.HostRow(il).tolist()
        nbin, ntern, nlong = row[:3]
        p = 3
        bins = row[p:p + nbin]; p += nbin
        terns = [tuple(row[p + 2 * i:p + 2 * i + 2]) for i in range(ntern)]; p += 2 * ntern
        longs = []
        for _ in range(nlong):
            blocker, size = row[p], row[p + 1]
            longs.append((blocker, tuple(row[p + 2:p + 2 + size])))
            p += 2 + size
        assert p == len(row)
        exp_bins = sorted({ilit(c[1 - i]) for c, i in occ.get(il, []) if len(c) == 2})
        assert bins == exp_bins, il
        exp_terns = sorted(tuple(sorted(ilit(x) for k, x in enumerate(c) if k != i)) for c, i in occ.get(il, []) if len(c) == 3)
        assert sorted(tuple(sorted(x)) for x in terns) == exp_terns, il
        exp_longs = sorted((ilit(c[(i + 1) % len(c)]), tuple(ilit(x) for x in c)) for c, i in occ.get(il, []) if len(c) > 3)
        assert sorted(longs) == exp_longs, il


@pytest.mark.parametrize("name", ["regr_20", "regr_47", "regr_66", "hand_taut_dup", "hand_long", "hand_units", "hand_alias", "fuzz_3"])
def test_normaliser_and_rows_match_python_model(golden, name):
    words = golden[f"{name}/words"]
    kept, mapped, unit_val, contradictory = py_normalise(words)
    t = BatchedTopor()
    t.AddClauses(words)
    rc = t.HostPrepare()
    assert (rc == 1) == contradictory
    if contradictory:
        return
    info = t.DbInfo()
    max_var = int(np.abs(words).max())
    assert info["max_var"] == max_var
    assert info["n_mapped_vars"] == len(mapped)
    assert info["n_level0"] == len(unit_val)          # before level-0 propagation: the direct units only
    assert info["n_tern"] == sum(len(c) == 3 for c in kept)
    assert info["n_long"] == sum(len(c) > 3 for c in kept)
    assert info["n_bin"] == len({frozenset(c) for c in kept if len(c) == 2})
    check_rows(t, kept, max_var)


@pytest.mark.parametrize("name,expect", [("hand_empty", 1), ("hand_unsat_units", 1), ("hand_units", 0)])
def test_intake_contradictions(golden, name, expect):
    t = BatchedTopor()
    t.AddClauses(golden[f"{name}/words"])
    assert t.HostPrepare() == expect


def test_add_and_add_clause_agree_with_bulk():
    words = np.array([1, -2, 3, 0, 2, 2, -4, 0, 5, 0, -5, 6, 7, 8, 0], dtype=np.int32)
    a, b, c = BatchedTopor(), BatchedTopor(), BatchedTopor()
    a.AddClauses(words)
    for w in words.tolist():
        b.lib.bbcp_add(b.handle, w)
    cur = []
    for w in words.tolist():
        cur.append(w)
        if w == 0:
            c.AddClause(cur + [9, 9])  # reading stops at the first 0 (Topor.hpp:33-38)
            cur = []
    for t in (a, b, c):
        assert t.HostPrepare() == 0
    rows = [[t.HostRow(i).tolist() for i in range(2, 18)] for t in (a, b, c)]
    assert rows[0] == rows[1] == rows[2]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    t = BatchedTopor()
    t.AddClause([1, 2, 0])
    with pytest.raises(BbcpError):
        t.Finalize()
    assert t.IsError() and "CUDA" in t.GetStatusExplanation()
    with pytest.raises(BbcpError):
        t.ProbeAll()


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "intel_sat_solver_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "oracle" not in src.lower(), fn


def test_normaliser_on_seeded_random_clause_streams():
    """300 small random clause streams with duplicate literals, tautologies, unit clauses, repeated clauses and the
    occasional empty clause: the host database (intake result-equivalent to AddUserClause, Topi.cc:377-653) agrees
    with the independent Python model on status, counts and every row."""
    rng = np.random.default_rng(20260101)
    n_contra = n_rows = 0
    for trial in range(300):
        nv = int(rng.integers(3, 12))
        words = []
        for _ in range(int(rng.integers(1, 25))):
            k = int(rng.choice([1, 2, 2, 2, 3, 3, 3, 4, 5, 6])) if rng.random() < 0.995 else 0
            lits = [int(rng.integers(1, nv + 1)) * (1 if rng.random() < 0.5 else -1) for _ in range(k)]
            if lits and rng.random() < 0.15:
                lits.append(lits[0])            # duplicate literal
            if lits and rng.random() < 0.08:
                lits.append(-lits[-1])          # tautology
            words += lits + [0]
        words = np.asarray(words, dtype=np.int32)
        kept, mapped, unit_val, contradictory = py_normalise(words)
        t = BatchedTopor()
        t.AddClauses(words)
        rc = t.HostPrepare()
        assert (rc == 1) == contradictory, (trial, words.tolist())
        if contradictory:
            n_contra += 1
            continue
        info = t.DbInfo()
        max_var = int(np.abs(words).max()) if words.size and np.abs(words).max() > 0 else 0
        assert info["n_mapped_vars"] == len(mapped) and info["n_level0"] == len(unit_val), (trial, words.tolist())
        assert info["n_tern"] == sum(len(c) == 3 for c in kept) and info["n_long"] == sum(len(c) > 3 for c in kept)
        assert info["n_bin"] == len({frozenset(c) for c in kept if len(c) == 2})
        if max_var:
            check_rows(t, kept, max_var)
            n_rows += 1
    assert n_contra > 10 and n_rows > 100


def test_row_build_does_not_depend_on_the_thread_count(golden):
    """The multi-threaded flattening (owner-computes by literal block) produces the same database, entry for entry
    and in the same order, as one thread."""
    import subprocess, sys, json
    prog = (
        "import sys, json, numpy as np; sys.path.insert(0, %r); "
        "from intel_sat_solver_b200 import BatchedTopor; g = np.load(%r); out = {}\n"
        "for name in ['fuzz_10', 'fuzz_3', 'regr_66', 'hand_long']:\n"
        "    t = BatchedTopor(); t.AddClauses(g[name + '/words']); t.HostPrepare(); mv = t.max_var\n"
        "    out[name] = [t.HostRow(l).tolist() for l in range(2, 2 * mv + 2)]\n"
        "print(json.dumps(out))" % (ROOT, os.path.join(ROOT, "tests", "golden", "golden.npz")))
    rows = {}
    for th in ("1", "4", "8"):
        env = dict(os.environ, BBCP_HOST_THREADS=th)
        rows[th] = json.loads(subprocess.run([sys.executable, "-c", prog], check=True, capture_output=True, text=True, env=env).stdout)
    assert rows["1"] == rows["4"] == rows["8"]
    assert sum(len(r) for r in rows["1"]["fuzz_10"]) > 10000


def test_intake_by_chunk_equals_the_pass_in_stream_order():
    """The multi-threaded intake (clause chunks normalised independently; applicable when no clause is or becomes a
    unit) gives the same database as the defining pass in stream order: counts, mapped variables and every row, on
    streams with duplicate literals, tautologies (incl. variables that occur in nothing else), clauses beyond the
    quadratic-dedupe length. Streams with unit or empty clauses must fall back."""
    import subprocess, sys, json
    prog = r"""
import sys, json, os, numpy as np
sys.path.insert(0, %r)
from intel_sat_solver_b200 import BatchedTopor
rng = np.random.default_rng(77)
out = []
for trial in range(40):
    nv = int(rng.integers(20, 400))
    words = []
    for _ in range(int(rng.integers(50, 600))):
        k = int(rng.choice([2, 2, 3, 3, 3, 4, 6, 20, 40]))
        lits = [int(rng.integers(1, nv + 1)) * (1 if rng.random() < 0.5 else -1) for _ in range(k)]
        if rng.random() < 0.2: lits.append(lits[0])
        if rng.random() < 0.1: lits.insert(int(rng.integers(0, len(lits))), -lits[-1])
        if rng.random() < 0.05: lits = [nv + 1 + len(words) %% 7, -(nv + 1 + len(words) %% 7)] + lits[:1]   # tautology on an otherwise unused variable
        words += lits + [0]
    kind = trial %% 4
    if kind == 2: words[len(words) // 2:len(words) // 2] = [5, 0]   # a unit clause in the middle
    if kind == 3 and trial %% 8 == 3: words[10:10] = [0]    # an empty clause (or a split one)
    t = BatchedTopor(); t.AddClauses(np.asarray(words, dtype=np.int32))
    rc = t.HostPrepare()
    rec = {"rc": rc, "info": {k: v for k, v in t.DbInfo().items() if k != "device_bytes"}}
    if rc == 0:
        rec["rows"] = [t.HostRow(l).tolist() for l in range(2, 2 * t.max_var + 2)]
    out.append(rec)
print(json.dumps(out))
""" % ROOT
    res, traces = {}, {}
    for th in ("1", "2", "8"):
        env = dict(os.environ, BBCP_HOST_THREADS=th, BBCP_HOST_TRACE="1")
        r = subprocess.run([sys.executable, "-c", prog], check=True, capture_output=True, text=True, env=env)
        res[th] = json.loads(r.stdout)
        traces[th] = r.stderr
    assert res["1"] == res["2"] == res["8"]
    assert "by chunk" not in traces["1"]
    assert traces["8"].count("by chunk") >= 15 and traces["8"].count("stream order") >= 10   # both paths were exercised
    assert sum(r["rc"] == 0 for r in res["1"]) >= 30


def _scc_ids(nl, succ):
    """Tarjan, recursive-free reference (small inputs): component id per literal."""
    import sys
    index, low, comp, stack, on = {}, {}, {}, [], set()
    counter = [0]
    for root in range(2, nl):
        if root in index:
            continue
        work = [(root, 0)]
        index[root] = low[root] = counter[0]; counter[0] += 1; stack.append(root); on.add(root)
        while work:
            v, i = work.pop()
            if i < len(succ[v]):
                work.append((v, i + 1))
                w = succ[v][i]
                if w not in index:
                    index[w] = low[w] = counter[0]; counter[0] += 1; stack.append(w); on.add(w)
                    work.append((w, 0))
                elif w in on:
                    low[v] = min(low[v], index[w])
            else:
                if work:
                    low[work[-1][0]] = min(low[work[-1][0]], low[v])
                if low[v] == index[v]:
                    while True:
                        w = stack.pop(); on.discard(w); comp[w] = v
                        if w == v:
                            break
    return comp


@pytest.mark.parametrize("name", ["regr_20", "regr_47", "regr_66", "regr_72", "fuzz_3", "fuzz_10", "hand_long"])
def test_probe_schedule_is_children_first(golden, name):
    """bbcp_host_order: a permutation of all internal literals in which q precedes p whenever p -> q is a binary
    implication and the two are not in one strongly connected component (equivalent literals share a height)."""
    t = BatchedTopor()
    t.AddClauses(golden[f"{name}/words"])
    if t.HostPrepare() != 0:
        pytest.skip("contradictory at intake")
    nl = 2 * t.max_var + 2
    order = t.HostOrder().tolist()
    succ = {l: [] for l in range(nl)}
    nbin = 0
    for l in range(2, nl):
        row = t.HostRow(l ^ 1).tolist()      # the row walked when l becomes true
        succ[l] = row[3:3 + row[0]]
        nbin += row[0]
    if nbin == 0:
        assert order == []
        return
    assert sorted(order) == list(range(2, nl))
    pos = {l: i for i, l in enumerate(order)}
    comp = _scc_ids(nl, succ)
    edges = cross = 0
    for p in range(2, nl):
        for q in succ[p]:
            edges += 1
            if comp[p] != comp[q]:
                cross += 1
                assert pos[q] < pos[p], (p, q)
    assert edges == nbin and cross > 0


def test_literals_beyond_the_engine_range_are_refused():
    """Internal literals are 2*var+neg in 30 bits (the two bits above are the trail-entry mode): |literal| >= 2^29 is refused
    with BBCP_ERR_INDEX_TOO_NARROW (the reference's STATUS_INDEX_TOO_NARROW), sticky, and nothing is allocated for it."""
    import ctypes as C
    lib = _lib.load()
    for bad in (1 << 29, -(1 << 29), (1 << 31) - 1, -(1 << 31)):
        h = lib.bbcp_create(0)
        a = (C.c_int32 * 2)(1, 2)
        assert lib.bbcp_add_clause(h, a, 2) == 0
        assert lib.bbcp_add(h, bad) == -3
        assert lib.bbcp_status(h) == -3 and b"2^29" in lib.bbcp_error_string(h)
        b = (C.c_int32 * 3)(3, bad, 0)
        assert lib.bbcp_add_clause(h, b, 3) == -3 and lib.bbcp_add_clauses(h, b, 3) == -3
        lib.bbcp_release(h)
    h = lib.bbcp_create(0)
    ok = (C.c_int32 * 3)((1 << 29) - 1, -5, 0)
    assert lib.bbcp_add_clauses(h, ok, 3) == 0 and lib.bbcp_status(h) == 0
    lib.bbcp_release(h)
