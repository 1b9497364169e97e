"""GPU parity tests (run with -m gpu on the B200 box). Everything goes through the C ABI (ctypes ->
libbbcp.so -> CUDA kernels) and is compared bit-exactly with
  * the committed golden vectors produced by the reference's own CTopi::BCP() (tests/golden/), and
  * the checker binary run live on seeded synthetic instances of the BASELINE.json shapes
    (oracle/_ref/ref_bcp = the reference itself when it travelled with the snapshot, else the C port).
"""
import os
import tempfile

import numpy as np
import pytest

from intel_sat_solver_b200 import BatchedTopor, TToporLitVal, TToporReturnVal, fileio
import gpu_util

pytestmark = pytest.mark.gpu


def engine_for(words):
    t = BatchedTopor()
    t.AddClauses(words)
    rc = t.Finalize()
    return t, rc


def test_all_goldens_probe_and_cube_parity(golden, golden_names):
    n_probe = n_failed = n_cube = 0
    for name in golden_names:
        g = lambda k: golden[f"{name}/{k}"]
        t, rc = engine_for(g("words"))
        assert rc == int(g("status")[0]), name
        if rc:
            # sticky contradictory status: batches are refused with the same code (Topi.cc:182, 449, 1015)
            assert t.lib.bbcp_probe_all(t.handle, None, None) == 1 and t.Solve([1]) == TToporReturnVal.RET_UNSAT
            continue
        assert np.array_equal(t.Level0(), g("level0")), name
        res = t.ProbeAll()
        gpu_util.assert_same_results(res, g("probe_flag"), g("probe_off"), g("probe_lits"), name + " probes")
        # failed-literal set == the probes flagged 1
        idx = np.nonzero(g("probe_flag") == 1)[0]
        expect_failed = np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)
        assert sorted(t.FailedLiterals().tolist(), key=lambda l: (abs(l), l < 0)) == sorted(expect_failed.tolist(), key=lambda l: (abs(l), l < 0)), name
        failed, unit = t.Bitmaps()
        fb = np.unpackbits(failed.view(np.uint8), bitorder="little")
        ub = np.unpackbits(unit.view(np.uint8), bitorder="little")
        il = np.nonzero(fb)[0]
        assert np.array_equal(np.nonzero(ub)[0], np.sort(il ^ 1)), name
        res = t.PropagateBatch(g("cube_lits"), g("cube_off"))
        gpu_util.assert_same_results(res, g("cube_flag"), g("cube_impl_off"), g("cube_impl_lits"), name + " cubes")
        assert res.info["n_ok"] + res.info["n_conflict"] + res.info["n_skipped"] == res.info["n"], name
        n_probe += int((g("probe_flag") <= 1).sum()); n_failed += idx.size; n_cube += g("cube_flag").size
    assert n_probe > 14000 and n_failed >= 77 and n_cube > 5000


def test_regression_corpus_totals(golden, golden_names):
    """4,694 valid probes and 77 failed literals over regr_1..72."""
    valid = failed = 0
    for name in golden_names:
        if not name.startswith("regr_"):
            continue
        t, rc = engine_for(golden[f"{name}/words"])
        assert rc == 0
        res = t.ProbeAll(implied=False)
        valid += int((res.flag <= 1).sum())
        failed += int((res.flag == 1).sum())
        assert res.info["n_conflict"] == int((res.flag == 1).sum())
    assert (valid, failed) == (4694, 77)


@pytest.mark.parametrize("kw", [dict(small_trail=32), dict(small_trail=32, mid_trail=32), dict(no_len_prune=True), dict(out_capacity=64),
                                dict(small_trail=32, mid_trail=64, out_capacity=16), dict(no_reuse=True), dict(no_reuse=True, small_trail=32, mid_trail=32),
                                dict(no_len_prune=True, small_trail=32)])
def test_option_variants_keep_results(golden, kw):
    """Tiny warp-private state (forces the CTA-per-probe shared-memory tier), tiny CTA state too (forces the
    global-slab tier), no length pruning, and an implied-literal buffer that overflows and regrows must all
    give the same bits."""
    for name in ["regr_47", "regr_72", "fuzz_5", "fuzz_10", "hand_long", "hand_alias"]:
        g = lambda k: golden[f"{name}/{k}"]
        t, rc = engine_for(g("words"))
        res = t.ProbeAll(**kw)
        gpu_util.assert_same_results(res, g("probe_flag"), g("probe_off"), g("probe_lits"), f"{name} {kw}")
        res = t.PropagateBatch(g("cube_lits"), g("cube_off"), **kw)
        gpu_util.assert_same_results(res, g("cube_flag"), g("cube_impl_off"), g("cube_impl_lits"), f"{name} cubes {kw}")


SHAPES = [
    ("3sat", dict(n=20000, m=84000, seed=1), None),
    ("aig", dict(frames=4, gates=3000, inputs=100, latches=100, window=300, seed=2), None),
    ("comm", dict(n=20000, m=80000, comms=20, seed=3), dict(ncubes=2048, split=8, extra=12)),
    ("plaw", dict(n=20000, m=80000, seed=5), None),
]


@pytest.mark.parametrize("kind,kw,cubekw", SHAPES, ids=[s[0] for s in SHAPES])
def test_synthetic_shapes_against_checker(kind, kw, cubekw):
    _synthetic_shape(kind, kw, cubekw, 0)


def _synthetic_shape(kind, kw, cubekw, width):
    with tempfile.TemporaryDirectory() as d:
        gkw = dict(kw)
        if cubekw:
            gkw.update(cubekw, cubes=os.path.join(d, "c.cubes"))
        inst, _ = gpu_util.gen(kind, d, **gkw)
        words = fileio.read_bcnf(inst)
        t, rc = engine_for(words)
        ref, _ = gpu_util.run_checker(inst)
        assert rc == ref.status
        res = t.ProbeAll()
        gpu_util.assert_same_results(res, ref.flag, ref.off, ref.lits, kind)
        # the reference counts a propagation only for literals whose negation ever had a row; on a
        # non-conflicting probe that is order-independent up to rows allocated by watch moves, so only
        # the looser invariant is asserted
        assert res.info["props_all"] >= res.info["props_ref"] > 0
        if cubekw:
            import numpy as np
            buf = np.fromfile(gkw["cubes"], dtype=np.uint64, count=3)
            nc, nl = int(buf[1]), int(buf[2])
            off = np.fromfile(gkw["cubes"], dtype=np.uint64, offset=24, count=nc + 1)
            lits = np.fromfile(gkw["cubes"], dtype=np.int32, offset=24 + 8 * (nc + 1), count=nl)
            cref, _ = gpu_util.run_checker(inst, gkw["cubes"])
            res = t.PropagateBatch(lits, off)
            gpu_util.assert_same_results(res, cref.flag, cref.off, cref.lits, kind + " cubes")
        # the same instance with tiny state budgets: deep probes cascade to the CTA-per-probe tiers
        res2 = t.ProbeAll(small_trail=32)
        gpu_util.assert_same_results(res2, ref.flag, ref.off, ref.lits, kind + " tier2")
        res3 = t.ProbeAll(small_trail=32, mid_trail=32)
        gpu_util.assert_same_results(res3, ref.flag, ref.off, ref.lits, kind + " tier3")
        assert res2.info["n_mid"] >= res.info["n_mid"] and res3.info["n_large"] >= res2.info["n_large"]
        if kind in ("aig", "plaw"):
            assert res2.info["n_mid"] > 0 and res3.info["n_large"] > 0
        # closure reuse (finished probes' results adopted by later probes) on and off: same bits
        res4 = t.ProbeAll(no_reuse=True)
        gpu_util.assert_same_results(res4, ref.flag, ref.off, ref.lits, kind + " no reuse")
        assert res4.info["adopted"] == 0 and res4.info["inherited"] == 0
        if kind == "aig":
            assert res.info["adopted"] > 0 and res.info["adopt_lits"] > res.info["adopted"]
        if kind == "plaw":
            assert res.info["inherited"] > 0


def test_probe_range_shards_compose(golden):
    g = lambda k: golden[f"fuzz_10/{k}"]
    t, _ = engine_for(g("words"))
    n = 2 * t.max_var
    cuts = [0, n // 3, n // 3 + 1, n // 2, n]
    flags, sizes, lits = [], [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        r = t.ProbeRange(a, b)
        flags.append(r.flag); sizes.append(np.diff(r.impl_off)); lits.append(r.impl_lits)
    assert np.array_equal(np.concatenate(flags), g("probe_flag"))
    assert np.array_equal(np.concatenate(sizes), np.diff(g("probe_off")))
    assert np.array_equal(np.concatenate(lits), g("probe_lits"))


def test_topor_style_single_cube_api(golden):
    """AddClause / Solve(assumps) / GetLitValue / GetLitDecLevel / Backtrack, checked against the golden
    cube results (what the reference's public API yields with a stop callback, SURVEY.md 8c route B)."""
    name = "regr_72"
    g = lambda k: golden[f"{name}/{k}"]
    words = g("words")
    t = BatchedTopor()
    cur = []
    for w in words.tolist():
        cur.append(w)
        if w == 0:
            t.AddClause(cur)
            cur = []
    level0 = set(g("level0").tolist())
    off, lits, flag = g("cube_off"), g("cube_lits"), g("cube_flag")
    ioff, ilits = g("cube_impl_off"), g("cube_impl_lits")
    maxv = int(np.abs(words).max())
    checked = 0
    for i in range(min(flag.size, 40)):
        cube = lits[int(off[i]):int(off[i + 1])].tolist()
        if flag[i] == 4:
            continue
        r = t.Solve(cube)
        if flag[i] in (1, 3):
            assert r == TToporReturnVal.RET_UNSAT
            continue
        assert r in (TToporReturnVal.RET_USER_INTERRUPT, TToporReturnVal.RET_SAT)
        implied = set(ilits[int(ioff[i]):int(ioff[i + 1])].tolist())
        for v in range(1, maxv + 1):
            val = t.GetLitValue(v)
            if v in implied or v in level0:
                assert val == TToporLitVal.VAL_SATISFIED
                assert t.GetLitDecLevel(v) == (0 if v in level0 else 1)
            elif -v in implied or -v in level0:
                assert val == TToporLitVal.VAL_UNSATISFIED
            else:
                assert val in (TToporLitVal.VAL_UNASSIGNED, TToporLitVal.VAL_DONT_CARE)
        t.Backtrack(0)
        assert t.GetLitValue(next(iter(implied))) != TToporLitVal.VAL_SATISFIED if implied - level0 and not (set(map(abs, implied)) & set(map(abs, level0))) else True
        checked += 1
    assert checked > 10 and t.GetPropagations() > 0 and not t.IsError()


def test_explicit_probe_list_equals_probe_universe(golden):
    """Depth-1 cubes are probes: same flags, same implied sets, same failed-literal bitmap as ProbeAll."""
    for name in ["regr_47", "fuzz_10", "fuzz_5"]:
        g = lambda k: golden[f"{name}/{k}"]
        t, _ = engine_for(g("words"))
        n = 2 * t.max_var
        idx = np.arange(n)
        lits = np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)
        off = np.arange(n + 1, dtype=np.uint64)
        res = t.PropagateBatch(lits, off)
        gpu_util.assert_same_results(res, g("probe_flag"), g("probe_off"), g("probe_lits"), name)
        fa, ua = t.Bitmaps()
        t.ClearBitmaps()
        assert not t.Bitmaps()[0].any()
        t.ProbeAll(implied=False)
        fb, ub = t.Bitmaps()
        assert np.array_equal(fa, fb) and np.array_equal(ua, ub) and fa.any() == bool((g("probe_flag") == 1).any())


def test_probe_list_and_u32_offsets(golden):
    """bbcp_probe_lits (4-byte probe list) and BBCP_OPT_OFF32 (u32 CSR offsets) give the bits of the universe pass;
    a 0 in the list and out-of-range variables are reported unmapped."""
    for name in ["regr_47", "fuzz_10", "hand_long"]:
        g = lambda k: golden[f"{name}/{k}"]
        t, _ = engine_for(g("words"))
        n = 2 * t.max_var
        idx = np.arange(n)
        lits = np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)
        res = t.ProbeLits(lits)
        gpu_util.assert_same_results(res, g("probe_flag"), g("probe_off"), g("probe_lits"), name)
        res32 = t.ProbeLits(lits, off32=True)
        assert res32.impl_off.dtype == np.uint32
        gpu_util.assert_same_results(res32, g("probe_flag"), g("probe_off").astype(np.uint32), g("probe_lits"), name + " off32")
        res32 = t.ProbeAll(off32=True)
        gpu_util.assert_same_results(res32, g("probe_flag"), g("probe_off").astype(np.uint32), g("probe_lits"), name + " universe off32")
        # fetching with the wrong width is refused, not silently mis-read
        assert t.lib.bbcp_fetch_implied(t.handle, None, None) == -4
        # permuted list with junk entries
        rng = np.random.default_rng(7)
        perm = rng.permutation(n)
        junk = np.concatenate([lits[perm], np.array([0, t.max_var + 5, -(t.max_var + 9)], dtype=np.int32)])
        r = t.ProbeLits(junk)
        assert np.array_equal(r.flag[:n], g("probe_flag")[perm]) and (r.flag[n:] == 4).all()
        sizes = np.diff(g("probe_off"))
        assert np.array_equal(np.diff(r.impl_off)[:n], sizes[perm])
        for k in range(0, n, max(1, n // 50)):
            p = perm[k]
            assert np.array_equal(r.implied(k), g("probe_lits")[int(g("probe_off")[p]):int(g("probe_off")[p + 1])])
        empty = t.ProbeLits(np.zeros(0, dtype=np.int32))
        assert empty.info["n"] == 0 and empty.impl_off.tolist() == [0]


def test_large_batch_scan_and_giant_sets():
    """Many tiles of the single-pass compaction (per-tile sums across > 100 tiles), sets copied by thread / warp /
    CTA / chunked, against the checker on a chain instance whose probes imply thousands of literals."""
    with tempfile.TemporaryDirectory() as d:
        # x1 -> x2 -> ... -> xN chain: probing x_i implies N-i+1 literals; plus noise ternaries
        N = 40000
        words = []
        for i in range(1, N):
            words += [-i, i + 1, 0]
        rng = np.random.default_rng(3)
        for _ in range(20000):
            a, b, c = rng.choice(np.arange(N + 1, N + 20001), 3, replace=False)
            words += [int(a), -int(b), int(c), 0]
        words = np.asarray(words, dtype=np.int32)
        used = set(np.abs(words).tolist())
        t, rc = engine_for(words)
        assert rc == 0
        lits = np.concatenate([np.arange(1, N + 1, 97), -np.arange(1, N + 1, 89), np.arange(N + 1, N + 20001)]).astype(np.int32)
        res = t.ProbeLits(lits)
        sizes = np.diff(res.impl_off).astype(np.int64)
        for k, l in enumerate(lits.tolist()):
            if 0 < l <= N:
                assert res.flag[k] == 0 and sizes[k] == N - l + 1
                if k % 40 == 0:
                    assert np.array_equal(res.implied(k), np.arange(l, N + 1, dtype=np.int32))
            elif l < 0:
                assert res.flag[k] == 0 and sizes[k] == -l
                if k % 40 == 0:
                    assert np.array_equal(res.implied(k), -np.arange(1, -l + 1, dtype=np.int32))
            elif l in used:
                assert res.flag[k] == 0 and sizes[k] == 1 and res.implied(k)[0] == l
            else:
                assert res.flag[k] == 4 and sizes[k] == 0
        assert res.info["n_large"] > 0 and res.info["n_mid"] > 0
        assert int(res.impl_off[-1]) == res.info["n_implied"] == int(sizes.sum())


def test_pipelined_host_call_matches_plain_call():
    """bbcp_probe_lits_to_host (chunked, upload / kernels / read-back overlapped, two result sets) against
    bbcp_probe_lits + fetch on an instance big enough for several chunks, and against the goldens on small ones."""
    with tempfile.TemporaryDirectory() as d:
        inst, _ = gpu_util.gen("comm", d, n=300000, m=1200000, comms=300, seed=11)
        t, rc = engine_for(fileio.read_bcnf(inst))
        assert rc == 0
        n = 2 * t.max_var
        rng = np.random.default_rng(5)
        idx = rng.permutation(n)
        lits = np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)
        ref = t.ProbeLits(lits)
        os.environ["BBCP_CHUNKS"] = "5"    # (the default keeps chunks above 512 K probes; force several here)
        for off32 in (False, True):
            res = t.ProbeLitsToHost(lits, impl_cap=int(ref.info["n_implied"]), off32=off32)
            assert res.info["n"] == n and res.info["launches"] >= 15     # five chunks
            assert np.array_equal(res.flag, ref.flag)
            assert np.array_equal(res.impl_off.astype(np.uint64), ref.impl_off)
            assert np.array_equal(res.impl_lits, ref.impl_lits)
        # a buffer that is too small is refused
        with pytest.raises(Exception):
            t.ProbeLitsToHost(lits, impl_cap=int(ref.info["n_implied"]) - 1)
        assert not t.IsError()
        del os.environ["BBCP_CHUNKS"]
    g = np.load(os.path.join(gpu_util.ROOT, "tests", "golden", "golden.npz"))
    for name in ["fuzz_10", "regr_47"]:
        t, _ = engine_for(g[f"{name}/words"])
        n = 2 * t.max_var
        idx = np.arange(n)
        lits = np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)
        res = t.ProbeLitsToHost(lits)
        gpu_util.assert_same_results(res, g[f"{name}/probe_flag"], g[f"{name}/probe_off"], g[f"{name}/probe_lits"], name)
        empty = t.ProbeLitsToHost(np.zeros(0, dtype=np.int32))
        assert empty.impl_off.tolist() == [0]


def test_pinned_caller_buffers_are_used_in_place():
    """bbcp_probe_lits_to_host with PINNED caller buffers: the triage kernel pulls the probe list out of the caller's memory
    and writes the flags it decides straight into the caller's flag buffer (BBCP_E2E=pull, the default; mirror: uploaded list,
    flags in place). Same bytes as the copy-engine path (BBCP_E2E=dma) and as the goldens: with and without queued probes,
    compact and standard format, aligned and misaligned views, lengths that are not multiples of the triage's vector width."""
    import torch
    g = np.load(os.path.join(gpu_util.ROOT, "tests", "golden", "golden.npz"))
    cases = []
    with tempfile.TemporaryDirectory() as d:
        # (a) random 3-SAT: every probe is decided by the triage (no binary clause): the flags never leave through a copy
        inst, _ = gpu_util.gen("3sat", d, n=40000, m=168000, seed=3)
        cases.append(("3sat", fileio.read_bcnf(inst)))
        # (b) community instance with binary clauses: queued probes, all tiers
        inst, _ = gpu_util.gen("comm", d, n=60000, m=240000, comms=60, seed=12)
        cases.append(("comm", fileio.read_bcnf(inst)))
    cases.append(("fuzz_10", g["fuzz_10/words"]))
    try:
        for name, words in cases:
            t, rc = engine_for(words)
            assert rc == 0
            n_all = 2 * t.max_var
            rng = np.random.default_rng(7)
            idx = rng.permutation(n_all)
            lits_all = np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)
            lits_all[::97] = 0                      # zeros name no variable: reported as unmapped
            for n, shift in ((n_all, 0), (n_all - 3, 0), (n_all - 5, 1), (min(n_all, 1001), 2)):
                lits = lits_all[:n]
                for compact in (False, True):
                    os.environ["BBCP_E2E"] = "dma"
                    ref = t.ProbeLitsToHost(lits, compact_self=compact, off32=True)
                    for mode, chunks in (("pull", None), ("pull", "3"), ("mirror", "2")):
                        os.environ["BBCP_E2E"] = mode
                        if chunks:
                            os.environ["BBCP_CHUNKS"] = chunks
                        # pinned buffers; `shift` moves the views off their 16-byte (list) and 8-byte (flags) alignment
                        p_lits = torch.empty(n + 8, dtype=torch.int32).pin_memory()
                        p_flag = torch.full((n + 16,), 0x55, dtype=torch.uint8).pin_memory()
                        p_off = torch.zeros(n + 1, dtype=torch.int32).pin_memory()
                        p_impl = torch.empty(max(int(ref.info["n_implied"]), 1), dtype=torch.int32).pin_memory()
                        v_lits = p_lits.numpy()[shift:shift + n]
                        v_lits[:] = lits
                        v_flag = p_flag.numpy()[shift:shift + n]
                        for rep in range(2):        # the second call sizes its chunks from the first one's result
                            res = t.ProbeLitsToHost(v_lits, flag_out=v_flag, off_out=p_off.numpy().view(np.uint32), impl_out=p_impl.numpy(),
                                                    compact_self=compact, off32=True)
                            assert np.array_equal(res.flag, ref.flag), (name, n, shift, compact, mode, rep)
                            if ref.info["n_implied"]:
                                assert np.array_equal(res.impl_off, ref.impl_off) and np.array_equal(res.impl_lits, ref.impl_lits)
                            assert res.info["n_implied"] == ref.info["n_implied"]
                        assert (p_flag.numpy()[:shift] == 0x55).all() and (p_flag.numpy()[shift + n:] == 0x55).all()   # nothing written past the view
                        os.environ.pop("BBCP_CHUNKS", None)
            if name == "3sat":
                assert ref.info["n_deep"] == 0
            else:
                assert ref.info["n_deep"] > 0
    finally:
        os.environ.pop("BBCP_E2E", None)
        os.environ.pop("BBCP_CHUNKS", None)
    # against the golden results of the reference
    n = 2 * t.max_var
    idx = np.arange(n)
    lits = torch.from_numpy(np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)).pin_memory()
    flag = torch.empty(n, dtype=torch.uint8).pin_memory()
    res = t.ProbeLitsToHost(lits.numpy(), flag_out=flag.numpy())
    gpu_util.assert_same_results(res, g["fuzz_10/probe_flag"], g["fuzz_10/probe_off"], g["fuzz_10/probe_lits"], "fuzz_10")


def test_probe_to_fixpoint_matches_checker_rounds(golden):
    """Unit-feedback rounds (survey 'next' row N2): probe all, add the negations of the failed literals as units,
    re-finalize, repeat. The same loop driven through the CPU checker gives the same number of rounds, the same
    number of units and the same final level-0 assignment (or the same refutation)."""
    from oracle_lib import Oracle
    seen_multi = False
    for name in ["fuzz_10", "fuzz_5", "fuzz_3", "fuzz_7", "regr_47", "regr_72", "regr_20"]:
        words = golden[f"{name}/words"]
        t, rc = engine_for(words)
        assert rc == 0
        got = t.ProbeToFixpoint()
        # checker loop
        w = words.copy()
        rounds = units = 0
        status = 0
        while True:
            o = Oracle(w)
            if o.status:
                status = 1
                break
            flag, _, _, _ = o.run()
            rounds += 1
            idx = np.nonzero(flag == 1)[0]
            if idx.size == 0:
                level0 = o.level0()
                break
            units += idx.size
            neg = np.where(idx & 1, (idx // 2 + 1), -(idx // 2 + 1)).astype(np.int32)    # negation of the failed literal
            w = np.concatenate([w, np.stack([neg, np.zeros_like(neg)], axis=1).ravel()])
        assert got["status"] == status, name
        assert got["units"] == units, name
        if status == 0:
            assert got["rounds"] == rounds, name
            assert np.array_equal(t.Level0(), level0), name
            seen_multi |= rounds > 1
        else:
            assert t.Solve([]) == TToporReturnVal.RET_UNSAT
    assert seen_multi


def test_probe_products_match_the_golden_implied_sets(golden):
    """Necessary literals and equivalences (survey 'next' row N4) computed on the device from the pass's implied sets
    equal the same products derived on the CPU from the REFERENCE's golden implied sets."""
    total_nec = total_eq = 0
    for name in ["fuzz_10", "fuzz_5", "fuzz_3", "fuzz_7", "regr_47", "regr_72", "regr_20", "hand_long", "ddtbug_32499"]:
        g = lambda k: golden[f"{name}/{k}"]
        flag, off, lits = g("probe_flag"), g("probe_off"), g("probe_lits")
        nec, eq = set(), set()
        for v in range(flag.size // 2):
            if flag[2 * v] != 0 or flag[2 * v + 1] != 0:
                continue
            a = set(lits[int(off[2 * v]):int(off[2 * v + 1])].tolist())
            b = set(lits[int(off[2 * v + 1]):int(off[2 * v + 2])].tolist())
            nec |= a & b
            eq |= {(v + 1, x) for x in a if -x in b and abs(x) != v + 1}
        t, rc = engine_for(g("words"))
        for kw in (dict(), dict(off32=True)):
            t.ProbeAll(fetch=False, **kw)
            got_nec, got_eq = t.ProbeProducts(equiv_cap=4)     # small buffer: the call grows it
            assert sorted(got_nec.tolist(), key=lambda l: (abs(l), l < 0)) == sorted(nec, key=lambda l: (abs(l), l < 0)), name
            assert {tuple(p) for p in got_eq.tolist()} == eq and len(got_eq) == len(eq), name
        total_nec += len(nec)
        total_eq += len(eq)
        t.ProbeLits(np.asarray([1], dtype=np.int32))
        assert t.lib.bbcp_probe_products(t.handle, None, None, None, 0, None) == -4      # not an all-literal pass
    assert total_nec > 0 and total_eq > 0


def test_decided_cubes_that_overflow_the_state_leave_it_clean(golden):
    """A cube longer than the tier's trail capacity that is ALSO decided without propagation (it names an unmapped
    variable / a literal false at level 0) must not leave table entries behind for the probes the same warp runs
    next: interleave many such cubes with every depth-1 probe and compare the probes with the goldens."""
    for name in ["fuzz_10", "fuzz_5"]:
        g = lambda k: golden[f"{name}/{k}"]
        t, _ = engine_for(g("words"))
        n = 2 * t.max_var
        free = [v for v in range(1, t.max_var + 1) if g("probe_flag")[2 * (v - 1)] in (0, 1)]
        assert len(free) > 100
        poison = [l * (1 if i % 3 else -1) for i, l in enumerate(free[:100])] + [t.max_var + 7]     # 100 literals + an unmapped variable
        cubes = []
        for i in range(n):
            v = i // 2 + 1
            cubes.append([v if i % 2 == 0 else -v])
            if i % 4 == 0:
                cubes.append(poison)
        lits, off = fileio.cubes_to_csr(cubes)
        res = t.PropagateBatch(lits, off, small_trail=32, mid_trail=32)
        probe_idx = [k for k, c in enumerate(cubes) if len(c) == 1]
        assert all(res.flag[k] == 4 for k, c in enumerate(cubes) if len(c) > 1)
        assert np.array_equal(res.flag[probe_idx], g("probe_flag")), name
        sizes = np.diff(res.impl_off)[probe_idx]
        assert np.array_equal(sizes, np.diff(g("probe_off"))), name
        got = np.concatenate([res.implied(k) for k in probe_idx]) if probe_idx else np.zeros(0, np.int32)
        assert np.array_equal(got, g("probe_lits")), name


def test_seeded_random_instances_against_checker():
    """60 seeded random instances (20-80 variables, clause lengths 1-7, duplicate literals, tautologies, units that
    propagate at level 0, sometimes refuted at level 0): every probe and 40 random cubes each (depth 1-10, with
    repeated, contradictory, level-0 and unmapped literals) against the CPU checker, default and tiny state budgets."""
    from oracle_lib import Oracle
    rng = np.random.default_rng(424242)
    n_items = n_conf = n_unsat = 0
    for trial in range(60):
        nv = int(rng.integers(20, 81))
        words = []
        for _ in range(int(rng.integers(nv, 4 * nv))):
            k = int(rng.choice([1, 2, 2, 2, 2, 3, 3, 3, 4, 5, 7], p=[.02, .2, .1, .1, .08, .15, .1, .1, .08, .05, .02]))
            lits = [int(rng.integers(1, nv + 1)) * (1 if rng.random() < 0.5 else -1) for _ in range(k)]
            if rng.random() < 0.05:
                lits.append(lits[0])
            if rng.random() < 0.03:
                lits.append(-lits[0])
            words += lits + [0]
        words = np.asarray(words, dtype=np.int32)
        o = Oracle(words)
        t, rc = engine_for(words)
        assert rc == o.status, trial
        if rc:
            n_unsat += 1
            continue
        assert np.array_equal(t.Level0(), o.level0()), trial
        flag, _, off, lits = o.run()
        cubes = []
        for _ in range(40):
            d = int(rng.integers(1, 11))
            c = [int(rng.integers(1, nv + 4)) * (1 if rng.random() < 0.5 else -1) for _ in range(d)]
            if rng.random() < 0.2:
                c.append(c[0])
            if rng.random() < 0.1:
                c.append(-c[0])
            cubes.append(c)
        cl, co = fileio.cubes_to_csr(cubes)
        cflag, _, coff, clits = o.run(cl, co)
        for kw in (dict(), dict(small_trail=32, mid_trail=32)):
            res = t.ProbeAll(**kw)
            gpu_util.assert_same_results(res, flag, off, lits, f"trial {trial} probes {kw}")
            res = t.PropagateBatch(cl, co, **kw)
            gpu_util.assert_same_results(res, cflag, coff, clits, f"trial {trial} cubes {kw}")
        n_items += flag.size + cflag.size
        n_conf += int((flag == 1).sum() + (cflag == 1).sum())
    assert n_items > 3000 and n_conf > 300 and n_unsat > 5


def test_counters_against_the_reference(tmp_path):
    """SURVEY 8a A14. m_Assignments (TopiAsg.cc:49), m_BCPs (TopiBcp.cc:176) and m_Implications (TopiBcp.cc:193-198)
    of the reference's probe loop against the engine's assignments / bcps / props_ref.
    Exact where the reference's own numbers do not depend on its propagation order and watch history: an instance
    of binary clauses only (every occurrence is a watch, so "row ever allocated" == "literal occurs") without failed
    literals. On mixed instances the engine counts a popped literal whose negation OCCURS in a clause, the reference
    one whose negation has ever WATCHED a clause, so engine >= reference on the probes that reach a fixpoint; the
    trail lengths of those probes (assignments_ok) are the implied-set sizes on both sides."""
    rng = np.random.default_rng(7)
    n, m = 4000, 9000
    a = rng.integers(1, n, size=m)
    b = a + rng.integers(1, 40, size=m)
    keep = b <= n
    words = np.stack([-a[keep], b[keep], np.zeros(keep.sum(), dtype=np.int64)], axis=1).astype(np.int32).ravel()   # a -> b, a < b: no failed literal
    inst = str(tmp_path / "dag.bcnf")
    fileio.write_bcnf(inst, words)
    ref, rinfo = gpu_util.run_checker(inst)
    assert rinfo["conflicts"] == 0 and rinfo["implied_lits"] > 3 * rinfo["ran"]
    t, rc = engine_for(words)
    assert rc == 0
    for kw in (dict(), dict(no_reuse=True), dict(small_trail=32, mid_trail=32)):
        res = t.ProbeAll(**kw)
        gpu_util.assert_same_results(res, ref.flag, ref.off, ref.lits, f"dag {kw}")
        i = res.info
        assert i["bcps"] == int((ref.flag <= 1).sum()), kw
        assert i["assignments"] == i["assignments_ok"] == rinfo["assignments"] == i["n_implied"], kw
        # (props_ref also holds the work of probes that outgrew a tier and were handed on; props_ref_ok counts the final run only)
        assert i["props_ref"] >= i["props_ref_ok"] == rinfo["implications"] == int(ref.props[ref.flag == 0].sum()), kw
    # mixed clause lengths, failed literals: the order-independent parts
    with tempfile.TemporaryDirectory() as d:
        for kind, kw in (("plaw", dict(n=6000, m=24000, seed=5)), ("aig", dict(frames=3, gates=1500, inputs=60, latches=60, window=200, seed=2))):
            inst, _ = gpu_util.gen(kind, d, **kw)
            ref, rinfo = gpu_util.run_checker(inst)
            t, rc = engine_for(fileio.read_bcnf(inst))
            res = t.ProbeAll()
            i = res.info
            ok = ref.flag == 0
            assert i["assignments_ok"] == i["n_implied"] == int(np.diff(ref.off)[ok].sum())
            assert i["bcps"] == int((ref.flag <= 1).sum())
            assert i["props_ref_ok"] >= int(ref.props[ok].sum()) > 0
            assert i["assignments"] >= i["assignments_ok"] and i["props_ref"] >= i["props_ref_ok"]


def test_device_unit_feedback_equals_host_rebuild(tmp_path):
    """bbcp_probe_to_fixpoint applies the failed-literal units ON THE DEVICE (level-0 facts marked in litinfo, rows left as
    they are); the round-1 loop re-finalized (host rebuild) after every round. Same rounds, same units, same level-0
    assignment, and afterwards the same probe results as a freshly built database with the units added."""
    import json, subprocess, sys
    multi = 0
    for kind, kw in (("plaw", dict(n=20000, m=80000, seed=5)), ("aig", dict(frames=4, gates=3000, inputs=100, latches=100, window=300, seed=2)),
                     ("comm", dict(n=20000, m=70000, comms=20, seed=3))):
        inst, _ = gpu_util.gen(kind, str(tmp_path), **kw)
        words = fileio.read_bcnf(inst)
        t, rc = engine_for(words)
        got = t.ProbeToFixpoint()
        prog = ("import sys, json; sys.path.insert(0, %r)\n"
                "from intel_sat_solver_b200 import BatchedTopor, fileio\n"
                "t = BatchedTopor(); t.AddClauses(fileio.read_bcnf(%r)); t.Finalize(); g = t.ProbeToFixpoint()\n"
                "print(json.dumps({'got': g, 'level0': t.Level0().tolist()}))" % (gpu_util.ROOT, inst))
        ref = json.loads(subprocess.run([sys.executable, "-c", prog], check=True, capture_output=True, text=True, env=dict(os.environ, BBCP_FIXPOINT="rebuild")).stdout)
        assert got == ref["got"], kind
        multi += got["status"] == 0 and got["rounds"] >= 2 and got["units"] > 0
        _after_fixpoint(t, words, got, ref)
    assert multi >= 1


def _after_fixpoint(t, words, got, ref):
    if got["status"] == 0:
        assert t.Level0().tolist() == ref["level0"]
        # the database with level-0 facts newer than its rows answers like a rebuilt one
        res = t.ProbeAll()
        t2 = BatchedTopor()
        t2.AddClauses(words)
        t2.AddClauses(np.stack([t.Level0(), np.zeros_like(t.Level0())], axis=1).ravel())
        assert t2.Finalize() == 0
        res2 = t2.ProbeAll()
        gpu_util.assert_same_results(res, res2.flag, res2.impl_off, res2.impl_lits, "dirty level 0 vs rebuilt")
        cl = np.asarray([5, -7, 11, 0, -5, 0, 13, 17, -19, 23], dtype=np.int32)
        co = np.asarray([0, 3, 4, 5, 10], dtype=np.uint64)
        ra, rb = t.PropagateBatch(cl, co), t2.PropagateBatch(cl, co)
        gpu_util.assert_same_results(ra, rb.flag, rb.impl_off, rb.impl_lits, "dirty level 0 vs rebuilt, cubes")


def test_compact_read_back_expands_to_the_same_results(tmp_path):
    """BBCP_OPT_COMPACT_SELF: self-implying probes come back as a flag bit instead of a stored one-literal set (and no
    offsets at all when nothing was stored); expanded on the host the results are the standard ones, bit for bit."""
    for kind, kw in (("3sat", dict(n=30000, m=126000, seed=1)), ("plaw", dict(n=20000, m=80000, seed=5)), ("aig", dict(frames=2, gates=2000, inputs=80, latches=80, window=200, seed=2))):
        inst, _ = gpu_util.gen(kind, str(tmp_path), **kw)
        t, rc = engine_for(fileio.read_bcnf(inst))
        n = 2 * t.max_var
        idx = np.arange(n)
        lits = np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)
        lits[::97] = 0                      # a 0 in a probe list names no variable
        std = t.ProbeLits(lits)
        for off32 in (False, True):
            off_buf = np.full(n + 1, 0xdeadbeef, dtype=np.uint32 if off32 else np.uint64)
            raw = t.ProbeLitsToHost(lits, impl_cap=std.info["n_implied"] + 16, off_out=off_buf, compact_self=True, off32=off32)
            assert (raw.flag & 0x80).any()
            if kind == "3sat":
                assert raw.info["n_implied"] == 0 and (off_buf[:n] == 0xdeadbeef).all()   # nothing stored: offsets never travel
            exp = BatchedTopor.ExpandCompact(lits, raw)
            gpu_util.assert_same_results(exp, std.flag, std.impl_off, std.impl_lits, f"{kind} compact off32={off32}")


def test_no_missed_implication_checker(golden, tmp_path):
    """The static analogue of WLAssertNoMissedImplications (TopiWL.cc:330-442): under every fixpoint the engine reports,
    no clause of the database is unsatisfied with fewer than two unassigned literals. Run over the regression corpus, the
    synthetic shapes, tiny tiers (every tier's code path), cubes, and a database whose level-0 facts are newer than its rows;
    a deliberately truncated implied set must be caught."""
    total = 0
    for name in ["regr_47", "regr_72", "regr_20", "fuzz_10", "fuzz_5", "hand_long", "ddtbug_32499"]:
        t, rc = engine_for(golden[f"{name}/words"])
        for kw in (dict(), dict(small_trail=32, mid_trail=32)):
            t.ProbeAll(fetch=False, **kw)
            r = t.CheckFixpoints()
            assert r["violations"] == 0 and r["items_checked"] > 0, (name, kw, r)
            total += r["clauses_checked"]
        t.PropagateBatch(golden[f"{name}/cube_lits"], golden[f"{name}/cube_off"], fetch=False)
        assert t.CheckFixpoints()["violations"] == 0, name
    for kind, kw in (("aig", dict(frames=3, gates=2000, inputs=80, latches=80, window=200, seed=2)), ("plaw", dict(n=8000, m=32000, seed=5)),
                     ("comm", dict(n=8000, m=32000, comms=10, seed=3))):
        inst, _ = gpu_util.gen(kind, str(tmp_path), **kw)
        t, rc = engine_for(fileio.read_bcnf(inst))
        for okw in (dict(), dict(small_trail=32), dict(small_trail=32, mid_trail=64), dict(no_reuse=True)):
            t.ProbeAll(fetch=False, **okw)
            r = t.CheckFixpoints(stride=7)
            assert r["violations"] == 0 and r["items_checked"] > 100, (kind, okw, r)
            total += r["clauses_checked"]
        if kind == "plaw":
            t.ProbeToFixpoint()
            t.ProbeAll(fetch=False)
            assert t.CheckFixpoints(stride=5)["violations"] == 0
    assert total > 10_000_000
    # negative control: an engine whose results are cut short is caught (drop the last literal of every set on the device)
    import torch
    t, rc = engine_for(golden["fuzz_10/words"])
    res = t.ProbeAll()
    off = torch.as_tensor(res.impl_off.astype(np.int64))
    sizes = np.diff(res.impl_off.astype(np.int64))
    big = np.nonzero(sizes > 3)[0]
    assert big.size > 0
    from intel_sat_solver_b200._lib import load
    ptr = load().bbcp_device_impl_lits(t.handle)
    class Arr:
        __cuda_array_interface__ = {"shape": (int(res.impl_off[-1]),), "typestr": "<i4", "data": (ptr, False), "version": 3}
    dev = torch.as_tensor(Arr(), device="cuda")
    # overwrite the last literal of each large set with a duplicate of its first one: the set loses a member
    last = torch.as_tensor(res.impl_off[big + 1].astype(np.int64) - 1, device="cuda")
    firsts = torch.as_tensor(res.impl_off[big].astype(np.int64), device="cuda")
    dev[last] = dev[firsts]
    torch.cuda.synchronize()
    assert t.CheckFixpoints()["violations"] > 0


def test_hyper_binary_resolvents_against_the_golden_implied_sets(golden):
    """N4: (v, x) is reported iff x is in the REFERENCE's implied set of probe v (golden) and x is not reachable from v
    through the binary clauses of the level-0-simplified database alone (a plain CPU closure over the engine's host rows)."""
    total = 0
    for name in ["fuzz_10", "fuzz_5", "fuzz_3", "regr_47", "regr_72", "regr_20", "hand_long", "ddtbug_32499"]:
        g = lambda k: golden[f"{name}/{k}"]
        t, rc = engine_for(g("words"))
        assert rc == 0
        got = {(int(a), int(b)) for a, b in t.ProbeHbr()}
        nl = 2 * t.max_var + 2
        succ = {}
        for l in range(2, nl):
            row = t.HostRow(l ^ 1).tolist()        # the row walked when l becomes true
            succ[l] = row[3:3 + row[0]]
        ext = lambda il: -(il >> 1) if il & 1 else il >> 1
        flag, off, lits = g("probe_flag"), g("probe_off"), g("probe_lits")
        expect = set()
        for i in range(flag.size):
            if flag[i] != 0:
                continue
            v = i // 2 + 1
            root = 2 * v + (i & 1)
            seen, stack, conflict = {root}, [root], False
            while stack:
                p = stack.pop()
                for q in succ[p]:
                    if q ^ 1 in seen:
                        conflict = True
                    if q not in seen:
                        seen.add(q); stack.append(q)
            assert not conflict, (name, i)          # a probe that reaches a fixpoint cannot fail through binaries alone
            for x in lits[int(off[i]):int(off[i + 1])].tolist():
                il = 2 * abs(x) + (x < 0)
                if il not in seen:
                    expect.add((ext(root), x))
        assert got == expect, (name, len(got), len(expect))
        total += len(expect)
        # the binary-only pass on its own: its sets are the closures computed above
        rb = t.ProbeAll(binary_only=True)
        for i in np.nonzero(flag == 0)[0][:50]:
            assert rb.flag[i] == 0 and set(rb.implied(int(i)).tolist()) <= set(lits[int(off[i]):int(off[i + 1])].tolist())
    assert total > 100


def test_refinalize_with_more_variables_after_a_last_tier_batch():
    """the last tier's slabs are sized by the variable range. A batch that used them, then clauses that raise
    max_var plus more level-0 units than the second tier holds (so that the internal level-0 unit batch of the next
    Finalize reaches the last tier), must not index the old, smaller slabs."""
    from oracle_lib import Oracle
    with tempfile.TemporaryDirectory() as d:
        inst, _ = gpu_util.gen("aig", d, frames=2, gates=1500, inputs=60, latches=60, window=200, seed=2)   # implied sets of hundreds of literals
        words = fileio.read_bcnf(inst)
    t, rc = engine_for(words)
    res = t.ProbeAll(small_trail=32, mid_trail=32)
    assert res.info["n_large"] > 0                      # the last tier ran: slabs exist for the old variable range
    mv = t.max_var
    extra = np.stack([np.arange(mv + 1, mv + 9001, dtype=np.int32), np.zeros(9000, dtype=np.int32)], axis=1).ravel()   # 9000 new unit clauses
    chain = np.asarray([-(mv + 1), mv + 9001, 0, -(mv + 9001), mv + 9002, 0], dtype=np.int32)                           # and something they propagate into
    t.AddClauses(extra)
    t.AddClauses(chain)
    assert t.Finalize() == 0
    w2 = np.concatenate([words, extra, chain])
    o = Oracle(w2)
    assert o.status == 0 and np.array_equal(t.Level0(), o.level0())
    flag, _, off, lits = o.run()
    res = t.ProbeAll(small_trail=32, mid_trail=32)
    gpu_util.assert_same_results(res, flag, off, lits, "after re-finalize with a larger variable range")
