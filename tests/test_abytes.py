"""The ABYTES lower bound of SURVEY.md 8d, recomputed on the CPU from the database and the checker's fixpoints,
agrees with the closed form bench.py uses for the C2 family (random 3-SAT without binary clauses: every probe
implies itself, no binary rows, every long watch visited: 16 B per probe + 24 B per long watch), and behaves on
instances with binaries, level-0 units and conflicts."""
import os
import subprocess
import tempfile

import numpy as np

import abytes as ab
from oracle_lib import Oracle
from intel_sat_solver_b200 import fileio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_closed_form_for_random_3sat():
    gen = os.path.join(ROOT, "tools", "cnfgen")
    if not os.path.exists(gen):
        subprocess.run(["gcc", "-O2", "-o", gen, os.path.join(ROOT, "tools", "cnfgen.c"), "-lm"], check=True)
    with tempfile.TemporaryDirectory() as d:
        inst = os.path.join(d, "i.bcnf")
        n, m = 3000, 12600
        subprocess.run([gen, "3sat", f"n={n}", f"m={m}", "seed=1", f"out={inst}"], check=True, capture_output=True)
        words = fileio.read_bcnf(inst)
        o = Oracle(words)
        flag, _, off, lits = o.run()
        a = ab.abytes(words, o.level0(), flag, off, lits)
        probes = int((flag == 0).sum())
        skipped = int((flag == 4).sum())
        assert probes + skipped == 2 * n and (np.diff(off)[flag == 0] == 1).all()
        # 8 (row extent) + 4 (trail) + 4 (input) per probe, (8 + 4 * (1 + 3)) per long watch; 2 watches per clause;
        # unmapped probes move their 4 input bytes only
        assert int(a.sum()) == 16 * probes + 4 * skipped + 24 * 2 * m
        # bench.py's closed form (all probes mapped) is the same expression
        assert 16 * (2 * n) + 24 * 2 * m - int(a.sum()) == 12 * skipped


def test_small_instance_by_hand():
    # (1 v 2) (-1 v 3) (-3 v 4 v 5) (-2 v -4 v 6 v 7): probe +1 implies 3; probe -1 implies 2
    words = np.asarray([1, 2, 0, -1, 3, 0, -3, 4, 5, 0, -2, -4, 6, 7, 0], dtype=np.int32)
    o = Oracle(words)
    flag, _, off, lits = o.run()
    a = ab.abytes(words, o.level0(), flag, off, lits)
    # probe +1: T = {1, 3}. row(-1): 1 binary -> 8 + 4; row(-3): watched by the ternary (other watch 4 not true) -> 8 + 8 + 4*(1+3)
    assert a[0] == (8 + 4) + (8 + 8 + 16) + 4 * 2 + 4
    # probe -1: T = {-1, 2}. row(1): 1 binary -> 8 + 4; row(-2): watched by the 4-clause (other watch -4 not true) -> 8 + 8 + 4*(1+4)
    assert a[1] == (8 + 4) + (8 + 8 + 20) + 4 * 2 + 4
    assert len(a) == 14


def test_conflict_probes_count_their_input_only(golden):
    g = lambda k: golden[f"fuzz_10/{k}"]
    a = ab.abytes(g("words"), g("level0"), g("probe_flag"), g("probe_off"), g("probe_lits"))
    assert (a[g("probe_flag") == 1] == 4).all() and a[g("probe_flag") == 0].min() >= 16


def test_c_tool_agrees_with_the_python_definition():
    """oracle/abytes (the C form used by bench.py's workload blocks on the checker's sampled fixpoints) against
    tests/abytes.py on instances with binaries, units, long clauses, conflicts and a strided sample over two shards."""
    import json
    import gpu_util
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
    tool = os.path.join(ROOT, "oracle", "abytes")
    port = os.path.join(ROOT, "oracle", "oracle_bcp")
    with tempfile.TemporaryDirectory() as d:
        for kind, kw, stride in (("aig", dict(frames=3, gates=400, inputs=30, latches=30, window=100, seed=2), 1),
                                 ("plaw", dict(n=1500, m=6000, seed=5), 3), ("comm", dict(n=1500, m=6000, comms=5, seed=3), 2)):
            inst, _ = gpu_util.gen(kind, d, **kw)
            words = fileio.read_bcnf(inst)
            outs = []
            for sh in range(2):
                out = os.path.join(d, f"{kind}{sh}.bres")
                subprocess.run([port, inst, "--out", out, "--probe-all", "--shard", str(sh), "2", "--stride", str(stride)], check=True, capture_output=True)
                outs.append(out)
            got = json.loads(subprocess.run([tool, inst] + outs, check=True, capture_output=True, text=True).stdout)
            o = Oracle(words)
            exp = items = 0
            for out in outs:
                r = fileio.read_bres(out)
                a = ab.abytes(words, o.level0(), r.flag, r.off, r.lits)
                ran = r.flag != 255
                exp += int(a[ran].sum())
                items += int(ran.sum())
            assert got["items"] == items and got["abytes_total"] == exp, kind
