"""CPU tests: the plain-C restatement (oracle/bcp_oracle.c) against the golden vectors produced by the
reference's own BCP (tests/golden/make_golden.py), plus an independent naive closure."""
import json
import os

import numpy as np
import pytest

from oracle_lib import Oracle, naive_up

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_manifest_matches_survey_counts():
    m = json.load(open(os.path.join(ROOT, "tests", "golden", "MANIFEST.json")))
    # 4,694 valid probes, 77 failed literals, 2,996 clauses over regr_1..72
    assert m["regr_totals"] == {"probes_valid": 4694, "failed": 77, "clauses": 2996}
    inst = m["instances"]
    assert (inst["regr_20"]["probes_valid"], inst["regr_20"]["failed"]) == (22, 2)
    assert (inst["regr_47"]["probes_valid"], inst["regr_47"]["failed"]) == (90, 16)
    assert (inst["regr_61"]["probes_valid"], inst["regr_61"]["failed"]) == (300, 0)
    assert (inst["regr_72"]["probes_valid"], inst["regr_72"]["failed"]) == (108, 7)


def test_oracle_matches_reference_on_all_goldens(golden, golden_names):
    n_probes = n_cubes = 0
    for name in golden_names:
        g = lambda k: golden[f"{name}/{k}"]
        o = Oracle(g("words"))
        assert o.status == int(g("status")[0]), name
        if o.status:
            continue
        assert np.array_equal(o.level0(), g("level0")), name
        flag, _, off, lits = o.run()
        assert np.array_equal(flag, g("probe_flag")), name
        assert np.array_equal(off, g("probe_off")), name
        assert np.array_equal(lits, g("probe_lits")), name
        flag, _, off, lits = o.run(g("cube_lits"), g("cube_off"))
        assert np.array_equal(flag, g("cube_flag")), name
        assert np.array_equal(off, g("cube_impl_off")), name
        assert np.array_equal(lits, g("cube_impl_lits")), name
        n_probes += int((g("probe_flag") <= 1).sum())
        n_cubes += g("cube_flag").size
    assert n_probes > 14000 and n_cubes > 5000


@pytest.mark.parametrize("name", ["regr_20", "regr_47", "regr_72", "hand_units", "hand_alias", "hand_long", "hand_taut_dup"])
def test_naive_closure_agrees(golden, name):
    words = golden[f"{name}/words"]
    level0 = set(golden[f"{name}/level0"].tolist())
    flag, off, lits = golden[f"{name}/probe_flag"], golden[f"{name}/probe_off"], golden[f"{name}/probe_lits"]
    for i in range(flag.size):
        if flag[i] > 1:
            continue
        l = (i // 2 + 1) * (-1 if i & 1 else 1)
        conflict, true = naive_up(words, [l])
        assert conflict == bool(flag[i]), (name, l)
        if not conflict:
            assert sorted(true - level0, key=abs) == lits[int(off[i]):int(off[i + 1])].tolist(), (name, l)


def test_shard_and_stride_select_probes(golden):
    o = Oracle(golden["regr_72/words"])
    n = 2 * o.max_var
    flag, _, _, _ = o.run(lo=10, hi=50, stride=4)
    ran = np.nonzero(flag != 255)[0]
    assert ran.tolist() == list(range(10, 50, 4))
    full, _, _, _ = Oracle(golden["regr_72/words"]).run()
    assert np.array_equal(flag[ran], full[ran]) and full.size == n
