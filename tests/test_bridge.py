"""Live-solver bridge (survey 'next' row N3, include/bbcp_bridge.hpp). CPU: the template compiles and links against
a stand-in solver and refuses to run without a device. GPU: bound to the UNMODIFIED reference solver
(oracle/_ref/bridge_demo, built in the build container from the reference sources where they lie), the reference
gives the same SAT/UNSAT answers with the GPU probing pass in front of it, and the GPU cube filter drops only
cubes the reference refutes."""
import json
import os
import subprocess

import numpy as np
import pytest

from intel_sat_solver_b200 import fileio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "bridge_demo")
LIBDIR = os.path.join(ROOT, "intel_sat_solver_b200")


def test_bridge_template_compiles_and_fails_loudly_without_device(tmp_path):
    exe = str(tmp_path / "bridge_mock")
    subprocess.run(["g++", "-std=c++20", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "bridge_mock_main.cpp"),
                    "-L", LIBDIR, "-lbbcp", f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True)
    import torch
    if not torch.cuda.is_available():
        assert subprocess.run([exe]).returncode == 3


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["regr_72", "regr_47", "regr_20", "fuzz_10", "fuzz_5", "hand_long"])
def test_reference_answers_the_same_behind_the_bridge(golden, tmp_path, name):
    if not os.path.exists(DEMO):
        pytest.skip("oracle/_ref/bridge_demo was not built (needs the reference sources)")
    g = lambda k: golden[f"{name}/{k}"]
    inst, cubes = str(tmp_path / "i.bcnf"), str(tmp_path / "c.cubes")
    fileio.write_bcnf(inst, g("words"))
    off, lits = g("cube_off"), g("cube_lits")
    keep = [i for i in range(off.size - 1) if 0 not in lits[int(off[i]):int(off[i + 1])].tolist()][:40]
    cl, co = fileio.cubes_to_csr([lits[int(off[i]):int(off[i + 1])].tolist() for i in keep])
    fileio.write_cubes(cubes, cl, co)
    r = subprocess.run([DEMO, inst, "--cubes", cubes], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    out = json.loads(r.stdout)
    assert out["agree"] == out["queries"] == len(keep) and out["filter_wrong"] == 0
    # every cube the goldens flag as refuted by unit propagation is filtered before the CPU sees it
    flags = g("cube_flag")[keep]
    assert out["cubes_filtered"] >= int(np.isin(flags, (1, 3)).sum())     # (">=": the probing units make UP stronger)
    # -1: failed-literal probing to a fixpoint refuted the instance on the GPU -- then the reference must say UNSAT everywhere
    assert out["units_fed"] >= 0 or (out["units_fed"] == -1 and out["unsat"] == out["queries"])


@pytest.mark.gpu
def test_bridge_learnt_clause_ingestion_in_solve_passes_and_drat_units(tmp_path):
    """N3: (a) the reference's learnt clauses (SetCbNewLearntCls, Topi.hpp:120) join the GPU database and GPU passes run
    inside the running Solve on a conflict-count schedule, answers unchanged; (b) every unit fed to the reference is logged
    as a DRAT line in derivation order, and each line is RUP -- re-checked here by unit propagation with the CPU checker:
    the negation of the unit must run into a conflict over the clauses plus the lines before it."""
    if not os.path.exists(DEMO):
        pytest.skip("oracle/_ref/bridge_demo was not built (needs the reference sources)")
    import gpu_util
    from oracle_lib import Oracle
    # (a) a random 3-SAT instance near the threshold: hundreds of conflicts
    inst, _ = gpu_util.gen("3sat", str(tmp_path), n=160, m=680, seed=11)
    r = subprocess.run([DEMO, inst, "--learnts"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    out = json.loads(r.stdout)
    assert out["agree"] == out["queries"] == 1 and out["learnts_taken"] > 0 and out["in_solve_passes"] > 0, out
    # (b) a Tseitin instance: satisfiable, with constant gates = failed literals
    os.rename(inst, str(tmp_path / "a.bcnf"))
    inst, _ = gpu_util.gen("aig", str(tmp_path), frames=2, gates=500, inputs=30, latches=30, window=120, seed=2)
    drat = str(tmp_path / "units.drat")
    r = subprocess.run([DEMO, inst, "--drat", drat, "--learnts"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    out = json.loads(r.stdout)
    lines = [int(l.split()[0]) for l in open(drat) if l.strip()]
    assert out["agree"] == out["queries"] and out["units_fed"] == len(lines) > 0, out
    words = fileio.read_bcnf(inst)
    checked = 0
    for k, u in enumerate(lines[:60]):
        w = np.concatenate([words, np.stack([np.asarray(lines[:k], dtype=np.int32), np.zeros(k, dtype=np.int32)], axis=1).ravel()]).astype(np.int32)
        o = Oracle(w)
        if o.status:          # the lines so far already refute the instance: everything after is trivially RUP
            break
        flag, _, _, _ = o.run(np.asarray([-u], dtype=np.int32), np.asarray([0, 1], dtype=np.uint64))
        assert flag[0] in (1, 3), (k, u, int(flag[0]))      # conflict, or already false at level 0
        checked += 1
    assert checked > 0
