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
