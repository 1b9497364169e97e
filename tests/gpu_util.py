"""Helpers shared by the GPU tests: build tools on demand, run the reference / restated CPU checkers."""
import json
import os
import subprocess
import tempfile

import numpy as np

from intel_sat_solver_b200 import fileio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CNFGEN = os.path.join(ROOT, "tools", "cnfgen")
REF_BCP = os.path.join(ROOT, "oracle", "_ref", "ref_bcp")
PORT_BCP = os.path.join(ROOT, "oracle", "oracle_bcp")


def ensure_tools():
    if not os.path.exists(CNFGEN):
        subprocess.run(["gcc", "-O2", "-o", CNFGEN, os.path.join(ROOT, "tools", "cnfgen.c"), "-lm"], check=True)
    if not os.path.exists(PORT_BCP):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)


def checker_binary():
    """The reference's own BCP when its prebuilt harness travelled with the repo, else the C restatement."""
    ensure_tools()
    return REF_BCP if os.path.exists(REF_BCP) else PORT_BCP


def gen(kind, tmpdir, **kw):
    ensure_tools()
    out = os.path.join(tmpdir, f"{kind}.bcnf")
    args = [CNFGEN, kind] + [f"{k}={v}" for k, v in kw.items()] + [f"out={out}"]
    info = json.loads(subprocess.run(args, check=True, capture_output=True, text=True).stdout)
    return out, info


def run_checker(inst, cubes=None, extra=()):
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.bres")
        cmd = [checker_binary(), inst, "--out", out] + (["--cubes", cubes] if cubes else []) + list(extra)
        info = json.loads(subprocess.run(cmd, check=True, capture_output=True, text=True).stdout)
        return fileio.read_bres(out), info


def assert_same_results(res, flag, off, lits, what=""):
    """res: BatchResult of the engine; flag/off/lits: the checker's. Bit-exact."""
    assert np.array_equal(res.flag, flag), f"{what}: flags differ at {np.nonzero(res.flag != flag)[0][:10]}"
    assert np.array_equal(res.impl_off, off), f"{what}: implied-set sizes differ"
    assert np.array_equal(res.impl_lits, lits), f"{what}: implied literals differ"
