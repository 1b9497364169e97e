#!/usr/bin/env python
"""Generate the committed golden vectors by EXECUTING THE REFERENCE (route A).

Run in the build container only (needs /root/reference and oracle/_ref/ref_bcp built by
``make -C oracle ref``):

    python tests/golden/make_golden.py

For every instance it stores the clause stream that was fed to the reference and what the reference's own
CTopi::BCP() produced: level-0 status and literals, the per-probe flag and implied set for every literal of
the probe universe (+1,-1,+2,-2,...), and the same for a cube batch made of (a) the file's own ``s`` query
lines, (b) seeded random cubes of depth 1..4 and of depth 20 (with duplicate / contradictory / level-0
literals arising naturally). Instances: regression_instances/regr_1..72 + the two ddtbug reproducers of the
reference, 12 seeded ``cnfuzz_incr`` outputs (the reference's own fuzzer, third_party/cnfuzzdd2013), and a
few hand-written intake edge cases (units, tautologies, duplicates, empty clause, level-0 conflicts).

Output: tests/golden/golden.npz  (keys "<instance>/<field>") and tests/golden/MANIFEST.json (sha256, counts).
"""
import glob
import hashlib
import json
import os
import random
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from intel_sat_solver_b200 import fileio  # noqa: E402

REF = "/root/reference"
REF_BCP = os.path.join(ROOT, "oracle", "_ref", "ref_bcp")
FUZZ = os.path.join(REF, "third_party", "cnfuzzdd2013", "cnfuzz_incr")

HAND = {
    # units chain to a level-0 fixpoint, then ordinary clauses
    "hand_units": "1 0\n-1 2 0\n-2 3 4 0\n-4 5 0\n5 6 7 0\n-6 -7 8 0\n-8 9 0\n",
    # tautology introduces vars 7,8 only -> unmapped; duplicate literals; duplicate binaries
    "hand_taut_dup": "1 2 2 3 0\n7 -7 8 0\n-1 -1 4 0\n-1 4 0\n4 -1 0\n-4 5 6 0\n9 1 -9 0\n-5 -6 0\n",
    # contradictory at level 0 through propagation
    "hand_unsat_l0": "1 0\n-1 2 0\n-2 3 0\n-3 -1 0\n4 5 0\n",
    # contradictory by complementary units
    "hand_unsat_units": "1 2 0\n3 0\n-3 0\n",
    # an explicit empty clause
    "hand_empty": "1 2 0\n0\n3 4 0\n",
    # fresh-variable unit after a level-0 literal exists (aliasing path, Topi.cc:559-572), then reuse of that var
    "hand_alias": "1 0\n9 0\n-9 2 3 0\n-2 4 0\n-3 4 0\n-10 0\n10 5 6 0\n-5 -4 0\n",
    # long clauses only, forcing watch moves
    "hand_long": "1 2 3 4 5 6 0\n-1 -2 -3 -4 -5 -6 0\n1 -2 3 -4 5 -6 7 8 0\n-7 -8 1 0\n-1 7 0\n-1 8 0\n2 3 4 5 0\n-2 9 0\n-3 9 0\n-9 -4 -5 1 0\n",
}


def run_ref(words, cubes=None):
    with tempfile.TemporaryDirectory() as d:
        inst = os.path.join(d, "i.bcnf")
        fileio.write_bcnf(inst, words)
        cmd = [REF_BCP, inst, "--out", os.path.join(d, "o.bres")]
        if cubes is not None:
            lits, off = fileio.cubes_to_csr(cubes)
            fileio.write_cubes(os.path.join(d, "c.cubes"), lits, off)
            cmd += ["--cubes", os.path.join(d, "c.cubes")]
        info = json.loads(subprocess.run(cmd, check=True, capture_output=True, text=True).stdout)
        assert info["delayed_triggers"] == 0, "SURVEY 8a row A9 violated"
        return fileio.read_bres(os.path.join(d, "o.bres"))


def make_cubes(name, words, queries, level0):
    rng = random.Random(hashlib.sha256(name.encode()).digest())
    maxv = int(np.abs(words).max()) if words.size else 1
    cubes = [q for q in queries]
    def rl():
        v = rng.randint(1, maxv + 1)  # +1: sometimes an unmapped variable
        return v if rng.random() < 0.5 else -v
    for _ in range(40):
        cubes.append([rl() for _ in range(rng.randint(1, 4))])
    for _ in range(12):
        cubes.append([rl() for _ in range(20)])
    for _ in range(4):  # with an explicit duplicate and a level-0 literal when there is one
        c = [rl() for _ in range(3)]
        c.append(c[0])
        if len(level0):
            c.append(int(level0[rng.randrange(len(level0))]))
        cubes.append(c)
    cubes.append([])  # the empty cube
    return cubes


def main():
    insts = {}
    for p in sorted(glob.glob(os.path.join(REF, "regression_instances", "*.cnf"))):
        name = os.path.basename(p)[:-4]
        if name.startswith("ddtbug"):
            name = "ddtbug_" + name.split("_")[-1]
        insts[name] = (fileio.read_text_cnf(p), fileio.read_text_queries(p))
    for seed in range(1, 13):
        with tempfile.NamedTemporaryFile("w+", suffix=".cnf") as f:
            subprocess.run([FUZZ, str(seed)], check=True, stdout=f)
            f.flush()
            insts[f"fuzz_{seed}"] = (fileio.read_text_cnf(f.name), fileio.read_text_queries(f.name))
    for name, txt in HAND.items():
        with tempfile.NamedTemporaryFile("w+", suffix=".cnf") as f:
            f.write(txt)
            f.flush()
            insts[name] = (fileio.read_text_cnf(f.name), [])

    out = {}
    manifest = {}
    for name, (words, queries) in insts.items():
        pr = run_ref(words)
        cubes = make_cubes(name, words, queries, pr.level0)
        cr = run_ref(words, cubes)
        cl, co = fileio.cubes_to_csr(cubes)
        out[f"{name}/words"] = words
        out[f"{name}/status"] = np.asarray([pr.status], dtype=np.int32)
        out[f"{name}/level0"] = pr.level0
        out[f"{name}/probe_flag"] = pr.flag
        out[f"{name}/probe_off"] = pr.off
        out[f"{name}/probe_lits"] = pr.lits
        out[f"{name}/cube_lits"] = cl
        out[f"{name}/cube_off"] = co
        out[f"{name}/cube_flag"] = cr.flag
        out[f"{name}/cube_impl_off"] = cr.off
        out[f"{name}/cube_impl_lits"] = cr.lits
        valid = int(((pr.flag == 0) | (pr.flag == 1)).sum())
        manifest[name] = {"clauses": int((words == 0).sum()), "status": pr.status, "level0": int(pr.level0.size),
                          "probes_valid": valid, "failed": int((pr.flag == 1).sum()),
                          "implied_lits": int(pr.lits.size), "cubes": len(cubes),
                          "cube_conflicts": int(((cr.flag == 1) | (cr.flag == 3)).sum())}
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    np.savez_compressed(path, **out)
    tot = {k: sum(m[k] for n, m in manifest.items() if n.startswith("regr_")) for k in ("probes_valid", "failed", "clauses")}
    with open(os.path.join(ROOT, "tests", "golden", "MANIFEST.json"), "w") as f:
        json.dump({"sha256": hashlib.sha256(open(path, "rb").read()).hexdigest(), "regr_totals": tot, "instances": manifest}, f, indent=1, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes;", len(insts), "instances; regr totals", tot)


if __name__ == "__main__":
    main()
