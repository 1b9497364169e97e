#!/usr/bin/env python
"""Propagation-only counterpart of the reference's command-line tool (Main.cc:202-247: ``<exe> <cnf>``) for the
files the reference consumes: DIMACS, or the IntelSAT incremental format where clause lines are interleaved with
``s <assumptions> 0`` query lines (Main.cc:207-221, 479). Clause lines are added in file order; every ``s`` line
becomes one cube of ONE GPU batch over the complete clause set (the convention of SURVEY.md 8c), and one result
line per query is printed with the reference's own words (Main.cc:122,152,177): ``s UNSATISFIABLE`` when unit
propagation of the assumptions derives a conflict, ``s SATISFIABLE`` when it assigns every variable, otherwise
``s USER_INTERRUPT`` (fixpoint reached, the search the reference would start here is out of scope). With
``--probe`` it also runs failed-literal probing to a fixpoint and prints the units found as ``u <lit> 0`` lines.

    python tools/bbcp_tool.py <file.cnf> [--probe]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from intel_sat_solver_b200 import BatchedTopor, fileio  # noqa: E402


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    path = argv[1]
    t = BatchedTopor()
    t.AddClauses(fileio.read_instance(path))
    rc = t.Finalize()
    queries = [] if path.endswith(".bcnf") else fileio.read_text_queries(path)
    if rc == 1:
        for _ in queries or [None]:
            print("s UNSATISFIABLE")
        return 20
    if "--probe" in argv:
        before = set(t.Level0().tolist())
        r = t.ProbeToFixpoint()
        if r["status"] == 1:
            print("s UNSATISFIABLE")
            return 20
        for l in t.Level0().tolist():
            if l not in before:
                print(f"u {l} 0")
        print(f"c probing: {r['rounds']} rounds, {r['units']} failed literals")
    if queries:
        lits, off = fileio.cubes_to_csr(queries)
        res = t.PropagateBatch(lits, off)
        n_assignable = t.DbInfo()["n_mapped_vars"] - len(t.Level0())
        for i in range(len(queries)):
            f = int(res.flag[i])
            if f in (1, 3):
                print("s UNSATISFIABLE")
            elif f in (0, 2) and int(res.impl_off[i + 1] - res.impl_off[i]) == n_assignable:
                print("s SATISFIABLE")
            else:
                print("s USER_INTERRUPT")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
