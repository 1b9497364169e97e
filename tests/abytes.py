"""ABYTES of SURVEY.md 8d -- TEST INFRASTRUCTURE (uses the CPU checker for the fixpoints).

The order-independent lower bound on the bytes any two-watched-literal BCP must move for a batch of probes/cubes,
computed on the CPU from the clause database and the fixpoints alone, so that the figure bench.py reports as
``roofline.abytes_2wl_*`` can be recomputed without trusting engine counters:

    ABYTES(probe) = sum over p in T of [ 8                      row-extent pair of not-p
                                        + 4 * nbin(not-p)        binary row entries
                                        + 8 * nlong(not-p)       (blocker, clauseRef) entries
                                        + sum over long clauses C watched by not-p whose other watch is not true
                                          in the fixpoint of 4 * (1 + |C|) ]                       clause fetch
                    + 4 * |T|                                    trail / result write
                    + 4 * |cube|                                 input
    T = literals true at the probe's fixpoint and not true at level 0. Conflict probes: 4 * |cube| only.

on the level-0-simplified database (clauses satisfied at level 0 removed, level-0-false literals stripped,
duplicate binaries kept once), canonical watches = the first two remaining literals of each clause of >= 3
literals in input order.

    python tests/abytes.py <instance.bcnf|.cnf> [--stride k]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def simplified_clauses(words, level0):
    """level-0-simplified clause list (lists of external literals), duplicate binaries kept once."""
    true0 = set(int(l) for l in level0)
    out, seen_bin, cur = [], set(), []
    for w in np.asarray(words).tolist():
        if w != 0:
            cur.append(w)
            continue
        c, cur = cur, []
        lits, taut = [], False
        for l in c:
            if -l in lits:
                taut = True
                break
            if l not in lits:
                lits.append(l)
        if taut or any(l in true0 for l in lits):
            continue
        lits = [l for l in lits if -l not in true0]
        if len(lits) < 2:
            continue
        if len(lits) == 2:
            key = (min(lits), max(lits))
            if key in seen_bin:
                continue
            seen_bin.add(key)
        out.append(lits)
    return out


def abytes(words, level0, flag, off, lits, cube_off=None):
    """Per-item ABYTES (numpy uint64 array). Probe universe when cube_off is None (|cube| = 1)."""
    cls = simplified_clauses(words, level0)
    nbin, watched = {}, {}
    for c in cls:
        if len(c) == 2:
            for l in c:
                nbin[l] = nbin.get(l, 0) + 1
        else:
            watched.setdefault(c[0], []).append((c[1], len(c)))
            watched.setdefault(c[1], []).append((c[0], len(c)))
    n = flag.size
    out = np.zeros(n, dtype=np.uint64)
    for i in range(n):
        cube = 1 if cube_off is None else int(cube_off[i + 1] - cube_off[i])
        total = 4 * cube
        if flag[i] == 0:
            T = lits[int(off[i]):int(off[i + 1])].tolist()
            Tset = set(T)
            for p in T:
                q = -p                      # the row of not-p is scanned when p becomes true
                longs = watched.get(q, ())
                total += 8 + 4 * nbin.get(q, 0) + 8 * len(longs)
                total += sum(4 * (1 + ln) for other, ln in longs if other not in Tset)
            total += 4 * len(T)
        out[i] = total
    return out


def main(argv):
    from intel_sat_solver_b200 import fileio
    from oracle_lib import Oracle
    if len(argv) < 2:
        print(__doc__)
        return 2
    words = fileio.read_instance(argv[1])
    stride = int(argv[argv.index("--stride") + 1]) if "--stride" in argv else 1
    o = Oracle(words)
    flag, _, off, lits = o.run(stride=stride)
    a = abytes(words, o.level0(), flag, off, lits)
    ran = flag != 255
    print({"probes": int(ran.sum()), "abytes_total": int(a[ran].sum()), "abytes_per_probe": float(a[ran].mean())})
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
