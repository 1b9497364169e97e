"""The text formats the reference consumes (SURVEY 'next' row N1): DIMACS and the IntelSAT incremental format
with ``s`` query lines (Main.cc:207-221). CPU: parsing. GPU: the propagation-only command-line tool replays a file
made from a golden instance and its golden cubes and prints the reference's result words."""
import os
import subprocess
import sys

import numpy as np
import pytest

from intel_sat_solver_b200 import fileio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TEXT = """c an incremental file
p cnf 5 4
1 -2 0
s 1 2 0
ot 1000
-1 3 4 0
r
2 5
lb 3
s -3 0
5 0
n 7
"""


def test_text_parser_keeps_clause_lines_in_order_and_collects_queries(tmp_path):
    p = tmp_path / "a.cnf"
    p.write_text(TEXT)
    words = fileio.read_text_cnf(str(p))
    assert words.tolist() == [1, -2, 0, -1, 3, 4, 0, 2, 5, 0, 5, 0]     # a clause line without its 0 is closed
    assert fileio.read_text_queries(str(p)) == [[1, 2], [-3]]
    lits, off = fileio.cubes_to_csr([[1, 2], [-3], []])
    assert lits.tolist() == [1, 2, -3] and off.tolist() == [0, 2, 3, 3]
    b = tmp_path / "a.bcnf"
    fileio.write_bcnf(str(b), words)
    assert np.array_equal(fileio.read_instance(str(b)), words) and np.array_equal(fileio.read_instance(str(p)), words)


@pytest.mark.gpu
def test_tool_replays_incremental_file(golden, tmp_path):
    name = "regr_72"
    g = lambda k: golden[f"{name}/{k}"]
    words, off, lits, flag = g("words").tolist(), g("cube_off"), g("cube_lits"), g("cube_flag")
    lines, cur = ["p cnf 0 0"], []
    for w in words:
        cur.append(str(w))
        if w == 0:
            lines.append(" ".join(cur))
            cur = []
    keep = [i for i in range(flag.size) if 0 not in lits[int(off[i]):int(off[i + 1])].tolist() and flag[i] != 4][:50]
    for i in keep:
        lines.append("s " + " ".join(str(int(l)) for l in lits[int(off[i]):int(off[i + 1])]) + " 0")
    p = tmp_path / "f.cnf"
    p.write_text("\n".join(lines) + "\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bbcp_tool.py"), str(p)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = [l for l in r.stdout.splitlines() if l.startswith("s ")]
    assert len(out) == len(keep)
    for i, line in zip(keep, out):
        if flag[i] in (1, 3):
            assert line == "s UNSATISFIABLE"
        else:
            assert line in ("s USER_INTERRUPT", "s SATISFIABLE")
