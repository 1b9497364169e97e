"""libbbcp_ipasir.so: the IPASIR names of the reference's C ABI (ToporIpasir.h) over the GPU engine. CPU: the library
loads and exports every IPASIR symbol the reference exports; GPU: ipasir_add / assume / solve / val / failed agree
with the golden cube results."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "intel_sat_solver_b200", "libbbcp_ipasir.so")
NAMES = ["ipasir_signature", "ipasir_init", "ipasir_release", "ipasir_add", "ipasir_assume", "ipasir_solve", "ipasir_val", "ipasir_failed",
         "ipasir_set_terminate", "ipasir_set_learn", "ipasir_set_parameter"]


def load():
    L = C.CDLL(LIB)
    L.ipasir_signature.restype = C.c_char_p
    L.ipasir_init.restype = C.c_void_p
    L.ipasir_release.argtypes = [C.c_void_p]
    L.ipasir_add.argtypes = [C.c_void_p, C.c_int32]
    L.ipasir_assume.argtypes = [C.c_void_p, C.c_int32]
    L.ipasir_solve.argtypes = [C.c_void_p]
    L.ipasir_val.argtypes = [C.c_void_p, C.c_int32]
    L.ipasir_val.restype = C.c_int32
    L.ipasir_failed.argtypes = [C.c_void_p, C.c_int32]
    return L


def test_ipasir_library_exports_the_reference_names():
    L = load()
    for n in NAMES:
        assert hasattr(L, n), n
    assert b"bbcp" in L.ipasir_signature()
    s = L.ipasir_init()
    assert s
    L.ipasir_release(s)


@pytest.mark.gpu
def test_ipasir_calls_agree_with_golden_cubes(golden):
    L = load()
    name = "regr_72"
    g = lambda k: golden[f"{name}/{k}"]
    s = L.ipasir_init()
    for w in g("words").tolist():
        L.ipasir_add(s, int(w))
    off, lits, flag = g("cube_off"), g("cube_lits"), g("cube_flag")
    ioff, ilits = g("cube_impl_off"), g("cube_impl_lits")
    level0 = set(g("level0").tolist())
    checked = 0
    for i in range(min(flag.size, 60)):
        cube = lits[int(off[i]):int(off[i + 1])].tolist()
        if flag[i] == 4 or 0 in cube:
            continue
        for l in cube:
            L.ipasir_assume(s, int(l))
        r = L.ipasir_solve(s)
        if flag[i] in (1, 3):
            assert r == 20
            assert all(L.ipasir_failed(s, int(l)) == 1 for l in cube)
            continue
        assert r in (0, 10)
        implied = set(ilits[int(ioff[i]):int(ioff[i + 1])].tolist()) | level0
        for l in implied:
            assert L.ipasir_val(s, int(l)) == l and L.ipasir_val(s, int(-l)) == l
        assert L.ipasir_failed(s, int(cube[0])) == 0 if cube else True
        checked += 1
    assert checked > 10
    L.ipasir_release(s)
