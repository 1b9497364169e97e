"""ctypes binding of oracle/liboracle_bcp.so (the plain-C restatement). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle_bcp.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_release.argtypes = [C.c_void_p]
        L.orc_add_words.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_add_clause.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_level0.argtypes = [C.c_void_p]
        L.orc_max_user_var.argtypes = [C.c_void_p]
        L.orc_level0_lits.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_level0_lits.restype = C.c_uint64
        L.orc_run_cubes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.orc_free.argtypes = [C.c_void_p]
        for f in ("orc_implications", "orc_assignments", "orc_bcps"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_uint64
        _LIB = L
    return _LIB


class Oracle:
    def __init__(self, words):
        self.L = lib()
        self.h = self.L.orc_create()
        w = np.ascontiguousarray(words, dtype=np.int32)
        self.L.orc_add_words(self.h, w.ctypes.data, w.size)
        self.status = self.L.orc_level0(self.h)
        self.max_var = self.L.orc_max_user_var(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_release(self.h)
            self.h = None

    def level0(self):
        n = self.L.orc_level0_lits(self.h, None, 0)
        out = np.zeros(n, dtype=np.int32)
        self.L.orc_level0_lits(self.h, out.ctypes.data, n)
        return out

    def run(self, cube_lits=None, cube_off=None, lo=0, hi=None, stride=1):
        """cube_lits None -> probe universe. Returns (flag, props, off, lits)."""
        if cube_lits is None:
            n = 2 * self.max_var
            lp = op = None
        else:
            cube_lits = np.ascontiguousarray(cube_lits, dtype=np.int32)
            cube_off = np.ascontiguousarray(cube_off, dtype=np.uint64)
            n = cube_off.size - 1
            keep = np.zeros(max(cube_lits.size, 1), dtype=np.int32)
            keep[:cube_lits.size] = cube_lits
            lp, op = keep.ctypes.data, cube_off.ctypes.data
        hi = n if hi is None else hi
        flag = np.full(n, 255, dtype=np.uint8)
        props = np.zeros(n, dtype=np.uint32)
        off = np.zeros(n + 1, dtype=np.uint64)
        if self.status:
            return flag, props, off, np.zeros(0, dtype=np.int32)
        ptr = C.c_void_p()
        self.L.orc_run_cubes(self.h, lp, op, n, lo, hi, stride, flag.ctypes.data, props.ctypes.data, off.ctypes.data, C.byref(ptr))
        nl = int(off[n])
        lits = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), shape=(max(nl, 1),))[:nl].copy() if ptr.value else np.zeros(0, np.int32)
        self.L.orc_free(ptr)
        return flag, props, off, lits


def naive_up(words, cube):
    """Route C: scan-all-clauses closure, an independent third opinion for tiny cases.
    Returns (conflict, set of true literals)."""
    clauses, cur = [], []
    for w in words.tolist():
        if w == 0:
            clauses.append(cur)
            cur = []
        else:
            cur.append(w)
    true = set()
    for l in cube:
        if -l in true:
            return True, true
        true.add(l)
    changed = True
    while changed:
        changed = False
        for c in clauses:
            if len(c) == 0:
                return True, true
            if any(l in true for l in c):
                continue
            free = [l for l in set(c) if -l not in true]
            if any(-l in free for l in free):
                continue  # tautology
            if not free:
                return True, true
            if len(free) == 1:
                true.add(free[0])
                changed = True
    return False, true
