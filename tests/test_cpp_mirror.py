"""The C++ host mirror of the reference facade (include/bbcp_topor.hpp): compiles as C++17 and C++20, links
against libbbcp.so, refuses to run without a device (CPU test), and reproduces the golden vectors when driven
the way a CTopor user drives the reference (GPU test)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "topor_mirror_main.cpp")
LIBDIR = os.path.join(ROOT, "intel_sat_solver_b200")


@pytest.fixture(scope="module")
def mirror_exe(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libbbcp.so")):
        subprocess.run(["make", "-C", os.path.join(LIBDIR, "csrc"), "-j4"], check=True, capture_output=True)
    exe = str(tmp_path_factory.mktemp("cpp") / "topor_mirror")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-L", LIBDIR, "-lbbcp",
                    f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True)
    return exe


def test_header_is_valid_cxx20_too():
    subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC], check=True)


def test_mirror_fails_loudly_without_device(mirror_exe, golden, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    w = tmp_path / "w.i32"
    golden["regr_72/words"].astype(np.int32).tofile(w)
    r = subprocess.run([mirror_exe, str(w), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "engine error -2" in r.stderr     # BBCP_ERR_CUDA: no CPU path exists


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["regr_72", "regr_47", "fuzz_10", "hand_long"])
def test_mirror_reproduces_goldens(mirror_exe, golden, tmp_path, name):
    w, o = tmp_path / "w.i32", tmp_path / "o.bin"
    golden[f"{name}/words"].astype(np.int32).tofile(w)
    r = subprocess.run([mirror_exe, str(w), str(o)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n, nl = np.fromfile(o, dtype=np.uint64, count=2)
    n, nl = int(n), int(nl)
    flag = np.fromfile(o, dtype=np.uint8, offset=16, count=n)
    off = np.fromfile(o, dtype=np.uint64, offset=16 + n, count=n + 1)
    lits = np.fromfile(o, dtype=np.int32, offset=16 + n + 8 * (n + 1), count=nl)
    assert np.array_equal(flag, golden[f"{name}/probe_flag"])
    assert np.array_equal(off, golden[f"{name}/probe_off"])
    assert np.array_equal(lits, golden[f"{name}/probe_lits"])
