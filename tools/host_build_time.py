#!/usr/bin/env python
"""Times the host-side database build (normalise + flatten; no GPU involved) on generated instances.
usage: host_build_time.py [threads ...]   (sets BBCP_HOST_THREADS for each run)"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from intel_sat_solver_b200 import BatchedTopor, fileio  # noqa: E402


def main():
    gen = os.path.join(ROOT, "tools", "cnfgen")
    if not os.path.exists(gen):
        subprocess.run(["gcc", "-O2", "-o", gen, os.path.join(ROOT, "tools", "cnfgen.c"), "-lm"], check=True)
    d = tempfile.mkdtemp()
    insts = []
    for kind, args in (("3sat", ["n=1000000", "m=4200000", "seed=1"]), ("plaw", ["n=2500000", "m=10000000", "seed=5"])):
        f = os.path.join(d, kind + ".bcnf")
        subprocess.run([gen, kind] + args + [f"out={f}"], check=True, capture_output=True)
        insts.append((kind, fileio.read_bcnf(f)))
    for th in sys.argv[1:] or ["1", "0"]:
        if th == "0":
            os.environ.pop("BBCP_HOST_THREADS", None)
        else:
            os.environ["BBCP_HOST_THREADS"] = th
        for kind, w in insts:
            best = 1e9
            for _ in range(2):
                t = BatchedTopor()
                t.AddClauses(w)
                t0 = time.perf_counter()
                t.HostPrepare()
                best = min(best, time.perf_counter() - t0)
            print(f"threads={th or 'all'} {kind}: host prepare {best:.2f} s", flush=True)


if __name__ == "__main__":
    main()
