#!/usr/bin/env python
"""Exploration harness (not the official bench): run probe passes on a generated instance and print the
engine's per-batch info. usage: explore.py <kind> key=value... [--chunk N] [--limit N] [--noimpl] [--trail T] [--check]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from intel_sat_solver_b200 import BatchedTopor, fileio  # noqa: E402
import gpu_util  # noqa: E402


def main():
    kind = sys.argv[1]
    kw = dict(a.split("=") for a in sys.argv[2:] if "=" in a and not a.startswith("--"))
    flags = [a for a in sys.argv[2:] if a.startswith("--")]
    def opt(name, default):
        for f in flags:
            if f.startswith(f"--{name}="):
                return int(f.split("=")[1])
        return default
    d = tempfile.mkdtemp()
    inst, info = gpu_util.gen(kind, d, **kw)
    print("instance", info, flush=True)
    t0 = time.perf_counter()
    t = BatchedTopor()
    t.AddClauses(fileio.read_bcnf(inst))
    t.Finalize()
    print("finalize_s", round(time.perf_counter() - t0, 2), t.DbInfo(), flush=True)
    n = 2 * t.max_var
    limit = opt("limit", n)
    chunk = opt("chunk", limit)
    implied = "--noimpl" not in flags
    trail = opt("trail", 0)
    for rep in range(opt("reps", 2)):
        tot_ms = 0.0
        agg = {}
        for a in range(0, limit, chunk):
            r = t.ProbeRange(a, min(a + chunk, limit), implied=implied, fetch=False, small_trail=trail, mid_trail=opt("mid", 0), kernel_times="--notimes" not in flags)
            tot_ms += r.info["total_ms"]
            for k, v in r.info.items():
                agg[k] = agg.get(k, 0) + v
        agg["probes_per_s"] = limit / (tot_ms / 1e3)
        agg["props_per_s"] = agg["props_ref"] / (tot_ms / 1e3)
        agg["GBps_kernels"] = agg["kernel_bytes"] / (max(agg["kernel_ms"], 1e-9) / 1e3) / 1e9
        print(f"rep{rep}", json.dumps(agg), flush=True)
    if "--check" in flags:
        stride = opt("stride", 1)
        ref, rinfo = gpu_util.run_checker(inst, extra=["--stride", str(stride)])
        r = t.ProbeRange(0, n)
        sel = np.nonzero(ref.flag != 255)[0]
        ok = np.array_equal(r.flag[sel], ref.flag[sel])
        sizes = np.diff(r.impl_off)
        rs = np.diff(ref.off)
        ok2 = np.array_equal(sizes[sel], rs[sel])
        ok3 = all(np.array_equal(r.implied(i), ref.implied(i)) for i in sel[:: max(1, len(sel) // 2000)])
        print("check", ok, ok2, ok3, "cpu_seconds", rinfo["seconds"], "cpu_probes_per_s", rinfo["ran"] / max(rinfo["seconds"], 1e-9), flush=True)


if __name__ == "__main__":
    main()
