#!/usr/bin/env python
"""Supplementary measurements on the other BASELINE.json configurations (C3 Tseitin/BMC, C4 cube batch, C5
power-law), 1 GPU. NOT the official bench (bench.py = C2): same engine calls, same timing rules (CUDA events on
the engine's stream, L2 flushed between steps, per-kernel times in a separate pass), one JSON line per workload,
with the reference's CPU path (oracle/_ref/ref_bcp, one process per core) timed beside it on a strided sample of
the same probe list -- and the sample's flags and implied-set sizes compared with the engine's while we are at it.

    python tools/bench_workloads.py [c3] [c4] [c5] [--scale F] [--steps K] [--cpu-seconds S]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from intel_sat_solver_b200 import BatchedTopor, fileio  # noqa: E402
import gpu_util  # noqa: E402


def workloads(scale):
    s = lambda x: max(1, int(x * scale))
    return {
        "c3": dict(kind="aig", name=f"C3 Tseitin/BMC AIG, {s(50)} frames x 66,667 gates (~{s(50) * 0.2:.1f} M clauses), seed 2, probe all literals",
                   gen=dict(frames=s(50), gates=66667, inputs=2000, latches=2000, window=5000, seed=2)),
        "c4": dict(kind="comm", name=f"C4 community CNF n={s(1250000)} m={s(5000000)} in {s(1000)} communities, seed 3; 65,536 cubes of depth 20",
                   gen=dict(n=s(1250000), m=s(5000000), comms=s(1000), seed=3), cubes=dict(ncubes=65536, split=16, extra=4)),
        "c5": dict(kind="plaw", name=f"C5 power-law CNF n={s(12500000)} m={s(50000000)} alpha 0.8, seed 5, probe all literals",
                   gen=dict(n=s(12500000), m=s(50000000), seed=5)),
    }


def run_cpu(inst, cubes, stride, cores):
    binary = gpu_util.checker_binary()
    outs, procs = [], []
    d = tempfile.mkdtemp()
    for i in range(cores):
        out = os.path.join(d, f"o{i}.bres")
        cmd = [binary, inst, "--out", out, "--shard", str(i), str(cores), "--stride", str(stride)] + (["--cubes", cubes] if cubes else ["--probe-all"])
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True), out))
    infos = [json.loads(p.communicate()[0]) for p, _ in procs]
    res = [fileio.read_bres(o) for _, o in procs]
    return infos, res, "reference" if binary.endswith("ref_bcp") else "port"


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for key in args.which:
        w = workloads(args.scale)[key]
        d = tempfile.mkdtemp()
        gkw = dict(w["gen"])
        cubes = None
        if "cubes" in w:
            cubes = os.path.join(d, "c.cubes")
            gkw.update(w["cubes"], cubes=cubes)
        t0 = time.perf_counter()
        inst, ginfo = gpu_util.gen(w["kind"], d, **gkw)
        t1 = time.perf_counter()
        eng = BatchedTopor()
        eng.AddClauses(fileio.read_bcnf(inst))
        rc = eng.Finalize()
        t2 = time.perf_counter()
        assert rc == 0, "contradictory at level 0"
        dbi = eng.DbInfo()
        if cubes:
            hdr = np.fromfile(cubes, dtype=np.uint64, count=3)
            nc, nl = int(hdr[1]), int(hdr[2])
            coff = np.fromfile(cubes, dtype=np.uint64, offset=24, count=nc + 1)
            clits = np.fromfile(cubes, dtype=np.int32, offset=24 + 8 * (nc + 1), count=nl)
            n_items = nc
            step = lambda **kw: eng.PropagateBatch(clits, coff, fetch=False, **kw)
        else:
            n_items = 2 * eng.max_var
            step = lambda **kw: eng.ProbeAll(fetch=False, **kw)
        ms, infos = [], []
        for it in range(2 + args.steps):
            flush.fill_(it)
            torch.cuda.synchronize()
            r = step()
            if it >= 2:
                ms.append(r.info["total_ms"])
        flush.fill_(7)
        torch.cuda.synchronize()
        kt = step(kernel_times=True).info
        res = step()    # results kept on the device for the parity sample below
        info = res.info
        sec = sum(ms) / len(ms) / 1e3
        kern_ms = {k: kt[k] for k in ("triage_ms", "deep_ms", "mid_ms", "block_ms", "compact_ms")}
        prop_bytes = info["kernel_bytes"] - info["triage_bytes"]
        prop_ms = kt["deep_ms"] + kt["block_ms"]
        line = {
            "metric": "probes/sec, batched BCP to fixpoint" if not cubes else "cubes/sec, batched BCP to fixpoint", "workload": w["name"],
            "value": n_items / sec, "unit": "probes/s" if not cubes else "cubes/s", "props_per_sec": info["props_ref"] / sec, "n_gpus": 1,
            "steps": args.steps, "ms_per_step": 1e3 * sec, "n_items": n_items, "n_ok": info["n_ok"], "n_conflict": info["n_conflict"],
            "n_skipped": info["n_skipped"], "n_implied": info["n_implied"], "tiers": {"queued": info["n_deep"], "second": info["n_mid"], "third": info["n_large"]},
            "db": {k: dbi[k] for k in ("n_bin", "n_tern", "n_long", "row_slots", "device_bytes")}, "gen_s": round(t1 - t0, 1), "db_build_s": round(t2 - t1, 1),
            "kernel_ms": kern_ms, "gpu_launches": info["launches"],
            "reuse": {k: info[k] for k in ("adopted", "adopt_lits", "inherited")}, "props_all": info["props_all"], "slots": info["slots"],
            "roofline": {"bound": "hbm", "kernel": "k_deep + k_deep2 + k_big (propagation)", "bytes_per_launch": prop_bytes, "kernel_ms": prop_ms,
                         "achieved": prop_bytes / max(prop_ms, 1e-9) / 1e6, "peak": peak, "unit": "GB/s", "frac": prop_bytes / max(prop_ms, 1e-9) / 1e6 / peak,
                         "note": "engine byte counters (row headers, slots, clause units, stored literals); served mostly from L2"},
            "l2": "flushed between timed steps (256 MiB write)", "dtype": "u32", "data": "synthetic",
        }
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            # stride sized from a first guess of the CPU rate so that the sample takes about --cpu-seconds per core
            probe_rate_guess = {"c3": 2e3, "c4": 2e4, "c5": 3e5}[key]   # items/s per core, order of magnitude
            stride = max(1, int(n_items / (probe_rate_guess * args.cpu_seconds * cores)))
            cinfo, cres, kind = run_cpu(inst, cubes, stride, cores)
            ran = sum(i["ran"] for i in cinfo)
            tmax = max(i["seconds"] for i in cinfo)
            line["cpu_baseline"] = {"value": ran / tmax, "unit": line["unit"], "cores": cores, "kind": kind,
                                    "sample": f"every {stride}-th item ({ran} items), one process per core over contiguous shards, CNF load untimed",
                                    "props_per_sec": sum(i["implications"] for i in cinfo) / tmax}
            # parity of the sample against the engine's device-resident results
            flag = np.empty(n_items, dtype=np.uint8)
            off = np.empty(n_items + 1, dtype=np.uint64)
            eng.lib.bbcp_fetch_flags(eng.handle, flag.ctypes.data)
            lits = np.empty(max(info["n_implied"], 1), dtype=np.int32)
            eng.lib.bbcp_fetch_implied(eng.handle, off.ctypes.data, lits.ctypes.data)
            sizes = np.diff(off)
            checked = bad = 0
            for r in cres:
                sel = np.nonzero(r.flag != 255)[0]
                rs = np.diff(r.off)
                bad += int((flag[sel] != r.flag[sel]).sum()) + int((sizes[sel] != rs[sel]).sum())
                for i in sel[:: max(1, len(sel) // 200)]:
                    bad += int(not np.array_equal(lits[int(off[i]):int(off[i + 1])], r.lits[int(r.off[i]):int(r.off[i + 1])]))
                checked += len(sel)
            line["parity_sample"] = {"items_checked": checked, "mismatches": bad}
        print(json.dumps(line), flush=True)
        del eng


if __name__ == "__main__":
    main()
