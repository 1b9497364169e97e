#!/usr/bin/env python
"""bench.py -- batched BCP to fixpoint: probes/s (and propagations/s) on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workloads c3,c4,c5|none]

One "step" = one all-literal failed-literal-probing pass over the instance: every literal of the probe universe
is propagated to its unit-propagation fixpoint over the shared clause database; per probe a conflict flag and
the implied-literal set come back, plus the failed-literal / unit bitmaps.

Headline workload (config.workload): C2 of SURVEY.md 8d -- random 3-SAT, 1,000,000 variables, clause/variable ratio
4.2, seed 1 (tools/cnfgen.c), probe all 2,000,000 literals on one B200. With N GPUs the headline is a WEAK-scaling
run: N x 1M variables / N x 4.2M clauses, the database replicated on every GPU, the probe universe cut into N
contiguous literal ranges (one per rank, aligned to failed-literal-bitmap words), and the all-gather of the bitmap
slices done by the engine itself over NVLink peer memory inside the batch's last kernel (BBCP_OPT_MERGE).

Keys of the JSON line:
  value        probes/s with everything resident in HBM: CUDA events on the engine's stream from the first kernel of
               the batch to the end of its last one (incl. the merge), summed over the K timed steps, max over ranks.
               Probe-list upload and result read-back are NOT inside (they are what `e2e` adds).
  e2e          the same metric through ONE C-ABI call with HOST buffers (pinned probe list in; flags, offsets and
               implied literals out), wall clock, max over ranks.
  roofline     the kernel of the C2 step with the longest device time on the engine's own bytes model -- C2 has no
               binary clause, so the step is a table look-up + result compaction, NOT a row scan: the line says so
               (`abytes_2wl_frac_of_step` > 1 is what that looks like against SURVEY 8d's two-watched-literal bytes).
               The propagation kernels' roofline is in `workloads`.
  workloads    (N = 1) BASELINE configs C3 (Tseitin/BMC, 10 M clauses), C4 (65,536 cubes x depth 20 on a 5 M-clause
               community CNF) and C5 (50 M-clause power-law) at full size, each timed the same way (own CUDA events,
               L2 flushed before every step, own clock samples), with the reference's CPU path on a strided sample
               of the same items (`cpu_baseline`), a bit-exact comparison of that sample with the engine's results
               (`parity_sample`), and SURVEY 8d's ABYTES roofline from the checker's fixpoints (oracle/abytes).
               (N > 1) C5 at full size, probe ranges sharded over the N ranks with the in-kernel bitmap merge.
  merge_check  (N > 1) sharded passes with BBCP_OPT_MERGE on instances that HAVE failed literals (the committed
               goldens fuzz_10 / fuzz_5 / regr_47 and a 200 k-variable power-law instance): every rank's merged
               bitmap against the reference-derived golden bitmap / against one rank's unsharded pass / against an
               NCCL all-gather of the same slices.
  cpu_baseline the reference's own CTopi::BCP() (oracle/_ref/ref_bcp; the C port if the harness did not travel) on
               the box's host cores, one process per core over probe shards, timed in this same run.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CNFGEN = os.path.join(ROOT, "tools", "cnfgen")
REF_BCP = os.path.join(ROOT, "oracle", "_ref", "ref_bcp")
PORT_BCP = os.path.join(ROOT, "oracle", "oracle_bcp")
ABYTES = os.path.join(ROOT, "oracle", "abytes")
METRIC = "probes/sec, batched BCP to fixpoint (all-literal failed-literal probing)"
TMP = os.path.join(tempfile.gettempdir(), "bbcp_bench")


def ensure_tools():
    if not os.path.exists(CNFGEN):
        subprocess.run(["gcc", "-O2", "-o", CNFGEN, os.path.join(ROOT, "tools", "cnfgen.c"), "-lm"], check=True)


def generate(kind, tag, **kw):
    """Deterministic synthetic instance (tools/cnfgen.c); cached under the temp dir, written atomically."""
    ensure_tools()
    os.makedirs(TMP, exist_ok=True)
    name = kind + "_" + "_".join(f"{k}{v}" for k, v in sorted(kw.items()) if k != "cubes") + f"_{tag}"
    path = os.path.join(TMP, name + ".bcnf")
    cubes = os.path.join(TMP, name + ".cubes") if "ncubes" in kw else None
    if not os.path.exists(path):
        tmp = path + f".{os.getpid()}.tmp"
        args = [CNFGEN, kind] + [f"{k}={v}" for k, v in kw.items()] + [f"out={tmp}"]
        if cubes:
            args.append(f"cubes={cubes}.{os.getpid()}.tmp")
        subprocess.run(args, check=True, capture_output=True)
        if cubes:
            os.replace(f"{cubes}.{os.getpid()}.tmp", cubes)
        os.replace(tmp, path)
    return path, cubes


def make_instance(nvars, nclauses, seed, tag):
    return generate("3sat", tag, n=nvars, m=nclauses, seed=seed)[0]


def workload_name(nvars, nclauses, seed, n_gpus):
    w = f"C2 random 3-SAT n={nvars} m={nclauses} seed={seed}, probe all {2 * nvars} literals"
    if n_gpus > 1:
        w += f" (weak scaling: {n_gpus} x the 1-GPU instance, DB replicated, probe ranges sharded)"
    return w


def config_for(nvars, nclauses, seed, n_gpus):
    """The same dict in both arms (the driver compares them)."""
    return {"workload": workload_name(nvars, nclauses, seed, n_gpus), "instance": {"kind": "3sat", "n": nvars, "m": nclauses, "seed": seed},
            "probes_per_step": 2 * nvars, "results": "flag + sorted implied-literal set per probe, failed/unit bitmaps"}


class ClockSampler:
    """ONE nvidia-smi process (rank 0 only) sampling SM clocks / throttle reasons of the GPUs in use."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, indices):
        self.indices, self.rows, self.proc, self.marks = set(indices), [], None, {}

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) >= 8 and r[0].isdigit() and int(r[0]) in self.indices:
                self.rows.append(r)

    def mark(self):
        return len(self.rows)

    def wait_first(self, timeout=4.0):
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def under_load(self, fn, want=3, limit=2.0):
        """The timed regions here are micro- to milliseconds long, nvidia-smi samples every 200 ms: keep the SAME step running
        (untimed) until a few samples have been taken under that load."""
        m0, t0 = self.mark(), time.perf_counter()
        while self.proc and self.mark() - m0 < want and time.perf_counter() - t0 < limit:
            fn()
        return self.summary(m0)

    def summary(self, lo=0, hi=None):
        rows = self.rows[lo:hi]
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        return self.summary()


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_binary():
    if os.path.exists(REF_BCP):
        return REF_BCP, "reference"
    if not os.path.exists(PORT_BCP):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
    return PORT_BCP, "port"


def start_cpu_shards(inst, cores, stride=1, repeat=1, cubes=None, out_dir=None):
    """One checker process per host core over contiguous item shards; each loads the CNF (untimed) and times only
    its propagation loop. Returns the running processes (collect with finish_cpu_shards)."""
    binary, kind = cpu_binary()
    procs = []
    for i in range(cores):
        cmd = [binary, inst] + (["--cubes", cubes] if cubes else ["--probe-all"]) + ["--shard", str(i), str(cores), "--stride", str(stride), "--repeat", str(repeat)]
        out = None
        if out_dir:
            out = os.path.join(out_dir, f"shard{i}.bres")
            cmd += ["--out", out]
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True), out))
    return procs, kind


def finish_cpu_shards(procs, kind, repeat=1):
    outs = [json.loads(p.communicate()[0]) for p, _ in procs]
    passes = [max(o["seconds_list"][r] for o in outs) for r in range(repeat)]
    return {"kind": kind, "cores": len(procs), "ran": sum(o["ran"] for o in outs), "pass_seconds": passes,
            "implications_per_pass": sum(o["implications"] for o in outs) / repeat, "conflicts": sum(o["conflicts"] for o in outs),
            "assignments_per_pass": sum(o["assignments"] for o in outs) / repeat, "files": [f for _, f in procs]}


def run_cpu_shards(inst, repeat, cores=None, stride=1):
    procs, kind = start_cpu_shards(inst, cores or os.cpu_count() or 1, stride, repeat)
    return finish_cpu_shards(procs, kind, repeat)


def reference_arm(args, rank, world):
    if rank != 0:
        return
    nvars, ncls = args.vars * world, int(args.vars * world * args.ratio)
    inst = make_instance(nvars, ncls, args.seed, "ref")
    r = run_cpu_shards(inst, args.warmup + args.steps)
    timed = r["pass_seconds"][args.warmup:]
    total = sum(timed)
    value = r["ran"] * len(timed) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "probes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": config_for(nvars, ncls, args.seed, world),
        "props_per_sec": r["implications_per_pass"] * len(timed) / total,
        "cpu_baseline": {"value": value, "unit": "probes/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": f"all {r['ran']} probes per step, one process per core over contiguous shards, CNF load untimed"},
        "e2e": {"value": value, "unit": "probes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm: helpers
def stats(xs):
    xs = sorted(xs)
    return {"min": xs[0], "median": xs[len(xs) // 2], "max": xs[-1]} if xs else None


def peak_hbm():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def timed_passes(eng, call, steps, warmup=2):
    """`call(opts_kw)` runs one pass and returns its BatchResult. L2 is flushed ON THE ENGINE'S STREAM before every
    pass (bbcp_l2_flush), so the pass's kernels are already queued when the device gets to them."""
    ms, last = [], None
    for it in range(warmup + steps):
        eng.lib.bbcp_l2_flush(eng.handle, 0)
        last = call({})
        if it >= warmup:
            ms.append(last.info["total_ms"])
    eng.lib.bbcp_l2_flush(eng.handle, 0)
    kt = call({"kernel_times": True}).info
    return ms, last, kt


WORKLOADS = {
    "c3": dict(kind="aig", name="C3 Tseitin/BMC AIG, 50 frames x 66,667 gates (10.0 M clauses), seed 2, probe all literals",
               gen=dict(frames=50, gates=66667, inputs=2000, latches=2000, window=5000, seed=2), rate_guess=2e3),
    "c4": dict(kind="comm", name="C4 community CNF n=1,250,000 m=5,000,000 in 1,000 communities, seed 3; 65,536 cubes of depth 20",
               gen=dict(n=1250000, m=5000000, comms=1000, seed=3, ncubes=65536, split=16, extra=4), rate_guess=4e4),
    "c5": dict(kind="plaw", name="C5 power-law CNF n=12,500,000 m=50,000,000 alpha 0.8, seed 5, probe all literals",
               gen=dict(n=12500000, m=50000000, seed=5), rate_guess=5e4),
}


def read_cubes(path):
    hdr = np.fromfile(path, dtype=np.uint64, count=3)
    nc, nl = int(hdr[1]), int(hdr[2])
    return np.fromfile(path, dtype=np.int32, offset=24 + 8 * (nc + 1), count=nl), np.fromfile(path, dtype=np.uint64, offset=24, count=nc + 1)


def workload_block(key, scale_cpu_seconds, clocks, cores, pre=None):
    """One BASELINE configuration at full size on one GPU. `pre` = CPU checker processes already started for it."""
    from intel_sat_solver_b200 import BatchedTopor, fileio
    w = WORKLOADS[key]
    t0 = time.perf_counter()
    inst, cubes = generate(w["kind"], "wl", **w["gen"])
    t1 = time.perf_counter()
    eng = BatchedTopor()
    eng.AddClauses(fileio.read_bcnf(inst))
    rc = eng.Finalize()
    t2 = time.perf_counter()
    assert rc == 0, "contradictory at level 0"
    dbi = eng.DbInfo()
    if cubes:
        clits, coff = read_cubes(cubes)
        n_items = coff.size - 1
        call = lambda kw: eng.PropagateBatch(clits, coff, fetch=False, **kw)
    else:
        n_items = 2 * eng.max_var
        call = lambda kw: eng.ProbeAll(fetch=False, **kw)
    mark = clocks.mark() if clocks else 0
    ms, res, kt = timed_passes(eng, call, steps=2)
    clk = clocks.summary(mark) if clocks else None
    if clocks and (not clk or clk["samples"] < 2):
        clk = clocks.under_load(lambda: call({}))     # a pass of a few milliseconds: keep it running for a few samples
    res = call({})     # results kept on the device for the parity sample
    info = res.info
    sec = sum(ms) / len(ms) / 1e3
    peak, peak_src = peak_hbm()
    prop_ms = kt["deep_ms"] + kt["block_ms"]
    prop_bytes = info["kernel_bytes"] - info["triage_bytes"]
    unit = "cubes/s" if cubes else "probes/s"
    line = {"workload": w["name"], "value": n_items / sec, "unit": unit, "props_per_sec": info["props_ref"] / sec, "ms_per_step": 1e3 * sec,
            "steps": len(ms), "step_ms": ms, "n_items": n_items, "n_ok": info["n_ok"], "n_conflict": info["n_conflict"], "n_skipped": info["n_skipped"],
            "n_implied": info["n_implied"], "tiers": {"queued": info["n_deep"], "second_tier": info["n_mid"], "last_tier": info["n_large"]},
            "closure_reuse": {k: info[k] for k in ("adopted", "adopt_lits", "inherited")},
            "counters": {k: info[k] for k in ("props_all", "props_ref", "assignments", "bcps", "slots", "clause_fetches")},
            "db": {k: dbi[k] for k in ("n_bin", "n_tern", "n_long", "row_slots", "device_bytes")}, "gen_s": round(t1 - t0, 1), "db_build_s": round(t2 - t1, 1),
            "kernel_ms": {k: kt[k] for k in ("triage_ms", "deep_ms", "mid_ms", "block_ms", "compact_ms")}, "gpu_launches": info["launches"],
            "l2": "flushed before every timed step (256 MiB write on the engine's stream)", "clocks": clk}
    if pre is not None or scale_cpu_seconds > 0:
        with tempfile.TemporaryDirectory() as d:
            if pre is None:
                stride = max(1, int(n_items / (w["rate_guess"] * scale_cpu_seconds * cores)))
                procs, kind = start_cpu_shards(inst, cores, stride, 1, cubes, d)
            else:
                procs, kind, stride, d2 = pre
            r = finish_cpu_shards(procs, kind)
            tmax = r["pass_seconds"][0]
            line["cpu_baseline"] = {"value": r["ran"] / tmax, "unit": unit, "cores": r["cores"], "kind": kind,
                                    "sample": f"every {stride}-th item ({r['ran']} items), one process per core over contiguous shards, CNF load untimed",
                                    "props_per_sec": r["implications_per_pass"] / tmax}
            # parity of the sample against the engine's device-resident results: flags and set sizes of every sampled item,
            # the literal sets themselves on ~200 items per shard
            flag = np.empty(n_items, dtype=np.uint8)
            off = np.empty(n_items + 1, dtype=np.uint64)
            eng.lib.bbcp_fetch_flags(eng.handle, flag.ctypes.data)
            lits = np.empty(max(info["n_implied"], 1), dtype=np.int32)
            eng.lib.bbcp_fetch_implied(eng.handle, off.ctypes.data, lits.ctypes.data)
            sizes = np.diff(off)
            checked = bad = sets = 0
            for f in r["files"]:
                cr = fileio.read_bres(f)
                sel = np.nonzero(cr.flag != 255)[0]
                rs = np.diff(cr.off)
                bad += int((flag[sel] != cr.flag[sel]).sum()) + int((sizes[sel] != rs[sel]).sum())
                for i in sel[:: max(1, len(sel) // 200)]:
                    bad += int(not np.array_equal(lits[int(off[i]):int(off[i + 1])], cr.lits[int(cr.off[i]):int(cr.off[i + 1])]))
                    sets += 1
                checked += len(sel)
            line["parity_sample"] = {"items_checked": checked, "sets_compared": sets, "mismatches": bad}
            # SURVEY 8d's ABYTES on the checker's fixpoints of the sample, scaled to the whole batch by the sampling stride
            try:
                if not os.path.exists(ABYTES):
                    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "abytes"], check=True, capture_output=True)
                a = json.loads(subprocess.run([ABYTES, inst] + (["--cubes", cubes] if cubes else []) + r["files"], check=True, capture_output=True, text=True).stdout)
                abytes = a["abytes_total"] * (n_items / max(a["items"], 1))
                line["roofline"] = {"bound": "hbm", "kernel": "whole step (triage + propagation tiers + compaction)", "achieved": abytes / sec / 1e9, "peak": peak,
                                    "unit": "GB/s", "frac": abytes / sec / 1e9 / peak, "peak_source": peak_src,
                                    "abytes_per_launch": abytes, "abytes_basis": f"oracle/abytes on the {a['items']} sampled items x {n_items / max(a['items'], 1):.1f}",
                                    "engine_bytes_per_launch": prop_bytes, "engine_bytes_frac_of_propagation": prop_bytes / max(prop_ms, 1e-9) / 1e6 / peak,
                                    "traffic": load_traffic(key)}
            except (OSError, subprocess.CalledProcessError, ValueError) as e:
                line["roofline"] = {"error": f"abytes tool failed: {e}"}
            if pre is not None:
                d2.cleanup()
    del eng
    return line


def load_traffic(key):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key)
    except (OSError, ValueError):
        return None


# ------------------------------------------------------------------------------------------- GPU arm
def merge_check(rank, world, dist, torch):
    """Sharded all-literal passes with the in-batch merge on instances that have failed literals."""
    from intel_sat_solver_b200 import BatchedTopor, fileio, sharding
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    bad = inst = failed_bits = 0
    cases = [(n, golden[f"{n}/words"], golden[f"{n}/probe_flag"]) for n in ("fuzz_10", "fuzz_5", "regr_47")]
    path, _ = generate("plaw", "mc", n=200000, m=800000, seed=5) if rank == 0 else (None, None)
    if world > 1:
        dist.barrier()
    path, _ = generate("plaw", "mc", n=200000, m=800000, seed=5)
    cases.append(("plaw_200k", fileio.read_bcnf(path), None))
    for name, words, gflag in cases:
        t = BatchedTopor(device=torch.cuda.current_device())
        t.AddClauses(words)
        assert t.Finalize() == 0
        s = sharding.shard_for(rank, world, t.max_var)
        sharding.peer_attach(t, rank, world)
        res = t.ProbeRange(s.first, s.last, merge=s.words_per_rank, rendezvous=True)
        failed, unit = t.Bitmaps()
        nwords = sharding.bitmap_words(t.max_var)
        if gflag is not None:      # the reference's own failed-literal set
            bits = np.zeros(nwords * 32, dtype=np.uint8)
            bits[np.nonzero(gflag == 1)[0] + 2] = 1
            expect = np.packbits(bits, bitorder="little").view(np.uint32)
            bad += int(not np.array_equal(res.flag, gflag[s.first:s.last]))
        else:                      # one rank's unsharded pass, broadcast
            t1 = BatchedTopor(device=torch.cuda.current_device())
            t1.AddClauses(words)
            t1.Finalize()
            t1.ProbeAll(implied=False, fetch=False)
            e = torch.from_numpy(t1.Bitmaps()[0].view(np.int32).copy()).cuda()
            if world > 1:
                dist.broadcast(e, 0)
            expect = e.cpu().numpy().view(np.uint32)
            del t1
        bad += int(not np.array_equal(failed, expect))
        ub = np.unpackbits(unit.view(np.uint8), bitorder="little")
        fb = np.unpackbits(expect.view(np.uint8), bitorder="little")
        bad += int(not np.array_equal(np.nonzero(ub)[0], np.sort(np.nonzero(fb)[0] ^ 1)))
        local = torch.from_numpy(failed.view(np.int32).copy()).cuda()
        bad += int(not np.array_equal(sharding.allgather_bitmap(local, s).cpu().numpy().view(np.uint32), expect))
        inst += 1
        failed_bits += int(fb.sum())
        t.lib.bbcp_peer_detach(t.handle)
        if world > 1:
            dist.barrier()
        del t
    tot = torch.tensor([bad], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    return {"instances": inst, "world": world, "failed_literals_in_expected_bitmaps": failed_bits, "mismatches": int(tot.item())}


def c5_sharded(rank, world, dist, torch, clocks):
    """BASELINE config 5 at full size: 50 M clauses, 25 M probes, contiguous probe ranges over the ranks, bitmap merge in
    the batch's last kernel (strong scaling: the same job on N GPUs)."""
    from intel_sat_solver_b200 import BatchedTopor, fileio, sharding
    w = WORKLOADS["c5"]
    if rank == 0:
        generate(w["kind"], "wl", **w["gen"])
    dist.barrier()
    inst, _ = generate(w["kind"], "wl", **w["gen"])
    t0 = time.perf_counter()
    eng = BatchedTopor(device=torch.cuda.current_device())
    eng.AddClauses(fileio.read_bcnf(inst))
    assert eng.Finalize() == 0
    build_s = time.perf_counter() - t0
    s = sharding.shard_for(rank, world, eng.max_var)
    sharding.peer_attach(eng, rank, world)
    mark = clocks.mark() if clocks else 0
    ms = []
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        eng.lib.bbcp_l2_flush(eng.handle, 0)
        r = eng.ProbeRange(s.first, s.last, fetch=False, merge=s.words_per_rank, rendezvous=True)
        if it >= 1:
            ms.append(r.info["total_ms"])
    info = r.info
    failed, _ = eng.Bitmaps()
    nfailed = int(np.unpackbits(failed.view(np.uint8)).sum())
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {"rank": rank, "ms": ms, "n": s.last - s.first, "n_conflict": info["n_conflict"], "props_ref": info["props_ref"],
                                      "n_implied": info["n_implied"], "failed_bits_after_merge": nfailed, "db_build_s": round(build_s, 1)})
    eng.lib.bbcp_peer_detach(eng.handle)
    dist.barrier()
    del eng
    if rank != 0:
        return None
    t_step = [max(p["ms"][k] for p in per_rank) for k in range(len(ms))]
    sec = sum(t_step) / len(t_step) / 1e3
    n = sum(p["n"] for p in per_rank)
    return {"workload": w["name"] + f", sharded over {world} GPUs (contiguous literal ranges, DB replicated, in-kernel bitmap all-gather)", "scaling": "strong",
            "value": n / sec, "unit": "probes/s", "ms_per_step": 1e3 * sec, "props_per_sec": sum(p["props_ref"] for p in per_rank) / sec, "n_items": n,
            "failed_literals": sum(p["n_conflict"] for p in per_rank), "merged_bitmap_agrees": len({p["failed_bits_after_merge"] for p in per_rank}) == 1
            and per_rank[0]["failed_bits_after_merge"] == sum(p["n_conflict"] for p in per_rank),
            "per_rank_ms": [stats(p["ms"]) for p in per_rank], "load_imbalance": max(min(p["ms"]) for p in per_rank) / max(1e-9, min(min(p["ms"]) for p in per_rank)),
            "db_build_s_per_rank": [p["db_build_s"] for p in per_rank], "clocks": clocks.summary(mark) if clocks else None}


def ours_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from intel_sat_solver_b200 import BatchedTopor, fileio, sharding
    from intel_sat_solver_b200._lib import BatchInfo, Opts

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cores = os.cpu_count() or 1
    wl_keys = [k for k in args.workloads.split(",") if k in WORKLOADS] if args.workloads != "none" else []

    pre_c5 = None
    nvars, ncls = args.vars * world, int(args.vars * world * args.ratio)
    if rank == 0:
        make_instance(nvars, ncls, args.seed, "c2")
    if world > 1:
        dist.barrier()
    inst = make_instance(nvars, ncls, args.seed, "c2")
    t0 = time.perf_counter()
    words = fileio.read_bcnf(inst)
    eng = BatchedTopor(device=local_rank)
    eng.AddClauses(words)
    rc = eng.Finalize()
    build_s = time.perf_counter() - t0
    assert rc == 0, "instance contradictory at level 0"
    del words
    L, h = eng.lib, eng.handle
    dbi = eng.DbInfo()

    universe = 2 * nvars
    shard = sharding.shard_for(rank, world, nvars)
    assert L.bbcp_bitmap_words(h) == sharding.bitmap_words(nvars)
    first, last = shard.first, shard.last
    n_local = last - first
    # world > 1: BBCP_OPT_MERGE (16) = the bitmap all-gather over NVLink peer memory is the last kernel of the batch;
    # BBCP_OPT_RENDEZVOUS (32) = device-side barrier with all peers right before the timed batch
    mg = (16 | 32) if world > 1 else 0
    opts = Opts(1 | mg, 0, 0, 0, shard.words_per_rank)
    opts_timed = Opts(1 | 8 | mg, 0, 0, 0, shard.words_per_rank)   # + BBCP_OPT_KERNEL_TIMES: per-kernel CUDA events (separate pass)
    info = BatchInfo()
    if world > 1:
        sharding.peer_attach(eng, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = None
    if rank == 0:
        clocks = ClockSampler(range(world))    # ONE sampler for the box (rank 0), not one per rank
        clocks.start()
        clocks.wait_first()
    kstats = {"triage_ms": [], "deep_ms": [], "compact_ms": [], "block_ms": [], "merge_ms": []}
    step_ms = []
    mark_c2 = clocks.mark() if clocks else 0
    for it in range(args.warmup + args.steps):
        barrier()
        # L2 flush on the ENGINE's stream (256 MiB write): the step's kernels queue up behind it, so the host has issued
        # all of them before the device starts the step, and the device-side rendezvous then aligns the ranks
        L.bbcp_l2_flush(h, 0)
        rc = L.bbcp_probe_range(h, first, last, C.byref(opts), C.byref(info))
        assert rc == 0, eng.GetStatusExplanation()
        if it >= args.warmup:
            step_ms.append(info.total_ms)   # CUDA events on the engine's stream: first kernel .. end of the last (merge incl.)
    launches = info.launches * args.steps
    n_conflict, n_implied, props_ref, props_all = info.n_conflict, info.n_implied, info.props_ref, info.props_all
    kernel_bytes, triage_bytes = info.kernel_bytes, info.triage_bytes
    for it in range(args.steps):     # per-kernel durations for the roofline (extra events cost stream time: separate pass)
        barrier()
        L.bbcp_l2_flush(h, 0)
        rc = L.bbcp_probe_range(h, first, last, C.byref(opts_timed), C.byref(info))
        assert rc == 0, eng.GetStatusExplanation()
        for k in kstats:
            kstats[k].append(getattr(info, k))
    barrier()
    # clocks under this load: the timed steps are ~30 us each, so the same step keeps running (untimed, all ranks) until the
    # sampler has a few rows
    for _ in range(30000):     # (a fixed count: with peers every rank must make the same sequence of merging calls)
        rc = L.bbcp_probe_range(h, first, last, C.byref(opts), C.byref(info))
        assert rc == 0, eng.GetStatusExplanation()
    barrier()
    clk = clocks.summary(mark_c2) if clocks else None
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    probes = torch.tensor([float(n_local)], dtype=torch.float64, device=dev)
    props = torch.tensor([float(props_ref)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(probes)
        dist.all_reduce(props)
    total_s = float(total_ms.item()) / 1e3
    value = float(probes.item()) * args.steps / total_s
    per_rank = [None] * world
    mine = {"rank": rank, "step_ms": stats(step_ms), "kernel_ms": {k: stats(v) for k, v in kstats.items()}, "db_build_s": round(build_s, 2)}
    if world > 1:
        dist.all_gather_object(per_rank, mine)
    else:
        per_rank = [mine]

    # ---- e2e: the call a user makes, host buffers in and out, every step
    idx = np.arange(first, last, dtype=np.int64)
    h_lits = torch.from_numpy(np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)).pin_memory()
    o_flag = torch.empty(n_local + 8, dtype=torch.uint8).pin_memory()
    o_off = torch.empty(n_local + 1, dtype=torch.int32).pin_memory()   # u32 CSR offsets (BBCP_OPT_OFF32)
    o_lits = torch.empty(max(int(n_implied) * 2, 1024), dtype=torch.int32).pin_memory()

    def e2e_loop(flags):
        opts_e2e = Opts(flags, 0, 0, 0, 0)
        secs = []
        for it in range(args.warmup + args.steps):
            barrier()
            L.bbcp_l2_flush(h, 0)
            L.bbcp_sync(h)
            t0 = time.perf_counter()
            rc = L.bbcp_probe_lits_to_host(h, h_lits.data_ptr(), n_local, C.byref(opts_e2e), o_flag.data_ptr(), o_off.data_ptr(), o_lits.data_ptr(),
                                           o_lits.numel(), C.byref(info))
            assert rc == 0, eng.GetStatusExplanation()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                secs.append(dt)
        tot = torch.tensor([sum(secs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item())

    # (a) the standard result format: flags, u32 CSR offsets, implied literals (BBCP_OPT_IMPLIED | BBCP_OPT_OFF32)
    full_total = e2e_loop(1 | 4)
    d2h_full = n_local + (n_local + 1) * 4 + int(info.n_implied) * 4
    assert int(o_off[n_local]) == info.n_implied and int((o_flag[:n_local] == 0).sum()) == info.n_ok
    # (b) the compact read-back (+ BBCP_OPT_COMPACT_SELF): a probe that implies only itself comes back as a flag bit; bit-exact
    # after host-side expansion (tests/test_gpu_parity.py::test_compact_read_back_expands_to_the_same_results). Headline e2e.
    e2e_sum = e2e_loop(1 | 4 | 128)
    e2e_value = float(probes.item()) * args.steps / e2e_sum
    h2d = h_lits.numel() * 4
    d2h = n_local + ((n_local + 1) * 4 + int(info.n_implied) * 4 if info.n_implied else 4)
    assert int(((o_flag[:n_local] & 0x7f) == 0).sum()) == info.n_ok and int((o_flag[:n_local] & 0x80).ne(0).sum()) + info.n_implied >= info.n_ok
    if world > 1:
        eng.lib.bbcp_peer_detach(h)
        dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = run_cpu_shards(inst, 2, cores)
        v = r["ran"] / r["pass_seconds"][-1]
        cpu = {"value": v, "unit": "probes/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"all {r['ran']} probes once (after one warm pass), one process per core over contiguous shards, CNF load untimed",
               "props_per_sec": r["implications_per_pass"] / r["pass_seconds"][-1], "failed_literals": r["conflicts"] // 2,
               "implications": r["implications_per_pass"], "engine_props_ref": int(props_ref)}
    del eng

    # the C5 checker processes need minutes to load 50 M clauses: start them now (the C2 CPU leg above had every core),
    # on a few cores, and collect them last; the other workloads' CPU legs leave those cores alone
    if rank == 0 and world == 1 and "c5" in wl_keys and not args.no_cpu:
        w = WORKLOADS["c5"]
        c5_cores = min(4, max(1, cores // 4))
        inst5, _ = generate(w["kind"], "wl", **w["gen"])
        d5 = tempfile.TemporaryDirectory()
        stride5 = max(1, int(25_000_000 / (w["rate_guess"] * 6.0 * c5_cores)))
        procs, kind = start_cpu_shards(inst5, c5_cores, stride5, 1, None, d5.name)
        pre_c5 = (procs, kind, stride5, d5)
        cores = max(1, cores - c5_cores)

    mcheck = merge_check(rank, world, dist, torch) if world > 1 else None
    workloads = {}
    if world == 1:
        for key in wl_keys:
            try:
                workloads[key] = workload_block(key, 0 if args.no_cpu else 8.0, clocks, cores, pre_c5 if key == "c5" else None)
            except Exception as e:   # a workload block must not take the headline line down with it
                workloads[key] = {"error": repr(e)}
    elif "c5" in wl_keys:
        try:
            workloads["c5"] = c5_sharded(rank, world, dist, torch, clocks)
        except Exception as e:
            workloads["c5"] = {"error": repr(e)}

    if rank == 0:
        clocks.stop()
        peak, peak_src = peak_hbm()
        avg = lambda k: sum(kstats[k]) / max(len(kstats[k]), 1) / 1e3
        tri, dee, com, blk = avg("triage_ms"), avg("deep_ms"), avg("compact_ms"), avg("block_ms")
        # engine bytes model, per launch: triage; propagation (engine counters); compaction = 4 B
        # size read + 8 B offset + 4 B literal written per item here (every implied set of this workload is the probe's
        # own literal, regenerated, so nothing is read back)
        prop_bytes = kernel_bytes - triage_bytes
        compact_bytes = n_local * (4 + 8) + int(n_implied) * 4
        cands = [("k_triage (8 probes per thread: one byte of per-literal facts -> flag, size; NOT a row scan)", triage_bytes, tri),
                 ("k_deep (warp-per-probe propagate; empty queue on this workload)", prop_bytes, dee + blk),
                 ("k_scan_gather (result compaction: sizes -> CSR offsets + literals)", compact_bytes, com)]
        kname, kbytes, t1 = max(cands, key=lambda c: c[2])
        achieved = kbytes / t1 / 1e9
        traffic = load_traffic(kname.split(" ")[0])
        # SURVEY.md 8d ABYTES for this instance family in closed form: T={p}, no binaries, every long watch visited:
        # 16 + 24*nlong(not p); sum of nlong over literals = 2m
        abytes_2wl = 16 * universe + 24 * 2 * ncls
        line = {
            "metric": METRIC, "value": value, "unit": "probes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": config_for(nvars, ncls, args.seed, world),
            "value_excludes": "probe-list upload and result read-back (inputs resident; see e2e)",
            "details": {"l2": "flushed before every timed step (256 MiB write on the engine's stream)",
                        "db": {k: dbi[k] for k in ("n_bin", "n_tern", "n_long", "row_slots", "device_bytes")}, "db_build_s": round(build_s, 2),
                        "per_rank": per_rank},
            "props_per_sec": float(props.item()) * args.steps / total_s,
            "failed_literals": int(n_conflict),
            "e2e": {"value": e2e_value, "unit": "probes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_sum / args.steps, "format": "compact read-back (BBCP_OPT_COMPACT_SELF): flags with a self-implied bit; offsets + literals only for stored sets",
                    "transport": "pinned host buffers used in place by bbcp_probe_lits_to_host: the triage kernel reads the probe list over PCIe "
                                 "(h2d bytes) and writes the flags into the caller's buffer (d2h bytes); BBCP_E2E=dma selects copy-engine transfers",
                    "full_format": {"value": float(probes.item()) * args.steps / full_total, "ms_per_step": 1e3 * full_total / args.steps, "d2h_bytes_per_step": d2h_full,
                                    "format": "flags + u32 CSR offsets + implied literals"}},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "bytes_model": "engine bytes (DESIGN.md section 4), not SURVEY 8d ABYTES: see note",
                         "note": "C2 has no binary clause: every probe is decided from one byte of per-literal facts (length pruning), so this step times "
                                 "triage + result compaction and never scans a row; the BCP kernels' roofline on SURVEY 8d's ABYTES is in workloads.c3/c4/c5",
                         "bytes_per_launch": kbytes, "kernel_ms": 1e3 * t1, "triage_ms": 1e3 * tri, "deep_ms": 1e3 * dee, "compact_ms": 1e3 * com, "merge_ms": 1e3 * avg("merge_ms"),
                         "all_kernels": {"bytes": triage_bytes + prop_bytes + compact_bytes, "ms": 1e3 * (tri + dee + blk + com),
                                         "frac": (triage_bytes + prop_bytes + compact_bytes) / max(tri + dee + blk + com, 1e-12) / 1e9 / peak},
                         "traffic": traffic,
                         "abytes_2wl_per_launch": abytes_2wl if world == 1 else None,
                         "abytes_2wl_frac_of_step": (abytes_2wl / (total_s / args.steps) / 1e9 / peak) if world == 1 else None},
            "cpu_baseline": cpu,
            "clocks": clk,
            "workloads": workloads,
        }
        if mcheck is not None:
            line["merge_check"] = mcheck
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vars", type=int, default=1_000_000, help="variables per GPU (C2: 1,000,000)")
    ap.add_argument("--ratio", type=float, default=4.2)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--workloads", default="c3,c4,c5", help="BASELINE configs measured beside the C2 headline (N=1: c3,c4,c5; N>1: c5 sharded); 'none' skips")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        ours_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
