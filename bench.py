#!/usr/bin/env python
"""bench.py -- batched BCP to fixpoint: probes/s (and propagations/s) on BASELINE.json's configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one all-literal failed-literal-probing pass over the instance: every literal of the probe universe
is propagated to its unit-propagation fixpoint over the shared clause database; per probe a conflict flag and
the implied-literal set come back, plus the failed-literal / unit bitmaps.

Workload (config.workload): C2 of SURVEY.md 8d -- random 3-SAT, 1,000,000 variables, clause/variable ratio 4.2,
seed 1 (tools/cnfgen.c), probe all 2,000,000 literals on one B200. With N GPUs the run is a WEAK-scaling one:
N x 1M variables / N x 4.2M clauses, the database replicated on every GPU, the probe universe cut into N
contiguous literal ranges (one per rank, aligned to failed-literal-bitmap words), and the all-gather of the bitmap
slices done by the engine itself over NVLink peer memory inside the batch's last kernel (BBCP_OPT_MERGE; the CUDA
IPC handles are exchanged once through torch.distributed); a device-side rendezvous aligns the ranks before
each timed batch.

Keys: value = probes/s with everything resident in HBM (device-timed, CUDA events on the engine's stream from the
first kernel to the end of the last one incl. the merge, max over ranks); e2e = the same metric through ONE C-ABI
call with HOST buffers (bbcp_probe_lits_to_host: pinned 4-byte probe list in; flags, u32 CSR offsets and implied
literals out; upload / kernels / read-back of successive chunks overlapped); roofline = the kernel of the step with
the longest device time (per-kernel events in a separate pass) on the bytes model of DESIGN.md section 4, against
the measured HBM copy bandwidth; cpu_baseline = the reference's own CTopi::BCP() (oracle/_ref/ref_bcp; the C port
if the harness did not travel) on the box's host cores, one process per core over probe shards, timed in this
same run.
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
METRIC = "probes/sec, batched BCP to fixpoint (all-literal failed-literal probing)"


def ensure_tools():
    if not os.path.exists(CNFGEN):
        subprocess.run(["gcc", "-O2", "-o", CNFGEN, os.path.join(ROOT, "tools", "cnfgen.c"), "-lm"], check=True)


def make_instance(nvars, nclauses, seed, tag):
    ensure_tools()
    d = os.path.join(tempfile.gettempdir(), "bbcp_bench")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, f"3sat_{nvars}_{nclauses}_{seed}_{tag}.bcnf")
    if not os.path.exists(path):
        tmp = path + f".{os.getpid()}.tmp"
        subprocess.run([CNFGEN, "3sat", f"n={nvars}", f"m={nclauses}", f"seed={seed}", f"out={tmp}"], check=True, capture_output=True)
        os.replace(tmp, path)
    return path


def workload_name(nvars, nclauses, seed, n_gpus):
    w = f"C2 random 3-SAT n={nvars} m={nclauses} seed={seed}, probe all {2 * nvars} literals"
    if n_gpus > 1:
        w += f" (weak scaling: {n_gpus} x the 1-GPU instance, DB replicated, probe ranges sharded)"
    return w


class ClockSampler:
    """nvidia-smi samples of SM clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_binary():
    if os.path.exists(REF_BCP):
        return REF_BCP, "reference"
    if not os.path.exists(PORT_BCP):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
    return PORT_BCP, "port"


def run_cpu_shards(inst, n_probes, repeat, cores=None, stride=1):
    """One checker process per host core over contiguous probe shards; each loads the CNF (untimed) and times
    only its probe loop. Returns per-pass aggregate probes/s (= probes / max shard time) and props/s."""
    binary, kind = cpu_binary()
    cores = cores or os.cpu_count() or 1
    procs = [subprocess.Popen([binary, inst, "--probe-all", "--shard", str(i), str(cores), "--stride", str(stride), "--repeat", str(repeat)],
                              stdout=subprocess.PIPE, text=True) for i in range(cores)]
    outs = [json.loads(p.communicate()[0]) for p in procs]
    passes = []
    for r in range(repeat):
        tmax = max(o["seconds_list"][r] for o in outs)
        passes.append(tmax)
    ran = sum(o["ran"] for o in outs)
    impl_last = sum(o["implications"] for o in outs) / repeat
    return {"kind": kind, "cores": cores, "ran": ran, "pass_seconds": passes, "implications_per_pass": impl_last, "conflicts": sum(o["conflicts"] for o in outs)}


def reference_arm(args, rank, world):
    if rank != 0:
        return
    nvars, ncls = args.vars * world, int(args.vars * world * args.ratio)
    inst = make_instance(nvars, ncls, args.seed, "ref")
    n_probes = 2 * nvars
    r = run_cpu_shards(inst, n_probes, args.warmup + args.steps)
    timed = r["pass_seconds"][args.warmup:]
    total = sum(timed)
    value = r["ran"] * len(timed) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "probes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": {"workload": workload_name(nvars, ncls, args.seed, world)},
        "props_per_sec": r["implications_per_pass"] * len(timed) / total,
        "cpu_baseline": {"value": value, "unit": "probes/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": f"all {r['ran']} probes per step, one process per core over contiguous shards, CNF load untimed"},
        "e2e": {"value": value, "unit": "probes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
class DevArr:
    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def ours_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from intel_sat_solver_b200 import BatchedTopor, fileio
    from intel_sat_solver_b200._lib import BatchInfo, Opts

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    nvars, ncls = args.vars * world, int(args.vars * world * args.ratio)
    inst = make_instance(nvars, ncls, args.seed, f"r{local_rank}")
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

    # shard = a contiguous slice of failed-literal-bitmap words (intel_sat_solver_b200/sharding.py)
    from intel_sat_solver_b200 import sharding
    universe = 2 * nvars
    shard = sharding.shard_for(rank, world, nvars)
    words_total = L.bbcp_bitmap_words(h)
    assert words_total == sharding.bitmap_words(nvars)
    first, last = shard.first, shard.last
    n_local = last - first

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # world > 1: BBCP_OPT_MERGE (16) = the bitmap all-gather over NVLink peer memory is the last kernel of the batch;
    # BBCP_OPT_RENDEZVOUS (32) = device-side barrier with all peers right before the timed batch
    mg = (16 | 32) if world > 1 else 0
    opts = Opts(1 | mg, 0, 0, 0, shard.words_per_rank)
    opts_timed = Opts(1 | 8 | mg, 0, 0, 0, shard.words_per_rank)   # + BBCP_OPT_KERNEL_TIMES: per-kernel CUDA events (separate pass)
    info = BatchInfo()
    if world > 1:
        sharding.peer_attach(eng, rank, world)   # CUDA IPC handles exchanged once; the merge itself is the engine's kernel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        """all inputs resident: kernels + compaction (+ on-stream all-gather of the bitmap slices)"""
        rc = L.bbcp_probe_range(h, first, last, C.byref(opts), C.byref(info))
        assert rc == 0, eng.GetStatusExplanation()
        return info.total_ms   # CUDA events on the engine's stream: first kernel .. end of the merge kernel

    clocks = ClockSampler(local_rank)
    stats = {"triage_ms": [], "deep_ms": [], "compact_ms": [], "block_ms": [], "merge_ms": [], "kernel_bytes": 0, "triage_bytes": 0, "launches": 0, "props_ref": 0, "props_all": 0}
    step_ms = []
    for it in range(args.warmup + args.steps):
        flush.fill_(it & 0xff)       # L2 flush between iterations
        barrier()
        if it == 0:
            clocks.start()           # sampled from the first warm-up step to the end of the e2e loop (the timed steps are microseconds long)
        ms = device_step()
        if it >= args.warmup:
            step_ms.append(ms)
            stats["kernel_bytes"], stats["triage_bytes"] = info.kernel_bytes, info.triage_bytes
            stats["launches"] += info.launches
            stats["props_ref"], stats["props_all"] = info.props_ref, info.props_all
    n_conflict, n_implied = info.n_conflict, info.n_implied
    # per-kernel durations for the roofline: the same steps again with CUDA events between the kernels
    # (BBCP_OPT_KERNEL_TIMES; kept out of the loop above because the extra events cost stream time)
    for it in range(args.steps):
        flush.fill_(it & 0xff)
        barrier()
        rc = L.bbcp_probe_range(h, first, last, C.byref(opts_timed), C.byref(info))
        assert rc == 0, eng.GetStatusExplanation()
        for k in ("triage_ms", "deep_ms", "compact_ms", "block_ms", "merge_ms"):
            stats[k].append(getattr(info, k))
    barrier()
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    probes = torch.tensor([float(n_local)], dtype=torch.float64, device=dev)
    props = torch.tensor([float(stats["props_ref"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(probes)
        dist.all_reduce(props)
    total_s = float(total_ms.item()) / 1e3
    value = float(probes.item()) * args.steps / total_s

    # ---- e2e: the call a user makes, host buffers in and out, every step
    idx = np.arange(first, last, dtype=np.int64)
    h_lits = torch.from_numpy(np.where(idx & 1, -(idx // 2 + 1), idx // 2 + 1).astype(np.int32)).pin_memory()
    o_flag = torch.empty(n_local + 8, dtype=torch.uint8).pin_memory()
    o_off = torch.empty(n_local + 1, dtype=torch.int32).pin_memory()   # u32 CSR offsets (BBCP_OPT_OFF32)
    opts_e2e = Opts(1 | 4, 0, 0, 0, 0)
    o_lits = torch.empty(max(int(n_implied) * 2, 1024), dtype=torch.int32).pin_memory()
    e2e_s = []
    for it in range(args.warmup + args.steps):
        flush.fill_(it & 0xff)
        barrier()
        t0 = time.perf_counter()
        # ONE call, host buffers in and out: chunks of the list are uploaded, propagated and read back overlapped
        rc = L.bbcp_probe_lits_to_host(h, h_lits.data_ptr(), n_local, C.byref(opts_e2e), o_flag.data_ptr(), o_off.data_ptr(), o_lits.data_ptr(),
                                       o_lits.numel(), C.byref(info))
        assert rc == 0, eng.GetStatusExplanation()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            e2e_s.append(dt)
    clk = clocks.stop()
    e2e_total = torch.tensor([sum(e2e_s)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    e2e_value = float(probes.item()) * args.steps / float(e2e_total.item())
    h2d = h_lits.numel() * 4
    d2h = n_local + (n_local + 1) * 4 + int(info.n_implied) * 4
    # sanity on the last e2e step: every probe of this workload implies exactly itself
    assert int(o_off[n_local]) == info.n_implied and int((o_flag[:n_local] == 0).sum()) == info.n_ok

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = run_cpu_shards(inst, universe, 2)
        v = r["ran"] / r["pass_seconds"][-1]
        cpu = {"value": v, "unit": "probes/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"all {r['ran']} probes once (after one warm pass), one process per core over contiguous shards, CNF load untimed",
               "props_per_sec": r["implications_per_pass"] / r["pass_seconds"][-1], "failed_literals": r["conflicts"] // 2}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        avg = lambda k: sum(stats[k]) / max(len(stats[k]), 1) / 1e3
        tri, dee, com, blk = avg("triage_ms"), avg("deep_ms"), avg("compact_ms"), avg("block_ms")
        # bytes model of DESIGN.md section 4, per launch: triage; propagation (engine counters); compaction =
        # 4 B size read + 8 B offset + 4 B literal written per item here (+ the tile sums, negligible)
        # (every implied set of this workload is the probe's own literal, regenerated, so nothing is read back)
        prop_bytes = stats["kernel_bytes"] - stats["triage_bytes"]
        compact_bytes = n_local * (4 + 8) + int(n_implied) * 4
        cands = [("k_triage (8 probes per thread, streaming)", stats["triage_bytes"], tri),
                 ("k_deep (warp-per-probe propagate, shared-memory state)", prop_bytes, dee + blk),
                 ("k_scan_gather (single-pass scan + gather on tile sums)", compact_bytes, com)]
        # the dominant kernel of the step = the one with the longest device time
        kname, kbytes, t1 = max(cands, key=lambda c: c[2])
        achieved = kbytes / t1 / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kname.split(" ")[0])
        except (OSError, ValueError):
            pass
        # SURVEY.md 8d ABYTES for this instance family in closed form: T={p}, no binaries, every long
        # watch visited: 16 + 24*nlong(not p); sum of nlong over literals = 2m
        abytes_2wl = 16 * universe + 24 * 2 * ncls
        line = {
            "metric": METRIC, "value": value, "unit": "probes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(nvars, ncls, args.seed, world), "l2": "flushed between timed iterations (256 MiB device write)",
                       "db": {k: dbi[k] for k in ("n_bin", "n_tern", "n_long", "row_slots", "device_bytes")},
                       "results": "flag + sorted implied-literal set per probe, failed/unit bitmaps", "db_build_s": round(build_s, 2)},
            "props_per_sec": float(props.item()) * args.steps / total_s,
            "failed_literals": int(n_conflict),
            "e2e": {"value": e2e_value, "unit": "probes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * float(e2e_total.item()) / args.steps},
            "gpu_launches": stats["launches"],
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback",
                         "bytes_per_launch": kbytes, "kernel_ms": 1e3 * t1, "triage_ms": 1e3 * tri, "deep_ms": 1e3 * dee, "compact_ms": 1e3 * com, "merge_ms": 1e3 * avg("merge_ms"),
                         "all_kernels": {"bytes": stats["triage_bytes"] + prop_bytes + compact_bytes, "ms": 1e3 * (tri + dee + blk + com),
                                         "frac": (stats["triage_bytes"] + prop_bytes + compact_bytes) / max(tri + dee + blk + com, 1e-12) / 1e9 / peak},
                         "traffic": traffic,
                         "abytes_2wl_per_launch": abytes_2wl if world == 1 else None,
                         "abytes_2wl_frac_of_step": (abytes_2wl / (total_s / args.steps) / 1e9 / peak) if world == 1 else None},
            "cpu_baseline": cpu,
            "clocks": clk,
        }
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
