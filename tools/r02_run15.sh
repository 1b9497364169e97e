#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_j.log 2>&1; tail -6 gpurun_out/r02_pytest_j.log
python - <<'PY'
# per-round cost of failed-literal probing to a fixpoint: unit feedback on the device against the round-1 host rebuild
import os, sys, time, json, subprocess
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import gpu_util, tempfile
from intel_sat_solver_b200 import BatchedTopor, fileio
d = tempfile.mkdtemp()
out = {}
for kind, kw in (("aig", dict(frames=50, gates=66667, inputs=2000, latches=2000, window=5000, seed=2)), ("3sat", dict(n=1000000, m=4200000, seed=1)), ("comm", dict(n=1250000, m=5000000, comms=1000, seed=3))):
    inst, _ = gpu_util.gen(kind, d, **kw)
    words = fileio.read_bcnf(inst)
    for mode in ("device", "rebuild"):
        prog = ("import sys, json, time; sys.path.insert(0, %r)\nfrom intel_sat_solver_b200 import BatchedTopor, fileio\n"
                "t = BatchedTopor(); t.AddClauses(fileio.read_bcnf(%r)); t.Finalize(); t.ProbeAll(implied=False, fetch=False)\n"
                "t0 = time.perf_counter(); g = t.ProbeToFixpoint(); dt = time.perf_counter() - t0\n"
                "print(json.dumps({'got': g, 'seconds': dt, 'level0': int(t.Level0().size)}))" % (os.getcwd(), inst))
        env = dict(os.environ)
        if mode == "rebuild": env["BBCP_FIXPOINT"] = "rebuild"
        r = json.loads(subprocess.run([sys.executable, "-c", prog], check=True, capture_output=True, text=True, env=env).stdout)
        out[f"{kind}/{mode}"] = {**r["got"], "level0": r["level0"], "seconds": round(r["seconds"], 4), "ms_per_round": round(1e3 * r["seconds"] / max(r["got"]["rounds"], 1), 3)}
json.dump(out, open("gpurun_out/r02_fixpoint_rounds.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
