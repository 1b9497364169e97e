#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
BBCP_LIB=$PWD/intel_sat_solver_b200/libbbcp_checks.so python - <<'PY' > gpurun_out/r02_phase_ticks.json 2> gpurun_out/r02_phase_ticks.err
import sys, os, json, ctypes as C, tempfile
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, gpu_util
from intel_sat_solver_b200 import BatchedTopor, fileio
d = tempfile.mkdtemp()
inst, _ = gpu_util.gen("aig", d, frames=50, gates=66667, inputs=2000, latches=2000, window=5000, seed=2)
t = BatchedTopor(); t.AddClauses(fileio.read_bcnf(inst)); t.Finalize()
t.ProbeAll(fetch=False); t.ProbeAll(fetch=False)
buf = (C.c_uint64 * 8)()
t.lib.bbcp_debug_phase_ticks(t.handle, buf)
r = t.ProbeAll(fetch=False, kernel_times=True)
t.lib.bbcp_debug_phase_ticks(t.handle, buf)
ticks = list(buf)
tot = sum(ticks) or 1
names = ["narrow mode", "lookup+adoption (CTA)", "row walks (CTA)", "load", "emission: sweep", "cleanup+publish+ticket", "emission: alloc + range", "SWEEP WORDS (count, not ticks)"]
print(json.dumps({"kernel_ms": {k: r.info[k] for k in ("deep_ms", "mid_ms", "block_ms")}, "n_large": r.info["n_large"],
                  "phase_share": {n: round(x / tot, 4) for n, x in zip(names, ticks)}, "ticks": ticks}))
PY
cat gpurun_out/r02_phase_ticks.json; tail -2 gpurun_out/r02_phase_ticks.err
