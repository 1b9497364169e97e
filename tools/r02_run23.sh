#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_l.log 2>&1; tail -3 gpurun_out/r02_pytest_l.log
timeout 300 python tools/bench_workloads.py c3 --steps 2 --cpu-seconds 4 > gpurun_out/r02_c3_v12.json 2> gpurun_out/r02_c3_v12.err
grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*' gpurun_out/r02_c3_v12.json; tail -1 gpurun_out/r02_c3_v12.err | cut -c1-200
bash tools/r02_run22.sh
