#!/bin/bash
# sweep of the last tier with 16-byte loads (C3 pass + parity sample), GPU parity tests that reach the last tier, e2e time line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_pytest_m.log 2>&1; tail -3 gpurun_out/r02_pytest_m.log
timeout 300 python tools/bench_workloads.py c3 --steps 2 --cpu-seconds 4 > gpurun_out/r02_c3_v13.json 2> gpurun_out/r02_c3_v13.err
grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*' gpurun_out/r02_c3_v13.json; tail -1 gpurun_out/r02_c3_v13.err | cut -c1-200
timeout 300 python tools/e2e_trace.py > gpurun_out/r02_e2e_trace.log 2>&1; cat gpurun_out/r02_e2e_trace.log | cut -c1-260
