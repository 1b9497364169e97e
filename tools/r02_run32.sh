#!/bin/bash
# C3 pass + parity sample after a kernel change, behind the GPU parity tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-v14}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_pytest_$TAG.log 2>&1; tail -3 gpurun_out/r02_pytest_$TAG.log
timeout 300 python tools/bench_workloads.py c3 --steps 2 --cpu-seconds 4 > gpurun_out/r02_c3_$TAG.json 2> gpurun_out/r02_c3_$TAG.err
grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*' gpurun_out/r02_c3_$TAG.json; tail -1 gpurun_out/r02_c3_$TAG.err | cut -c1-200
