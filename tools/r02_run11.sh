#!/bin/bash
# round-2 run 11: hot-row mode
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_h.log 2>&1; tail -3 gpurun_out/r02_pytest_h.log
timeout 300 python tools/bench_workloads.py c5 --scale 0.2 --steps 2 > gpurun_out/r02_c5s_v7.json 2> gpurun_out/r02_c5s_v7.err
timeout 300 python tools/bench_workloads.py c5 --no-cpu --steps 2 > gpurun_out/r02_c5_v7.json 2> gpurun_out/r02_c5_v7.err
timeout 300 python tools/bench_workloads.py c3 --steps 1 --cpu-seconds 5 > gpurun_out/r02_c3_v7.json 2> gpurun_out/r02_c3_v7.err
timeout 300 python tools/bench_workloads.py c4 --steps 2 > gpurun_out/r02_c4_v7.json 2> gpurun_out/r02_c4_v7.err
for f in c5s_v7 c5_v7 c3_v7 c4_v7; do echo $f; grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*\|"slots": [0-9]*' gpurun_out/r02_$f.json; tail -1 gpurun_out/r02_$f.err | cut -c1-200; done
