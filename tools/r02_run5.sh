#!/bin/bash
# round-2 run 6: SKIP-mode adoption (second walks for the smaller side)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_f.log 2>&1; tail -3 gpurun_out/r02_pytest_f.log
timeout 300 python tools/bench_workloads.py c3 --steps 1 --cpu-seconds 6 > gpurun_out/r02_c3_v6.json 2> gpurun_out/r02_c3_v6.err
BBCP_TIER2=block timeout 300 python tools/bench_workloads.py c3 --no-cpu --steps 1 > gpurun_out/r02_c3_v6_block.json 2> gpurun_out/r02_c3_v6_block.err
timeout 300 python tools/bench_workloads.py c5 --scale 0.2 --steps 2 > gpurun_out/r02_c5s_v6.json 2> gpurun_out/r02_c5s_v6.err
for f in c3_v6 c3_v6_block c5s_v6; do echo $f; grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*\|"tiers": {[^}]*}\|"slots": [0-9]*\|"props_all": [0-9]*' gpurun_out/r02_$f.json; tail -1 gpurun_out/r02_$f.err | cut -c1-200; done
