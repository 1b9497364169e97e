#!/bin/bash
# round-2 run 5: adaptive second-tier hash + narrow-frontier mode; tests; block tiers again for the A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_e.log 2>&1; tail -3 gpurun_out/r02_pytest_e.log
timeout 300 python tools/bench_workloads.py c3 --steps 1 --cpu-seconds 6 > gpurun_out/r02_c3_v5.json 2> gpurun_out/r02_c3_v5.err
BBCP_TIER2=block timeout 300 python tools/bench_workloads.py c3 --no-cpu --steps 1 > gpurun_out/r02_c3_v5_block.json 2> gpurun_out/r02_c3_v5_block.err
BBCP_NO_ORDER=1 timeout 300 python tools/bench_workloads.py c3 --no-cpu --steps 1 > gpurun_out/r02_c3_v5_noorder.json 2> gpurun_out/r02_c3_v5_noorder.err
timeout 300 python tools/bench_workloads.py c5 --scale 0.2 --no-cpu --steps 2 > gpurun_out/r02_c5s_v5.json 2> gpurun_out/r02_c5s_v5.err
timeout 300 python tools/bench_workloads.py c4 --steps 2 > gpurun_out/r02_c4_v5.json 2> gpurun_out/r02_c4_v5.err
for f in c3_v5 c3_v5_block c3_v5_noorder c5s_v5 c4_v5; do echo $f; grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*\|"tiers": {[^}]*}' gpurun_out/r02_$f.json; tail -1 gpurun_out/r02_$f.err | cut -c1-200; done
