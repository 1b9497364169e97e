#!/bin/bash
# round-2 exploration run 3: CTA-per-probe shared-hash tiers (k_block) against the warp-per-probe second tier (BBCP_TIER2=warp)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_c.log
tail -5 gpurun_out/r02_pytest_c.log
timeout 600 python tools/bench_workloads.py c3 --steps 1 --cpu-seconds 8 > gpurun_out/r02_c3_block.json 2> gpurun_out/r02_c3_block.err
BBCP_TIER2=warp timeout 600 python tools/bench_workloads.py c3 --no-cpu --steps 1 > gpurun_out/r02_c3_warp.json 2> gpurun_out/r02_c3_warp.err
BBCP_MID_TRAIL=4096 timeout 600 python tools/bench_workloads.py c3 --no-cpu --steps 1 > gpurun_out/r02_c3_block_m4096.json 2> gpurun_out/r02_c3_block_m4096.err
timeout 300 python tools/bench_workloads.py c5 --scale 0.2 --steps 2 > gpurun_out/r02_c5s_block.json 2> gpurun_out/r02_c5s_block.err
timeout 300 python tools/bench_workloads.py c4 --steps 2 --no-cpu > gpurun_out/r02_c4_block.json 2> gpurun_out/r02_c4_block.err
grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*\|"tiers": {[^}]*}' gpurun_out/r02_c3_block.json gpurun_out/r02_c3_warp.json gpurun_out/r02_c3_block_m4096.json gpurun_out/r02_c5s_block.json gpurun_out/r02_c4_block.json
CMD="python tools/bench_workloads.py c3 --scale 0.2 --no-cpu --steps 1"
$CMD > gpurun_out/r02_ncu_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_big|k_block' -c 3 -f -o gpurun_out/r02_prof_c3s_block $CMD > gpurun_out/r02_ncu3.log 2>&1
tail -3 gpurun_out/r02_ncu3.log
