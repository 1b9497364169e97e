#!/bin/bash
# round-2 exploration run 2: GPU test suite, second-tier capacity sweep on C3, ncu of the propagation kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_b.log
tail -5 gpurun_out/r02_pytest_b.log
for mt in 16384 32768; do
  BBCP_MID_TRAIL=$mt timeout 600 python tools/bench_workloads.py c3 --no-cpu --steps 1 > gpurun_out/r02_c3_mid$mt.json 2> gpurun_out/r02_c3_mid$mt.err
done
BBCP_BIG_SLABS_PER_SM=16 timeout 600 python tools/bench_workloads.py c3 --no-cpu --steps 1 > gpurun_out/r02_c3_slabs16.json 2> gpurun_out/r02_c3_slabs16.err
grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}' gpurun_out/r02_c3_mid*.json gpurun_out/r02_c3_slabs16.json
CMD="python tools/bench_workloads.py c3 --scale 0.2 --no-cpu --steps 1"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_big|k_deep' -c 3 -f -o gpurun_out/r02_prof_c3s $CMD > gpurun_out/r02_ncu.log 2>&1
tail -3 gpurun_out/r02_ncu.log
