#!/bin/bash
# round-2 exploration run 1: GPU test suite, then C3/C4/C5 with and without closure reuse
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_a.log
tail -5 gpurun_out/r02_pytest_a.log
for v in noreuse reuse noorder; do
  case $v in
    noreuse) export BBCP_NO_REUSE=1; unset BBCP_NO_ORDER;;
    reuse) unset BBCP_NO_REUSE; unset BBCP_NO_ORDER;;
    noorder) unset BBCP_NO_REUSE; export BBCP_NO_ORDER=1;;
  esac
  timeout 600 python tools/bench_workloads.py c3 --no-cpu --steps 2 > gpurun_out/r02_c3_$v.json 2> gpurun_out/r02_c3_$v.err
  timeout 300 python tools/bench_workloads.py c5 --scale 0.2 --no-cpu --steps 2 > gpurun_out/r02_c5s_$v.json 2> gpurun_out/r02_c5s_$v.err
done
unset BBCP_NO_REUSE BBCP_NO_ORDER
timeout 600 python tools/bench_workloads.py c3 --steps 1 --cpu-seconds 10 > gpurun_out/r02_c3_parity.json 2> gpurun_out/r02_c3_parity.err
timeout 300 python tools/bench_workloads.py c4 --steps 2 > gpurun_out/r02_c4.json 2> gpurun_out/r02_c4.err
grep -h -o '"ms_per_step": [0-9.]*\|"mismatches": [0-9]*\|"reuse": {[^}]*}' gpurun_out/r02_c*.json
