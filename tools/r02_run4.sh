#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export BBCP_DEBUG_SYNC=2
run() { name=$1; shift; timeout 120 env "$@" python -u tools/bench_workloads.py c3 --scale $SC --no-cpu --steps 1 > gpurun_out/r02_dbg_$name.log 2>&1; echo "$name rc=$?"; grep -A14 -E "still running|check failed|BBCP_CHECK|fault|illegal" gpurun_out/r02_dbg_$name.log | head -18 | cut -c1-300; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}' gpurun_out/r02_dbg_$name.log | head -2; }
SC=0.04 BBCP_LIB=$PWD/intel_sat_solver_b200/libbbcp_checks.so run f_mid512_noreuse BBCP_MID_TRAIL=512 BBCP_NO_REUSE=1
SC=0.1 BBCP_LIB=$PWD/intel_sat_solver_b200/libbbcp_checks.so run c_default X=1
unset BBCP_DEBUG_SYNC
timeout 300 python tools/bench_workloads.py c3 --steps 1 --cpu-seconds 8 > gpurun_out/r02_c3_block.json 2> gpurun_out/r02_c3_block.err
grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*\|"tiers": {[^}]*}\|"reuse": {[^}]*}' gpurun_out/r02_c3_block.json; tail -2 gpurun_out/r02_c3_block.err
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_d.log 2>&1; tail -3 gpurun_out/r02_pytest_d.log
