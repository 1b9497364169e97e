#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python tools/bench_workloads.py c3 --no-cpu --steps 1 > gpurun_out/r02_c3_v10_$name.json 2> gpurun_out/r02_c3_v10_$name.err; echo "$name $(grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}' gpurun_out/r02_c3_v10_$name.json | tr '\n' ' ') $(tail -1 gpurun_out/r02_c3_v10_$name.err | cut -c1-100)"; }
run base X=1
run mid16k BBCP_MID_TRAIL=16384
run slabs5 BBCP_BIG_SLABS_PER_SM=5
run small1024 BBCP_SMALL_TRAIL=1024
