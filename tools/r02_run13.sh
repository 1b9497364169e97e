#!/bin/bash
# round-2 run 13: the north star's staging ideas, measured: bulk-async staging of long rows, L2 persistence window
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift 2; env "$@" timeout 300 python tools/bench_workloads.py $wl --no-cpu --steps 2 > gpurun_out/r02_g1_${wl}_$name.json 2> gpurun_out/r02_g1_${wl}_$name.err; echo "$wl $name $(grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/r02_g1_${wl}_$name.json) $(tail -1 gpurun_out/r02_g1_${wl}_$name.err | cut -c1-120)"; }
for wl in c5 c3 c4; do
run base $wl X=1
run tma $wl BBCP_TMA_ROWS=1
run l2hdr $wl BBCP_L2_WINDOW=hdr
run l2lit $wl BBCP_L2_WINDOW=litinfo
done
BBCP_TMA_ROWS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "goldens or synthetic or option" > gpurun_out/r02_pytest_tma.log 2>&1; tail -2 gpurun_out/r02_pytest_tma.log
