#!/bin/bash
# C3 pass with environment variants "NAME=VALUE,..." given as arguments (one run each; no CPU leg beyond the parity sample)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  echo "== $v"
  env $(echo "$v" | tr ',' ' ') timeout 300 python tools/bench_workloads.py c3 --steps 2 --cpu-seconds 2 > gpurun_out/r02_c3_env$i.json 2> gpurun_out/r02_c3_env$i.err
  grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*\|"tiers": {[^}]*}' gpurun_out/r02_c3_env$i.json; tail -1 gpurun_out/r02_c3_env$i.err | cut -c1-200
done
