#!/bin/bash
# workload passes with environment variants: arguments "workload[@scale]:NAME=VALUE,..." (one run each)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
i=0
for a in "$@"; do
  i=$((i+1))
  w=${a%%:*}; v=${a#*:}
  wl=${w%%@*}; sc=""
  [ "$w" != "$wl" ] && sc="--scale ${w#*@}"
  echo "== $w $v"
  env $(echo "$v" | tr ',' ' ') timeout 400 python tools/bench_workloads.py $wl $sc --steps 2 --cpu-seconds 2 > gpurun_out/r02_wl_env$i.json 2> gpurun_out/r02_wl_env$i.err
  grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*\|"tiers": {[^}]*}' gpurun_out/r02_wl_env$i.json; tail -1 gpurun_out/r02_wl_env$i.err | cut -c1-200
done
