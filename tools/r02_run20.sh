#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_k.log 2>&1; tail -4 gpurun_out/r02_pytest_k.log
timeout 300 python tools/bench_workloads.py c3 --steps 2 --cpu-seconds 5 > gpurun_out/r02_c3_v9.json 2> gpurun_out/r02_c3_v9.err
grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}\|"mismatches": [0-9]*' gpurun_out/r02_c3_v9.json; tail -1 gpurun_out/r02_c3_v9.err | cut -c1-200
sed -n '/^python - <<.PY.$/,/^PY$/p' tools/r02_run15.sh | sed '1d;$d' > /tmp/fix.py; python /tmp/fix.py | tail -45
