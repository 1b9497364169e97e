#!/bin/bash
# round-2 run 12: deferred flag publication + first-tier capacity sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_i.log 2>&1; tail -3 gpurun_out/r02_pytest_i.log
for st in 512 256 128; do
BBCP_SMALL_TRAIL=$st timeout 300 python tools/bench_workloads.py c5 --no-cpu --steps 2 > gpurun_out/r02_c5_v8_$st.json 2> gpurun_out/r02_c5_v8_$st.err
BBCP_SMALL_TRAIL=$st timeout 300 python tools/bench_workloads.py c3 --no-cpu --steps 1 > gpurun_out/r02_c3_v8_$st.json 2> gpurun_out/r02_c3_v8_$st.err
BBCP_SMALL_TRAIL=$st timeout 300 python tools/bench_workloads.py c4 --no-cpu --steps 2 > gpurun_out/r02_c4_v8_$st.json 2> gpurun_out/r02_c4_v8_$st.err
for f in c5_v8_$st c3_v8_$st c4_v8_$st; do echo $f; grep -h -o '"ms_per_step": [0-9.]*\|"kernel_ms": {[^}]*}' gpurun_out/r02_$f.json; tail -1 gpurun_out/r02_$f.err | cut -c1-200; done
done
