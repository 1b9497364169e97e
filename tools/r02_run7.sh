#!/bin/bash
# round-2 run 7: full GPU suite + the complete default bench (C2 headline + workloads block) + reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_g.log 2>&1; tail -4 gpurun_out/r02_pytest_g.log
( time timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2> gpurun_out/r02_bench_n1.time
tail -3 gpurun_out/r02_bench_n1.err | cut -c1-300; cat gpurun_out/r02_bench_n1.time
timeout 120 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['full_format']['value'])
print('cpu',d['cpu_baseline'] and d['cpu_baseline']['value'])
for k,w in d['workloads'].items():
    print(k, {x:w.get(x) for x in ('ms_per_step','value','parity_sample','error','gen_s','db_build_s')}, w.get('roofline',{}).get('frac'), w.get('cpu_baseline',{}).get('value'))
PY
