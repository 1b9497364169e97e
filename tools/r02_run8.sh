#!/bin/bash
# round-2 multi-GPU run: merge parity tests for the worlds this box has, then the bench at N = $1 (merge_check + C5 sharded)
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -rs > gpurun_out/r02_pytest_mgpu_n$N.log 2>&1; tail -5 gpurun_out/r02_pytest_mgpu_n$N.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err ) 2> gpurun_out/r02_bench_n$N.time
tail -5 gpurun_out/r02_bench_n$N.err | cut -c1-300; cat gpurun_out/r02_bench_n$N.time
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r02_bench_n{N}.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'clocks', d['clocks'])
    print('merge_check', d.get('merge_check'))
    for r in d['details']['per_rank']: print(r['rank'], r['step_ms'], {k:v['median'] for k,v in r['kernel_ms'].items()})
    print('c5', d['workloads'].get('c5'))
except Exception as e: print('ERR', e)
PY
