#!/bin/bash
# quick multi-GPU headline check: C2 weak scaling only (no workloads), N = $1
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rep in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$rep bench.py --gpus $N --workloads none --no-cpu > gpurun_out/r02_quick_n${N}_$rep.json 2> gpurun_out/r02_quick_n${N}_$rep.err
python - $N $rep <<'PY'
import json,sys
N,rep=sys.argv[1:3]
try:
    d=json.loads(open(f'gpurun_out/r02_quick_n{N}_{rep}.json').read().strip().splitlines()[-1])
    print('N',N,'value',d['value'],'us',1e3*d['ms_per_step'],'e2e',d['e2e']['value'], 'merge_check', d.get('merge_check',{}).get('mismatches'))
    print('  rank0', d['details']['per_rank'][0]['step_ms'], {k:round(1e3*v['median'],1) for k,v in d['details']['per_rank'][0]['kernel_ms'].items()})
except Exception as e: print('ERR', e); print(open(f'gpurun_out/r02_quick_n{N}_{rep}.err').read()[-1500:])
PY
done
