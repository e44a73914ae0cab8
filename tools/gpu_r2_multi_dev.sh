#!/bin/bash
# Round 2, multi-GPU development call: the push cases of the real-rank tests (skipped with a second argument
# `notest`), then the N-GPU bench line.
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if [ "${2:-tests}" != notest ]; then
echo "== real-rank tests (push)"; date
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "push" 2>&1 | tail -40 > gpurun_out/r2p_pytest_g$N.log; tail -25 gpurun_out/r2p_pytest_g$N.log
fi
echo "== bench N=$N"; date
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2p_bench_g${N}.json 2> gpurun_out/r2p_bench_g${N}.err
tail -c 1500 gpurun_out/r2p_bench_g${N}.err
python - gpurun_out/r2p_bench_g${N}.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print('no line', e); sys.exit(0)
print('N',d['n_gpus'],d['config'].get('exchange'),'ms',round(d['ms_per_step'],4),'value',round(d['value']/1e9,3),'G/s e2e',d['e2e'] and round(d['e2e']['ms_per_step'],3))
print(' phases',{k:(round(v,4) if isinstance(v,float) else v) for k,v in d['rank0_phase_ms'].items()})
print(' parity',json.dumps(d.get('parity_vs_1gpu'))[:900])
for k,v in d.get('configs',{}).items():
    print(' ',k,json.dumps({a:b for a,b in v.items() if a not in ('workload','parallelism','timing','rank0_phase_ms','per_rank_value')})[:500])
PY
date
