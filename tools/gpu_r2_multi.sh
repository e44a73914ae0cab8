#!/bin/bash
# Round 2, multi-GPU call: real-rank parity tests (N = 2 or 4), then the N-GPU bench line (push / pull / nccl).
# usage: bash tools/gpu_r2_multi.sh <N> [modes...]
set -u
N=${1:-2}; shift
MODES=${*:-push pull}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -8
echo "== real-rank parity tests"; date
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2m_pytest_g$N.log; tail -8 gpurun_out/r2m_pytest_g$N.log
for m in $MODES; do
  echo "== bench N=$N exchange=$m"; date
  extra=""; if [ $m != push ]; then extra="--no-parity --no-config-legs"; fi
  INVPREF_EXCHANGE=$m timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 3 $extra > gpurun_out/r2m_bench_g${N}_$m.json 2> gpurun_out/r2m_bench_g${N}_$m.err
  tail -c 300 gpurun_out/r2m_bench_g${N}_$m.err
  python - gpurun_out/r2m_bench_g${N}_$m.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print('no line', e); sys.exit(0)
print('N',d['n_gpus'],d['config'].get('exchange'),'ms',round(d['ms_per_step'],4),'value',round(d['value']/1e9,3),'G/s e2e',d['e2e'] and round(d['e2e']['ms_per_step'],3))
print(' phases',{k:(round(v,4) if isinstance(v,float) else v) for k,v in d['rank0_phase_ms'].items()})
print(' nvlink',{k:(round(v,3) if isinstance(v,float) else v) for k,v in d['nvlink'].items() if k!='note'})
print(' parity',json.dumps(d.get('parity_vs_1gpu'))[:600])
print(' c4',json.dumps(d.get('configs',{}).get('c4'))[:700])
PY
done
date
