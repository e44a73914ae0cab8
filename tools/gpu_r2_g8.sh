#!/bin/bash
# Round 2, 8-GPU call: the default N-GPU bench line, then the same C5 leg with the step's two synchronisation points
# as NCCL all-reduces (INVPREF_SYNC=nccl) for comparison.
set -u
N=${1:-8}
bash tools/gpu_r2_multi_dev.sh $N notest
echo "== INVPREF_SYNC=nccl"; date
INVPREF_SYNC=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-parity --no-config-legs > gpurun_out/r2p_bench_g${N}_ncclsync.json 2> gpurun_out/r2p_bench_g${N}_ncclsync.err
python - gpurun_out/r2p_bench_g${N}_ncclsync.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('N',d['n_gpus'],'nccl sync: ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['ms_per_step'],3))
print(' phases',{k:(round(v,4) if isinstance(v,float) else v) for k,v in d['rank0_phase_ms'].items()})
PY
date
