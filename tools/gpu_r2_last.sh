#!/bin/bash
# Round 2, last call: what the driver runs at round end (smoke, GPU tests, bench and reference arm with its flags),
# plus the C5 line over 64 steps (">= 64 consecutive steps"; the final lazy flush amortised over them).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== smoke"; date
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench (driver flags)"; date
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -c 300 gpurun_out/r2y_bench.err
echo "== reference arm (driver flags)"; date
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_bench_reference.json 2> gpurun_out/r2y_bench_reference.err; tail -c 300 gpurun_out/r2y_bench_reference.err
echo "== bench 64 steps"; date
timeout 900 python bench.py --gpus 1 --steps 64 --warmup 5 --no-cpu-baseline --no-config-legs > gpurun_out/r2y_bench_64.json 2> gpurun_out/r2y_bench_64.err; tail -c 300 gpurun_out/r2y_bench_64.err
python - <<'PY'
import json
for f in ('gpurun_out/r2y_bench.json','gpurun_out/r2y_bench_64.json','gpurun_out/r2y_bench_reference.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'steps',d.get('steps'),'ms',round(d.get('ms_per_step',0),4),'value',round(d['value']/1e6,2),'M/s','e2e',d.get('e2e',{}).get('ms_per_step'), d.get('e2e',{}).get('value'), 'frac', d.get('roofline',{}).get('frac'))
    except Exception as e:
        print(f,'failed',e)
PY
date
