#!/bin/bash
# Round 2: hand-written radix sort / scans in the plan build: GPU tests, plan timing, default bench line.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-r2n}
echo "== plan timing + sort check"; date
timeout 600 python tools/time_plan.py > gpurun_out/${TAG}_plan.json 2> gpurun_out/${TAG}_plan.err; cat gpurun_out/${TAG}_plan.json; tail -5 gpurun_out/${TAG}_plan.err
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
echo "== bench default"; date
timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench.err
python - gpurun_out/${TAG}_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('C5', round(d['ms_per_step'],4), 'frac', d['roofline']['frac'], 'e2e', d['e2e'].get('ms_per_step'), d['e2e'].get('value'))
print(' phases', d.get('phase_ms'))
for k,v in d.get('configs',{}).items(): print(k, v.get('ms_per_step'), v.get('launches_per_step'))
PY
date
