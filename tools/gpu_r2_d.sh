#!/bin/bash
# Round 2, GPU call D: same-box A/B of the round-1 tree vs the current one (C5 step), GPU tests, default bench line.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'lazy ms', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline']['phase_ms'].items() if x>0.02}, '| dense', {k:round(x,4) for k,x in d['dense_adam']['phase_ms'].items() if x>0.02}, '| cluster ms', round(d['cluster']['ms'],3), d['cluster']['samples'])
PY
}
echo "== r1 tree"; date
(cd build/r1repo && timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > ../../gpurun_out/r2d_r1tree.json 2> ../../gpurun_out/r2d_r1tree.err); show gpurun_out/r2d_r1tree.json r1tree
echo "== current tree, same arguments"; date
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config-legs --nbatch 4 > gpurun_out/r2d_cur.json 2> gpurun_out/r2d_cur.err; show gpurun_out/r2d_cur.json current
echo "== r1 tree again"; date
(cd build/r1repo && timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > ../../gpurun_out/r2d_r1tree2.json 2> ../../gpurun_out/r2d_r1tree2.err); show gpurun_out/r2d_r1tree2.json r1tree
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2d_pytest.log; tail -5 gpurun_out/r2d_pytest.log
echo "== bench default"; date
timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 400 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print('C5', round(d['ms_per_step'],4), 'upass frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'cluster', round(d['cluster']['value']/1e9,3))
for k,v in d.get('configs',{}).items():
    print(k, {x: (round(v[x],4) if isinstance(v[x],float) else v[x]) for x in ('ms_per_step','launches_per_step','value') if x in v}, 'nograph', v.get('no_graph',{}).get('ms_per_step'), 'eager x', v.get('torch_eager_gpu',{}).get('speedup_of_value'), v.get('unavailable'))
PY
date
