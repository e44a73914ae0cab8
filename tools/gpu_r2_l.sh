#!/bin/bash
# Round 2, GPU call L: compile-time D = 40 instantiations, pipelined e2e read-back -- tests, smoke, phases of c2/c3/c4, default bench line.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'lazy ms', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline']['phase_ms'].items() if x>0.015})
PY
}
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for wl in c2 c3 c4; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-config-legs --nbatch 3 > gpurun_out/r2m1_$wl.json 2> gpurun_out/r2m1_$wl.err; show gpurun_out/r2m1_$wl.json "$wl"
done
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2m1_pytest.log; tail -6 gpurun_out/r2m1_pytest.log
echo "== bench default"; date
timeout 900 python bench.py > gpurun_out/r2m1_bench.json 2> gpurun_out/r2m1_bench.err; tail -c 400 gpurun_out/r2m1_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m1_bench.json').read().strip().splitlines()[-1])
print('C5', round(d['ms_per_step'],4), 'upass frac', round(d['roofline']['frac'],3), 'item frac', round(d['roofline']['item_pass']['frac'],3), 'step frac', round(d['roofline']['step']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'cluster', round(d['cluster']['value']/1e9,3), round(d['cluster']['roofline']['frac'],3), d['cluster']['ms_min_max'])
print(' phases', {k:round(x,4) for k,x in d['roofline']['phase_ms'].items() if x>0.015})
print(' dense', round(d['dense_adam']['ms_per_step'],4), round(d['dense_adam']['roofline_frac'],3))
for k,v in d.get('configs',{}).items():
    print(k, {x: (round(v[x],4) if isinstance(v[x],float) else v[x]) for x in ('ms_per_step','launches_per_step','value') if x in v}, 'nograph', v.get('no_graph',{}).get('ms_per_step'), 'eager x', v.get('torch_eager_gpu',{}).get('speedup_of_value'), 'cluster', v.get('cluster',{}).get('ms'), v.get('unavailable'))
PY
date
