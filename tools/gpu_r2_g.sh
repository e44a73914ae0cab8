#!/bin/bash
# Round 2, GPU call G: aligned rings -- A/B against the round-1 tree, GPU tests, default bench, compute-sanitizer.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'lazy ms', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline']['phase_ms'].items() if x>0.02}, '| dense rows_users', round(d['dense_adam']['phase_ms']['rows_users'],4), '| cluster ms', round(d['cluster']['ms'],3))
PY
}
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config-legs --nbatch 4 > gpurun_out/r2g_cur.json 2> gpurun_out/r2g_cur.err; show gpurun_out/r2g_cur.json current
(cd build/r1repo && timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > ../../gpurun_out/r2g_r1tree.json 2> ../../gpurun_out/r2g_r1tree.err); show gpurun_out/r2g_r1tree.json r1tree
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2g_pytest.log; tail -6 gpurun_out/r2g_pytest.log
echo "== bench default"; date
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 400 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print('C5', round(d['ms_per_step'],4), 'upass frac', round(d['roofline']['frac'],3), 'item frac', round(d['roofline']['item_pass']['frac'],3), 'step frac', round(d['roofline']['step']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'cluster', round(d['cluster']['value']/1e9,3), round(d['cluster']['roofline']['frac'],3))
print(' dense', round(d['dense_adam']['ms_per_step'],4), round(d['dense_adam']['roofline_frac'],3))
for k,v in d.get('configs',{}).items():
    print(k, {x: (round(v[x],4) if isinstance(v[x],float) else v[x]) for x in ('ms_per_step','launches_per_step','value') if x in v}, 'nograph', v.get('no_graph',{}).get('ms_per_step'), 'eager x', v.get('torch_eager_gpu',{}).get('speedup_of_value'), 'cluster', v.get('cluster',{}).get('ms'), v.get('unavailable'))
PY
echo "== sanitizer"; date
K='peer_memory_paths or fused_evaluator or (graph_epochs and c2 and True) or train_step_matches_reference or cluster_matches or bounded_plan'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_parallel.py tests/test_gpu_drivers.py tests/test_gpu_configs.py tests/test_gpu_trainer.py -x -q -k "$K" > gpurun_out/r2g_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2g_memcheck.log
K2='peer_memory_paths or fused_evaluator_ties or train_step_matches_reference or cluster_matches_reference'
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_parallel.py tests/test_gpu_drivers.py -x -q -k "$K2" > gpurun_out/r2g_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2g_racecheck.log
date
