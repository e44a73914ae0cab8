#!/bin/bash
# Round 2, GPU call H: L2 prefetch (user pass, re-assignment) on/off, GPU tests.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'lazy ms', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline']['phase_ms'].items() if x>0.02}, '| dense rows_users', round(d['dense_adam']['phase_ms']['rows_users'],4), '| cluster ms', round(d['cluster']['ms'],3), d['cluster'].get('ms_min_max'), d['cluster']['samples'])
PY
}
for rep in 1 2; do
for v in main nopf; do
  if [ $v = main ]; then unset INVPREF_LIB; else export INVPREF_LIB=$PWD/build/variants/libinvpref_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config-legs --nbatch 6 > gpurun_out/r2h_ab_$v.json 2> gpurun_out/r2h_ab_$v.err; show gpurun_out/r2h_ab_$v.json $v
done
done
unset INVPREF_LIB
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2h_pytest.log; tail -6 gpurun_out/r2h_pytest.log
date
