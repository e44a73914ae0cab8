#!/bin/bash
# Round 2, GPU call C: GPU test-suite, A/B of the kernel-arithmetic variants on the C5 step.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2c_pytest.log; tail -8 gpurun_out/r2c_pytest.log
for v in main r1math nodist noftz; do
  echo "== A/B $v"; date
  if [ $v = main ]; then unset INVPREF_LIB; else export INVPREF_LIB=$PWD/build/variants/libinvpref_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config-legs --nbatch 4 > gpurun_out/r2c_ab_$v.json 2> gpurun_out/r2c_ab_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2c_ab_$v.json').read().strip().splitlines()[-1])
print('$v', 'lazy ms', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline']['phase_ms'].items() if x>0.05}, 'dense rows_users', round(d['dense_adam']['phase_ms']['rows_users'],4), 'sweep', round(d['dense_adam']['phase_ms']['sweep_users'],4), 'cluster ms', round(d['cluster']['ms'],3))
PY
done
unset INVPREF_LIB
date
