#!/bin/bash
# Round 2, GPU call E: bisect of the user-pass slowdown (variants), evaluator + new tests.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'lazy ms', round(d['ms_per_step'],4), 'rows_users', round(d['roofline']['phase_ms']['rows_users'],4), '| dense rows_users', round(d['dense_adam']['phase_ms']['rows_users'],4))
PY
}
for v in main nosnake nosb nosnakesb; do
  if [ $v = main ]; then unset INVPREF_LIB; else export INVPREF_LIB=$PWD/build/variants/libinvpref_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config-legs --nbatch 4 > gpurun_out/r2e_ab_$v.json 2> gpurun_out/r2e_ab_$v.err; show gpurun_out/r2e_ab_$v.json $v
done
unset INVPREF_LIB
(cd build/r1repo && timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > ../../gpurun_out/r2e_r1tree.json 2> ../../gpurun_out/r2e_r1tree.err); show gpurun_out/r2e_r1tree.json r1tree
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2e_pytest.log; tail -15 gpurun_out/r2e_pytest.log
date
