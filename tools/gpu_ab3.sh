#!/bin/bash
# A/B/C on one box: default library vs several build/variants/libinvpref_<v>.so, C5 bench only, alternating runs.
# usage: bash tools/gpu_ab3.sh "<v1> <v2> ..." [reps]
set -u
VS=$1; R=${2:-2}
mkdir -p gpurun_out
for r in $(seq 1 $R); do
  for lib in main $VS; do
    if [ $lib = main ]; then unset INVPREF_LIB; else export INVPREF_LIB=build/variants/libinvpref_$lib.so; fi
    timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-config-legs > gpurun_out/ab_${lib}_$r.json 2> gpurun_out/ab_${lib}_$r.err
    python - gpurun_out/ab_${lib}_$r.json $lib <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ph=d['roofline'].get('phase_ms',{})
    print(sys.argv[2].ljust(8),'step',round(d['ms_per_step'],4),'upass',round(ph.get('rows_users',0),4),'items',round(ph.get('rows_items',0),4),'dense',round(d['dense_adam']['ms_per_step'],4),'dense_upass',round(d['dense_adam']['phase_ms'].get('rows_users',0),4),'e2e',round(d['e2e']['ms_per_step'],3))
except Exception as e:
    print(sys.argv[2],'failed',e, open(sys.argv[1].replace('.json','.err')).read()[-300:])
PY
  done
done
