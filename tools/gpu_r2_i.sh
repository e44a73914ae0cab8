#!/bin/bash
# Round 2, GPU call I: item-pass ring depth 4 / 6 / 8 (C5 and the dataset-scale configs).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'lazy ms', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline']['phase_ms'].items() if x>0.02})
PY
}
for v in main ring6 ring8; do
  if [ $v = main ]; then unset INVPREF_LIB; else export INVPREF_LIB=$PWD/build/variants/libinvpref_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config-legs --nbatch 6 > gpurun_out/r2i_ab_$v.json 2> gpurun_out/r2i_ab_$v.err; show gpurun_out/r2i_ab_$v.json $v
  for wl in c2 c4; do
    timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-config-legs --nbatch 3 > gpurun_out/r2i_ab_${v}_$wl.json 2> gpurun_out/r2i_ab_${v}_$wl.err; show gpurun_out/r2i_ab_${v}_$wl.json "$v $wl"
  done
done
date
