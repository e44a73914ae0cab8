#!/bin/bash
# Round 2, GPU call K: cp.async.bulk + mbarrier staging of the re-assignment kernel (A/B), GPU tests on both builds.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c=d['cluster']
print(sys.argv[2], 'cluster ms', round(c['ms'],3), c['ms_min_max'], c['samples'], 'G/s', round(c['value']/1e9,3), 'frac', round(c['roofline']['frac'],3), '| step', round(d['ms_per_step'],4))
PY
}
export INVPREF_LIB=$PWD/build/variants/libinvpref_bulk.so
timeout 900 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_configs.py tests/test_gpu_fullsize.py -m gpu -q -k "cluster" 2>&1 | tail -4
for nb in 4 23; do
  for v in main bulk; do
    if [ $v = main ]; then unset INVPREF_LIB; else export INVPREF_LIB=$PWD/build/variants/libinvpref_$v.so; fi
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config-legs --nbatch $nb > gpurun_out/r2k_${v}_$nb.json 2> gpurun_out/r2k_${v}_$nb.err; show gpurun_out/r2k_${v}_$nb.json "$v nb=$nb"
  done
done
unset INVPREF_LIB
echo "== pytest -m gpu (main)"; date
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2k_pytest.log; tail -4 gpurun_out/r2k_pytest.log
date
