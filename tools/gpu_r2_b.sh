#!/bin/bash
# Round 2, GPU call B: GPU test-suite (no -x), default bench line, ncu launch list + full captures exported to CSV
# on the box (the .ncu-rep files are too big to bring back).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2b_pytest.log; tail -12 gpurun_out/r2b_pytest.log
echo "== bench"; date
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 600 gpurun_out/r2b_bench.err
echo "== launch list c2"; date
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2b_launches_c2.csv \
    python bench.py --workload c2 --steps 6 --warmup 3 --no-cpu-baseline --no-config-legs > gpurun_out/r2b_ncu_c2.log 2>&1
echo "== ncu full upass c5"; date
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'upass_rows_staged' -s 4 -c 1 \
    -o /tmp/r2b_prof_c5 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config-legs > gpurun_out/r2b_ncu_c5.log 2>&1
ncu -i /tmp/r2b_prof_c5.ncu-rep --page raw --csv > gpurun_out/r2b_prof_c5_raw.csv 2>/dev/null
ncu -i /tmp/r2b_prof_c5.ncu-rep --page source --csv > gpurun_out/r2b_prof_c5_source.csv 2>/dev/null
echo "== ncu full c2 kernels"; date
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'upass_rows_staged|bwd_rows_ring|bwd_chunks_ring|tail_kernel|upass_chunks' -s 30 -c 5 \
    -o /tmp/r2b_prof_c2 -f python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-config-legs --dense-adam > gpurun_out/r2b_ncu_c2full.log 2>&1
ncu -i /tmp/r2b_prof_c2.ncu-rep --page raw --csv > gpurun_out/r2b_prof_c2_raw.csv 2>/dev/null
ncu -i /tmp/r2b_prof_c2.ncu-rep --page source --csv > gpurun_out/r2b_prof_c2_source.csv 2>/dev/null
ls -la gpurun_out | tail -20
date
