#!/bin/bash
# Round 2, GPU call F: ncu --set full of the dense-leg user pass, round-1 tree vs current tree, same launch.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(cd build/r1repo && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'upass_rows_staged' -s 2 -c 1 -o /tmp/r2f_r1 -f python bench.py --steps 2 --warmup 2 --no-cpu-baseline --nbatch 2 > ../../gpurun_out/r2f_r1.log 2>&1)
ncu -i /tmp/r2f_r1.ncu-rep --page raw --csv > gpurun_out/r2f_r1_raw.csv 2>/dev/null
ncu -i /tmp/r2f_r1.ncu-rep --page source --csv > gpurun_out/r2f_r1_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'upass_rows_staged' -s 2 -c 1 -o /tmp/r2f_cur -f python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-config-legs --nbatch 2 > gpurun_out/r2f_cur.log 2>&1
ncu -i /tmp/r2f_cur.ncu-rep --page raw --csv > gpurun_out/r2f_cur_raw.csv 2>/dev/null
ncu -i /tmp/r2f_cur.ncu-rep --page source --csv > gpurun_out/r2f_cur_source.csv 2>/dev/null
timeout 900 python -m pytest tests/test_gpu_trainer.py -m gpu -q 2>&1 | tail -5
ls -la gpurun_out | grep r2f
