#!/bin/bash
# Round 2, final single-GPU call: GPU tests, default bench line, launch lists (C5, C2), ncu --set full captures of the
# hot kernels exported to CSV on the box (the .ncu-rep files are too big to bring back).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; date
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2z_pytest.log; tail -4 gpurun_out/r2z_pytest.log
echo "== bench default"; date
timeout 1200 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 400 gpurun_out/r2z_bench.err
echo "== reference arm"; date
timeout 900 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err
echo "== launch list c5"; date
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2z_launches_c5.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-config-legs > gpurun_out/r2z_ncu_c5_list.log 2>&1
echo "== launch list c2 (trainer epochs, graph off so that every kernel is listed)"; date
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2z_launches_c2.csv \
    python bench.py --workload c2 --steps 6 --warmup 3 --no-cpu-baseline --no-config-legs > gpurun_out/r2z_ncu_c2_list.log 2>&1
echo "== ncu full: lazy user pass + item pass (c5)"; date
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'upass_rows_staged|bwd_rows_ring' -s 16 -c 2 \
    -o /tmp/r2z_prof_c5 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config-legs --nbatch 4 > gpurun_out/r2z_ncu_c5.log 2>&1
ncu -i /tmp/r2z_prof_c5.ncu-rep --page raw --csv > gpurun_out/r2z_prof_c5_raw.csv 2>/dev/null
echo "== ncu full: re-assignment kernels (c5)"; date
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cluster_kernel -s 2 -c 1 \
    -o /tmp/r2z_prof_cl -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-config-legs --nbatch 4 > gpurun_out/r2z_ncu_cl.log 2>&1
ncu -i /tmp/r2z_prof_cl.ncu-rep --page raw --csv > gpurun_out/r2z_prof_cl_raw.csv 2>/dev/null
ls -la gpurun_out | grep r2z
date
