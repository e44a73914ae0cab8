#!/bin/bash
# Round 2, GPU call A: full GPU test-suite, default bench line (C5 + C2/C3/C4 legs), launch lists, one ncu --set full
# capture of the fused user pass, compute-sanitizer on the small-shape tests.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; date
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.log; tail -5 gpurun_out/r2a_pytest.log
echo "== bench"; date
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.err; head -c 1500 gpurun_out/r2a_bench.json
echo "== launch list c2"; date
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2a_launches_c2.csv \
    python bench.py --workload c2 --steps 6 --warmup 3 --no-cpu-baseline --no-config-legs > gpurun_out/r2a_ncu_c2.log 2>&1
echo "== ncu full upass c5"; date
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'upass_rows_staged|bwd_rows_ring' -s 6 -c 4 \
    -o gpurun_out/r2a_prof_c5 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config-legs --dense-adam > gpurun_out/r2a_ncu_c5.log 2>&1
echo "== ncu full c2 kernels"; date
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'upass_rows_staged|bwd_rows_ring|bwd_chunks_ring|tail_kernel|upass_chunks' -s 30 -c 5 \
    -o gpurun_out/r2a_prof_c2 -f python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-config-legs --dense-adam > gpurun_out/r2a_ncu_c2full.log 2>&1
echo "== sanitizer"; date
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py -x -q -k "matches_reference or cluster or forward" > gpurun_out/r2a_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py -x -q -k "train_step_matches_reference or cluster_matches" > gpurun_out/r2a_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2a_racecheck.log
date
