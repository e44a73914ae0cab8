#!/bin/bash
# Round 2: compute-sanitizer over the kernels added late in the round -- hand-written radix sort + scans (sort.cuh),
# warp-aggregated touched bitmap (write_segments_kernel), the register-counter env histogram -- through the tests that
# exercise them (small shapes).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SEL='build_segments_bit_exact or (radix_sort_edges and not 300001 and not 100000) or (train_step_matches_reference and coat) or epoch_cluster_stat'
echo "== memcheck"; date
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_trainer.py \
    -x -q -k "$SEL" > gpurun_out/r2s_memcheck.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/r2s_memcheck.log
echo "== racecheck"; date
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py \
    -x -q -k "build_segments_bit_exact or (radix_sort_edges and not 300001 and not 100000)" > gpurun_out/r2s_racecheck.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/r2s_racecheck.log
date
