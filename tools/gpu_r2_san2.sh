#!/bin/bash
# compute-sanitizer over the item pass with staged own rows (ring rows kernel: Adam and export epilogues, lazy / dense,
# simulated-rank push exchange) on the small-shape tests.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== memcheck"; date
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_lazy.py tests/test_gpu_parallel.py \
    -x -q -k "train_step_matches_reference or peer_memory_paths or lazy_equals_dense or all_geometries" > gpurun_out/r2s2_memcheck.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2s2_memcheck.log
echo "== racecheck"; date
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_parallel.py \
    -x -q -k "(train_step_matches_reference and coat) or peer_memory_paths" > gpurun_out/r2s2_racecheck.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2s2_racecheck.log
date
