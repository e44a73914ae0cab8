#!/bin/bash
# Source-level ncu page (SASS + stall samples) of the item pass's ring rows kernel, exported to CSV on the box.
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_rows_ring -s 8 -c 1 \
    -o /tmp/r2x_items -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config-legs --nbatch 4 > gpurun_out/r2x_ncu_items.log 2>&1
ncu -i /tmp/r2x_items.ncu-rep --page source --csv > gpurun_out/r2x_items_source.csv 2>/dev/null
ncu -i /tmp/r2x_items.ncu-rep --page raw --csv > gpurun_out/r2x_items_raw.csv 2>/dev/null
ls -la gpurun_out/r2x_items_source.csv gpurun_out/r2x_items_raw.csv; tail -3 gpurun_out/r2x_ncu_items.log
