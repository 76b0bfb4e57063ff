#!/bin/bash
BENCH_ARGS="--only-headline --no-cache-build --full --sorted --plan-order" scripts/profile.sh c3 unstructured_tile_smem c3_plan_order | head -1 | cut -c1-160
python scripts/ncu_summary.py gpurun_out/c3_plan_order_full.ncu-rep gpurun_out/launches_c3_plan_order.csv > gpurun_out/r2b_c3_plan_order.md 2>/dev/null
cp gpurun_out/launches_c3_plan_order.csv gpurun_out/r2b_c3_plan_order_launches.csv
ncu -i gpurun_out/c3_plan_order_full.ncu-rep --page source --csv --print-source sass > gpurun_out/r2b_c3_sass.csv 2>/dev/null
rm -f gpurun_out/c3_plan_order_full.ncu-rep
