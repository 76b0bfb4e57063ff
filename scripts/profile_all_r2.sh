#!/bin/bash
# ncu evidence for every bench configuration at the bench's sizes (run under gpurun on ONE B200, ~9 min): launch list +
# one --set full capture of the dominant kernel each, summarised on the box by scripts/ncu_summary.py into
# gpurun_out/r2_<key>.md (copied to profiles/ afterwards); the .ncu-rep files are deleted (gpurun_out is capped at 64 MiB).
C="--only-headline --no-cache-build"
one() {   # workload, kernel regex, key, bench args
  BENCH_ARGS="$C $4" scripts/profile.sh "$1" "$2" "$3" | head -1 | cut -c1-120
  python scripts/ncu_summary.py gpurun_out/${3}_full.ncu-rep gpurun_out/launches_${3}.csv > gpurun_out/r2_${3}.md 2>/dev/null
  cp gpurun_out/launches_${3}.csv gpurun_out/r2_${3}_launches.csv 2>/dev/null
  rm -f gpurun_out/${3}_full.ncu-rep gpurun_out/launches_${3}.csv gpurun_out/ncu_list_${3}.log gpurun_out/ncu_full_${3}.log gpurun_out/plain_${3}.log
}
one c2 grid_pencil c2 ""
one c1 unstructured_aos c1 ""
one c3 unstructured_aos c3 "--full"
one c3 unstructured_tile c3_sorted "--full --sorted"
one c3 unstructured_tile c3_plan_order "--full --sorted --plan-order"
one c3 gradient_tile c3_gradient_plan_order "--full --sorted --plan-order --gradient fused"
one c4 grid_batch c4 ""
one c4 grid_slice c4_slice ""
one c5 unstructured_aos c5 "--points 500000000"
one c5 linear_tile c5_sorted "--sorted"
one c5 linear_tile c5_plan_order "--sorted --plan-order"
one c5 plan_unbucket c5_unbucket "--sorted"
ls -la gpurun_out | head -40
