#!/bin/bash
# ncu evidence for the kernels changed in the second half of round 2, at the bench's sizes (run under gpurun on ONE B200):
# launch list + one --set full capture of the dominant kernel each -> gpurun_out/r2b_<key>.md (copied to profiles/).
C="--only-headline --no-cache-build"
one() {   # workload, kernel regex, key, bench args
  BENCH_ARGS="$C $4" scripts/profile.sh "$1" "$2" "$3" | head -1 | cut -c1-120
  python scripts/ncu_summary.py gpurun_out/${3}_full.ncu-rep gpurun_out/launches_${3}.csv > gpurun_out/r2b_${3}.md 2>/dev/null
  cp gpurun_out/launches_${3}.csv gpurun_out/r2b_${3}_launches.csv 2>/dev/null
  rm -f gpurun_out/${3}_full.ncu-rep gpurun_out/launches_${3}.csv gpurun_out/ncu_list_${3}.log gpurun_out/ncu_full_${3}.log gpurun_out/plain_${3}.log
}
one c3 unstructured_tile_smem c3_plan_order "--full --sorted --plan-order"
one c3 unstructured_tile_smem c3_sorted "--full --sorted"
one c3 gradient_tile c3_gradient_plan_order "--full --sorted --plan-order --gradient fused"
one c4 grid_batch c4 ""
one c5 linear_tile c5_sorted "--sorted"
ls gpurun_out | grep r2b_ | head -30
