#!/bin/bash
# ncu of the config-5 iid kernel over the cell-record table at the bench's size (5e8 points)
BENCH_ARGS="--only-headline --no-cache-build --points 500000000" scripts/profile.sh c5 unstructured_rec c5 | head -1 | cut -c1-200
python scripts/ncu_summary.py gpurun_out/c5_full.ncu-rep gpurun_out/launches_c5.csv > gpurun_out/r2b_c5.md 2>/dev/null
cp gpurun_out/launches_c5.csv gpurun_out/r2b_c5_launches.csv
ncu -i gpurun_out/c5_full.ncu-rep --page source --csv --print-source sass > gpurun_out/r2b_c5_sass.csv 2>/dev/null
rm -f gpurun_out/c5_full.ncu-rep
