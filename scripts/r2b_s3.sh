#!/bin/bash
# ncu of the config-5 iid kernel over the cell-record table (1e8 points)
BENCH_ARGS="--only-headline --no-cache-build" scripts/profile.sh c5 unstructured_rec c5_rec | head -1 | cut -c1-200
python scripts/ncu_summary.py gpurun_out/c5_rec_full.ncu-rep gpurun_out/launches_c5_rec.csv > gpurun_out/r2b_c5_rec.md 2>/dev/null
ls -la gpurun_out/*.ncu-rep
