#!/bin/bash
python -m pytest tests -q -m gpu -x -k "grid or nurbs or c4 or guard" 2>&1 | tail -2
for nv in 0 1; do
for wl in c4 c4b; do
    if [ $nv = 1 ]; then export DIND_NO_BATCH_VEC=1; else unset DIND_NO_BATCH_VEC; fi
    python bench.py --workload $wl --no-cpu-baseline --only-headline --no-cache-build --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/tmp.json
    python - <<PY
import json
b=json.load(open("gpurun_out/tmp.json"))
print("novec=$nv $wl:", b["value"], "G/s", b["ms_per_step"], "ms", b["roofline"]["frac"], b["config"].get("kernel"))
PY
done
done
./scripts/microbench_gather > gpurun_out/r2b_microbench_gather.txt 2>&1; head -9 gpurun_out/r2b_microbench_gather.txt
