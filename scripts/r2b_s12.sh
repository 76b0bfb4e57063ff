#!/bin/bash
python -m pytest tests -q -m gpu -x -k "nurbs or c4 or guard or grid" 2>&1 | tail -3
for dc in 0 1; do
for wl in c4 c4b; do
    if [ $dc = 1 ]; then export DIND_NO_DEN_CACHE=1; else unset DIND_NO_DEN_CACHE; fi
    python bench.py --workload $wl --no-cpu-baseline --only-headline --no-cache-build --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/tmp.json
    python - <<PY
import json
b=json.load(open("gpurun_out/tmp.json"))
print("nocache=$dc $wl:", b["value"], "G/s", b["ms_per_step"], "ms", b["roofline"]["frac"], b["config"].get("kernel"), b["gpu_launches"])
PY
done
done
