#!/bin/bash
mkdir -p gpurun_out
for mb in 3 4 5 6; do
  DIND_CHUNK=$mb python bench.py --workload c5 --no-cpu-baseline --only-headline --no-cache-build --steps 5 --warmup 2 2>/dev/null | tail -1 > gpurun_out/tmp.json
  python - <<PY
import json
b=json.load(open("gpurun_out/tmp.json"))
print("minb=$mb", b["value"], b["ms_per_step"], b["config"].get("kernel"))
PY
done
