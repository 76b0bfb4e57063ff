#!/bin/bash
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
for wl in "c5 --sorted" "c1 --sorted"; do
  python bench.py --workload $wl --no-cpu-baseline --only-headline --steps 5 --warmup 2 2>/dev/null | tail -1 > gpurun_out/tmp.json
  python - <<PY
import json
b=json.load(open("gpurun_out/tmp.json"))
print("$wl:", b["value"], "Gpt/s", b["ms_per_step"], "ms", "cache_build", b["config"].get("cache_build")["ms"])
PY
done
