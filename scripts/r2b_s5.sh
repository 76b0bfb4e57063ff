#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_guard_bands.py tests/test_gpu_fullsize.py -q -m gpu -x -k "sorted or tile or plan or guard or c3 or c5" 2>&1 | tail -4
for wl in "c5 --sorted" "c3 --full --sorted" "c3 --full --sorted --gradient fused" "c1 --sorted"; do
    python bench.py --workload $wl --no-cpu-baseline --only-headline --steps 5 --warmup 2 2>/dev/null | tail -1 > gpurun_out/tmp.json
    python - <<PY
import json
b=json.load(open("gpurun_out/tmp.json"))
print("$wl:", b["value"], "Gpt/s", b["ms_per_step"], "ms", b["config"].get("kernel"), "cache_build", b["config"].get("cache_build"))
PY
done
