#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_guard_bands.py -q -m gpu -x -k "shared_memory_staging or tile_kernel_variants or guard" 2>&1 | tail -2
for sm in 0 1; do
for wl in "--sorted --plan-order --gradient fused" "--sorted --gradient fused"; do
    DIND_TILE_SMEM=$sm python bench.py --workload c3 --full $wl --no-cpu-baseline --only-headline --no-cache-build --steps 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/tmp.json
    python - <<PY
import json
b=json.load(open("gpurun_out/tmp.json"))
print("smem=$sm $wl:", b["value"], "G/s", b["ms_per_step"], "ms", b["config"].get("kernel"))
PY
done
done
