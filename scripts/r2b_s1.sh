#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cell_record" 2>&1 | tail -30 > gpurun_out/r2b_t1.log
tail -3 gpurun_out/r2b_t1.log
for pts in 100000000 500000000; do
  python bench.py --workload c5 --points $pts --no-cpu-baseline --only-headline --no-cache-build --steps 5 --warmup 2 2>gpurun_out/r2b_c5_$pts.err | tail -1 > gpurun_out/r2b_c5_$pts.json
  python - <<PY
import json
b=json.load(open("gpurun_out/r2b_c5_$pts.json"))
print("pts=$pts", b["value"], b["ms_per_step"], b["roofline"]["frac"], b["config"].get("kernel"), b["e2e"]["value"])
PY
done
