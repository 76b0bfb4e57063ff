#!/bin/bash
for pf in 1 2 3 4 6; do
    DIND_PF1=$pf python bench.py --workload c4 --no-cpu-baseline --only-headline --no-cache-build --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/tmp.json
    python - <<PY
import json
b=json.load(open("gpurun_out/tmp.json"))
print("pf=$pf c4:", b["value"], "G/s", b["ms_per_step"], "ms", b["roofline"]["frac"])
PY
done
