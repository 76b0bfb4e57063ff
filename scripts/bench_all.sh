#!/bin/bash
# Every bench configuration that DESIGN.md §6 quotes, one JSON line each -> gpurun_out/bench_all.jsonl
# (run under gpurun on ONE B200; ~3 min).
mkdir -p gpurun_out
: > gpurun_out/bench_all.jsonl
run() { echo "== $*" >&2; python bench.py "$@" 2>/dev/null | tail -1 | sed "s/^/{\"args\": \"$*\", \"line\": /; s/$/}/" >> gpurun_out/bench_all.jsonl; }
run --workload c2
run --workload c1 --no-cpu-baseline
run --workload c1 --sorted --no-cpu-baseline
run --workload c3
run --workload c3 --sorted --no-cpu-baseline
run --workload c3 --sorted --plan-order --no-cpu-baseline
run --workload c3 --gradient fused --no-cpu-baseline
run --workload c3 --gradient separate --no-cpu-baseline
run --workload c3 --gradient fused --sorted --no-cpu-baseline
run --workload c4 --no-cpu-baseline
run --workload c4b --no-cpu-baseline
run --workload c5
run --workload c5 --sorted --no-cpu-baseline
run --workload c5 --sorted --plan-order --no-cpu-baseline
python - <<'PY'
import json
for l in open("gpurun_out/bench_all.jsonl"):
    d = json.loads(l); b = d["line"]
    cb = b["config"].get("cache_build") or {}
    print(f'{d["args"]:52s} {b["value"]:9.3f} Gpt/s {b["ms_per_step"]:9.4f} ms  frac {b["roofline"]["frac"]:.3f}  e2e {b["e2e"]["value"]:8.3f}  cpu {(b.get("cpu_baseline") or {}).get("value")}  cache_build_ms {cb.get("ms")}  {b["config"]["kernel"]}')
PY
