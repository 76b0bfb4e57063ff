#!/bin/bash
# config 3 (10^8 points) through the cell-sorted plan, per points-per-item of the tile kernel (0 = one point per thread)
run() { echo "== $*"; env "$@" python bench.py --workload c3 --full --sorted --plan-order --no-cpu-baseline --only-headline --no-cache-build --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['config']['kernel'])"; }
for t in ${TILES:-0 2 3}; do
  run DIND_C3_TILE=$t
  run DIND_C3_TILE=$t DIND_PF1=0
done
