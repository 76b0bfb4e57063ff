#!/bin/bash
# config 5 through the cell-sorted plan: tile kernel (points per item) vs the one-point kernel (DIND_C3_TILE=0)
run() { echo "== $* $ORDER"; env "$@" python bench.py --workload c5 $FULL --sorted $ORDER --no-cpu-baseline --only-headline --no-cache-build --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['config']['kernel'])"; }
for ORDER in "--plan-order" ""; do
for t in ${TILES:-2 4}; do
  run DIND_C3_TILE=$t
done
done
