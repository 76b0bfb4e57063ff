#!/bin/bash
# Produces the ncu evidence for one bench workload (run under gpurun on ONE B200):
#   gpurun_out/launches_<wl>.csv     every launch with its device time (cold-cache, serialised)
#   gpurun_out/<wl>_full.ncu-rep     --set full capture of the dominant kernel (3 launches)
# usage: [BENCH_ARGS=...] scripts/profile.sh <workload> <kernel-regex> [tag]
set -u
WL=${1:-c2}
RE=${2:-grid_pencil}
TAG=${3:-$WL}
CMD="python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-}"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 3 -c 1 -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-200
