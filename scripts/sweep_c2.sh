#!/bin/bash
# Chunk-length sweep of the config-2 pencil kernel (run under gpurun on one B200).
# CHUNKS="26 32 37" scripts/sweep_c2.sh
for c in ${CHUNKS:-8 16 32 64}; do
  echo -n "chunk=$c: "
  DIND_CHUNK=$c python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['kernel'], d['roofline']['frac'])"
done
