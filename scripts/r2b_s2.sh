#!/bin/bash
mkdir -p gpurun_out
./scripts/microbench_gather > gpurun_out/r2b_microbench_gather.txt 2>&1
cat gpurun_out/r2b_microbench_gather.txt
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cell_record" 2>&1 | tail -40 > gpurun_out/r2b_t1.log
