#!/bin/bash
# 8-GPU visit on the last tree: peer-update pytest (world 8), configs[4] strong scaling with and without the split update
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_peer_update.py -m gpu -q 2>&1 | tail -3 | tee $OUT/r03p_peer_pytest_8gpu.log
LEVELS="0 8" EXTRA="--config 4 --steps 20 --warmup 3" bash profiles/r03_split.sh 8
for SL in 0 8; do cp $OUT/r03k_bench_8gpu_split$SL.json $OUT/r03p_c4_strong_8gpu_split$SL.json; done
