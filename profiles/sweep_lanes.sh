#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for R in 1 2 4; do for L in 8 32; do
  NB200_MARCH_CHUNKS=$R NB200_MARCH_LANES=$L python -m pytest tests/test_gpu_raymarching.py tests/test_gpu_ref_ext.py -x -q -k "march" 2>&1 | tail -1
  NB200_MARCH_CHUNKS=$R NB200_MARCH_LANES=$L python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/lanes_${L}_$R.json 2>/dev/null
  python -c "
import json
d=json.load(open('$OUT/lanes_${L}_$R.json')); print('chunks=$R lanes=$L', d['ms_per_step'], d['kernel_us']['march_count'])"
done; done
