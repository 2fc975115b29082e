#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for L in 4 8 16 32; do
  NB200_MARCH_LANES=$L python -m pytest tests/test_gpu_raymarching.py tests/test_gpu_ref_ext.py -x -q -k "march" 2>&1 | tail -1
  NB200_MARCH_LANES=$L python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/lanes_$L.json 2>/dev/null
  python -c "
import json
d=json.load(open('$OUT/lanes_$L.json')); print('lanes=$L', d['ms_per_step'], d['kernel_us']['march_count'])"
done
