#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_gridencoder.py tests/test_gpu_fused_step.py -x -q 2>&1 | tail -3
for H in 16 12 8 6 4 2; do
  NB200_GE_AGG_MAXHEADS=$H python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/sweep_agg_$H.json 2>$OUT/sweep_agg_$H.err
  python - <<PY
import json
d=json.load(open("$OUT/sweep_agg_$H.json"))
print("maxheads=$H ms/step=%.3f"%d["ms_per_step"], {k:d["kernel_us"][k] for k in ("grid_encode_backward","adam")})
PY
done
