#!/bin/bash
# rays-per-warp sweep of the march count kernel (NB200_MARCH_RPW), stage times from bench.py's live breakdown
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_raymarching.py tests/test_gpu_golden.py tests/test_gpu_ref_ext.py tests/test_gpu_fused_step.py -x -q 2>&1 | tail -5
for R in 32 16 8 4 2 0; do
  NB200_MARCH_RPW=$R python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/sweep_rpw_$R.json 2>$OUT/sweep_rpw_$R.err
  python - <<PY
import json
d=json.load(open("$OUT/sweep_rpw_$R.json"))
print("rpw=$R ms/step=%.3f"%d["ms_per_step"], {k:d["kernel_us"][k] for k in ("march_count","march_write")})
PY
done
