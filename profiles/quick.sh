#!/bin/bash
# quick GPU check: march / fused-step parity tests, then the bench's live stage breakdown.  $1 = tag
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-q}
python -m pytest tests/test_gpu_raymarching.py tests/test_gpu_golden.py tests/test_gpu_ref_ext.py tests/test_gpu_fused_step.py tests/test_gpu_render.py -x -q 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/${TAG}_bench.json 2>$OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print("ms/step=%.3f rays/s=%.3e e2e=%.3e samples=%d"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["samples_per_step"]))
print(d["kernel_us"])
PY
