#!/bin/bash
# tuning sweeps of round 2 (single GPU): $1 = tag
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02s}
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu $EXTRA > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_$name.json"))
    k=d.get("kernel_us",{})
    print("%-28s ms/step=%.4f  encf=%.1f fieldf=%.1f adam=%.1f encb=%.1f fieldb=%.1f march=%.1f"%("$name",d["ms_per_step"],k.get("grid_encode_forward",0),k.get("field_forward",0),k.get("adam",0),k.get("grid_encode_backward",0),k.get("field_backward",0),k.get("march_count",0)+k.get("march_write",0)))
except Exception as e:
    print("$name failed", e)
PY
}
run base NB200_X=0
run noprio NB200_SIDE_PRIORITY=0
for c in 1 2 3 4; do run prio_adam$c NB200_ADAM_CTAS_PER_SM=$c; done
EXTRA=--no-pipeline run nopipe NB200_X=0
