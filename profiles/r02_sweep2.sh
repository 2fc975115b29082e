#!/bin/bash
# tuning sweeps of round 2 (single GPU), second batch: narrow-and-deep Adam next to the march; L2 persisting window.  $1 = tag
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02s}
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu $EXTRA > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_$name.json"))
    k=d.get("kernel_us",{})
    print("%-28s ms/step=%.4f  encf=%.1f fieldf=%.1f adam=%.1f encb=%.1f fieldb=%.1f march=%.1f l2=%s"%("$name",d["ms_per_step"],k.get("grid_encode_forward",0),k.get("field_forward",0),k.get("adam",0),k.get("grid_encode_backward",0),k.get("field_backward",0),k.get("march_count",0)+k.get("march_write",0), d.get("config",{}).get("l2_persist")))
except Exception as e:
    print("$name failed", e)
PY
}
run base NB200_L2_PERSIST=0
for cfg in "56 512 4" "64 512 4" "72 512 4" "80 512 4" "88 512 4" "64 512 1" "74 512 2" "148 512 4" "148 256 4"; do
  set -- $cfg
  run deep_$1_$2_$3 NB200_L2_PERSIST=0 NB200_ADAM_GRID=$1 NB200_ADAM_THREADS=$2 NB200_ADAM_UNROLL=$3
done
