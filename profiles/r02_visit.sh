#!/bin/bash
# One single-GPU box visit of round 2.  $1 = tag; $2 = what: "tests" | "bench" | "all" (default all); $3 = pytest selection (optional)
set -u
TAG=${1:-r02}
WHAT=${2:-all}
SEL=${3:-tests}
OUT=gpurun_out
mkdir -p $OUT
if [ "$WHAT" = "tests" ] || [ "$WHAT" = "all" ]; then
  timeout 1500 python -m pytest $SEL -m gpu -q --timeout=600 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
  tail -40 $OUT/${TAG}_pytest.log
fi
if [ "$WHAT" = "bench" ] || [ "$WHAT" = "all" ]; then
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
  tail -3 $OUT/${TAG}_bench.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench.json"))
    print("ms/step=%.3f rays/s=%.3e e2e=%.3e samples=%d"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["samples_per_step"]))
    print(d.get("kernel_us"))
except Exception as e:
    print("bench parse failed", e)
PY
fi
