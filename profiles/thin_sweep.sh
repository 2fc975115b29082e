#!/bin/bash
# 1 GPU: the update through the (world = 1) peer kernel as a THIN grid next to the next step's march -- step time per knob set
OUT=gpurun_out
run() { python bench.py --steps 200 --warmup 10 --no-cpu --no-breakdown "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.4f  Mrays/s %.2f  e2e %.2f' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))"; }
echo "baseline k_fused_adam:"; run
echo "baseline, no pipeline:"; run --no-pipeline
for cfg in "1 256 4" "1 128 4" "2 128 4" "2 128 2" "1 64 4" "2 64 4" "4 64 4"; do
  set -- $cfg
  echo "peer kernel ctas/sm $1 threads $2 unroll $3:"
  NB200_PEER_CTAS_PER_SM=$1 NB200_PEER_THREADS=$2 NB200_PEER_UNROLL=$3 run --peer-at-1
done
