#!/bin/bash
# ncu captures of the bench command (run under gpurun on one B200).  Numbers printed by a run under ncu are never
# bench values; only the per-launch device times / counters in the reports are used.
#   $1 = tag (e.g. r01)
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-breakdown"
# every launch with its device time: skip the first 3 (warm-up) steps' launches roughly, keep 2 steps' worth
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_launches.log 2>&1
# full captures of the heaviest kernels, one launch each, taken in a warm step
for K in k_field_backward k_field_forward k_grid_bwd_d3c2 k_grid_fwd_d3c2 k_march_count_seg k_fused_adam; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o $OUT/${TAG}_$K -f $CMD > $OUT/${TAG}_$K.log 2>&1
done
ls -la $OUT | grep $TAG
