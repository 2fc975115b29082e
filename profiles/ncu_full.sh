#!/bin/bash
# ncu --set full captures (one warm launch each) of the kernels named in $KERNELS; $1 = tag
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-full}
CMD="python bench.py --steps 2 --warmup 3 --no-breakdown --no-graph"
for K in ${KERNELS:-k_field_backward k_field_forward k_grid_bwd_d3c2 k_march_count_seg}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o $OUT/${TAG}_$K -f $CMD > $OUT/${TAG}_$K.log 2>&1
done
ls -la $OUT | grep $TAG
