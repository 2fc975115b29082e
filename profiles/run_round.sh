#!/bin/bash
# One GPU-box visit: parity tests, the bench line, and the ncu launch list of the same bench command.
#   $1 = tag (e.g. r01b);  $2 = "full" to also take --set full captures of the heaviest kernels
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
python bench.py --steps 30 --warmup 5 --no-graph --no-breakdown > $OUT/${TAG}_bench_nograph.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_nograph.json
CMD="python bench.py --steps 2 --warmup 3 --no-breakdown"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_launches.log 2>&1
if [ "${2:-}" = "full" ]; then
  for K in ${KERNELS:-k_field_backward k_field_forward k_grid_bwd_d3c2 k_grid_fwd_d3c2 k_march_count k_march_write k_fused_adam}; do
    ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o $OUT/${TAG}_$K -f $CMD > $OUT/${TAG}_$K.log 2>&1
  done
fi
ls -la $OUT | grep $TAG
