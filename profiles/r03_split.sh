#!/bin/bash
# deferred split update: $1 = N; peer pytest at N = 2, then the bench line at several split levels
set -u
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29585"
if [ "$N" = "2" ] && [ -z "${EXTRA:-}" ]; then
  timeout 900 python -m pytest tests/test_gpu_peer_update.py -m gpu -q 2>&1 | tail -4 | tee $OUT/r03k_peer_pytest_${N}gpu.log
fi
for SL in ${LEVELS:-0 8 10 12}; do
  NB200_SPLIT_LEVEL=$SL timeout 300 $TR bench.py --gpus $N ${EXTRA:---steps 100 --warmup 5} > $OUT/r03k_bench_${N}gpu_split$SL.json 2> $OUT/r03k_bench_${N}gpu_split$SL.err
  python - $OUT/r03k_bench_${N}gpu_split$SL.json $SL <<'P'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); ok=True; print('split %s: n=%d ms/step %.4f  Mrays/s %.2f  e2e %.2f  update %s us (%s) loss %.5f' % (sys.argv[2], d['n_gpus'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d.get('update',{}).get('us'), d.get('update',{}).get('kind'), d['final_loss']))
if not ok: print('split', sys.argv[2], 'FAILED'); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
P
done
