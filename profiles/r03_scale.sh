#!/bin/bash
# final-tree multi-GPU line: $1 = N (2 | 4 | 8); at N = 2 also the peer-update pytest
set -u
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29585"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_peer_update.py -m gpu -q 2>&1 | tail -3 | tee $OUT/r03_peer_pytest_${N}gpu.log
fi
timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 5 > $OUT/r03_bench_${N}gpu.json 2> $OUT/r03_bench_${N}gpu.err
python - $OUT/r03_bench_${N}gpu.json <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('n=%d %s ms/step %.4f  Mrays/s %.2f  e2e %.2f  update %s' % (d['n_gpus'], d['scaling'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d.get('update',{})))
P
