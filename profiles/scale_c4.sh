#!/bin/bash
# BASELINE.json configs[4] (2^22 table, 1 M rays per step sharded over N ranks: strong scaling) on N GPUs
set -u
N=${1:-8}
EXTRA=${2:-}
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29580"
TAG=c4${EXTRA:+_weak}
timeout 300 $TR bench.py --gpus $N --config 4 $EXTRA --steps 20 --warmup 3 --no-cpu > $OUT/${TAG}_bench_$N.json 2> $OUT/${TAG}_bench_$N.err; echo "c4 bench rc=$?"
grep -v "^\*\*\*\|UserWarning\|return func\|NCCL version\|OMP_NUM_THREADS\|^$" $OUT/${TAG}_bench_$N.err | tail -5
python - $OUT/${TAG}_bench_$N.json <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['scaling'], 'ms/step %.4f'%d['ms_per_step'], 'Mrays/s %.2f'%(d['value']/1e6), 'e2e %.2f'%(d['e2e']['value']/1e6), d.get('update'), 'samples', d['samples_per_step'])
P
