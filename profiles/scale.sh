#!/bin/bash
# N-GPU bench lines, one per update kind (peer = P2P loads/stores, nvls = NVSwitch multicast, nccl = all-reduce + Adam)
set -u
N=${1:-8}
KINDS=${2:-"peer nvls nccl"}
STEPS=${3:-100}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
for K in $KINDS; do
  timeout 240 $TR bench.py --gpus $N --steps $STEPS --warmup 5 --update $K > $OUT/${K}_bench_$N.json 2> $OUT/${K}_bench_$N.err; echo "$K bench rc=$?"
  grep -v "^\*\*\*\|UserWarning\|return func\|NCCL version\|OMP_NUM_THREADS\|^$" $OUT/${K}_bench_$N.err | tail -5
  python - $OUT/${K}_bench_$N.json <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], 'ms/step %.4f'%d['ms_per_step'], 'Mrays/s %.2f'%(d['value']/1e6), 'e2e %.2f'%(d['e2e']['value']/1e6), d.get('update'), 'loss', d['final_loss'])
P
done
