#!/bin/bash
# 4-GPU visit: P2P vs multicast update on configs[1] (weak), configs[4] strong
set -u
N=4
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29585"
line() { python - "$1" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('   n=%d %s ms/step %.4f  Mrays/s %.2f  e2e %.2f  update %s  loss %.5f' % (d['n_gpus'], d['scaling'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d.get('update',{}), d['final_loss']))
P
}
run() { local tag=$1; local extra=$2; shift; shift
  echo "== $tag"
  env "$@" timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 5 $extra > $OUT/r02u_${N}gpu_$tag.json 2> $OUT/r02u_${N}gpu_$tag.err
  line $OUT/r02u_${N}gpu_$tag.json || tail -5 $OUT/r02u_${N}gpu_$tag.err
}
run auto "" NB200_X=0
run peer "--update peer" NB200_X=0
run c4_strong "--config 4 --steps 20 --warmup 3" NB200_X=0
