#!/bin/bash
# 8-GPU visit: update variants of the bench line (configs[1], weak).  $1 = N
set -u
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29585"
line() { python - "$1" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('   n=%d %s ms/step %.4f  Mrays/s %.2f  e2e %.2f  update %s  loss %.5f' % (d['n_gpus'], d['scaling'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d.get('update',{}).get('us'), d['final_loss']))
P
}
run() { # tag, extra args, env...
  local tag=$1; local extra=$2; shift; shift
  echo "== $tag"
  env "$@" timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 5 $extra > $OUT/r02t_${N}gpu_$tag.json 2> $OUT/r02t_${N}gpu_$tag.err
  line $OUT/r02t_${N}gpu_$tag.json
}
run peer "" NB200_X=0
run peer_split "" NB200_SPLIT_LEVEL=10
run nvls "--update nvls" NB200_X=0
run nvls_u4 "--update nvls" NB200_PEER_UNROLL=4
run nvls_wide "--update nvls" NB200_PEER_GRID=296 NB200_PEER_THREADS=256 NB200_PEER_UNROLL=4
