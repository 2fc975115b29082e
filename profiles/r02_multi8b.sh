#!/bin/bash
# 8-GPU visit: multicast update shapes on configs[1] (weak) + configs[3] (LGIE editing step) at 8 GPUs
set -u
N=8
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29585"
line() { python - "$1" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('   n=%d %s ms/step %.4f  Mrays/s %.2f  e2e %.2f  update %s  loss %.5f' % (d['n_gpus'], d['scaling'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d.get('update',{}).get('us'), d['final_loss']))
P
}
run() { local tag=$1; local extra=$2; shift; shift
  echo "== $tag"
  env "$@" timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 5 $extra > $OUT/r03b_${N}gpu_$tag.json 2> $OUT/r03b_${N}gpu_$tag.err
  line $OUT/r03b_${N}gpu_$tag.json || tail -5 $OUT/r03b_${N}gpu_$tag.err
}
run default "" NB200_X=0
run g32_u4 "" NB200_PEER_GRID=32 NB200_PEER_THREADS=512 NB200_PEER_UNROLL=4
run g48_u2 "" NB200_PEER_GRID=48 NB200_PEER_THREADS=512 NB200_PEER_UNROLL=2
run g48_u4 "" NB200_PEER_GRID=48 NB200_PEER_THREADS=512 NB200_PEER_UNROLL=4
run g96_u2 "" NB200_PEER_GRID=96 NB200_PEER_THREADS=512 NB200_PEER_UNROLL=2
run c3_edit "--config 3 --steps 50" NB200_X=0
