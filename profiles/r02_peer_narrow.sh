#!/bin/bash
# N-GPU visit: the peer-memory update as a NARROW grid (few SMs, deep loads) next to the next step's march.  $1 = N
set -u
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581"
line() { python - "$1" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('   n=%d ms/step %.4f  Mrays/s %.2f  e2e %.2f  update %s' % (d['n_gpus'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d.get('update',{}).get('us')))
P
}
for cfg in "0 256 0" "32 512 0" "32 256 0" "64 256 0" "64 512 2" "16 512 0"; do
  set -- $cfg
  echo "== NB200_PEER_GRID=$1 NB200_PEER_THREADS=$2 NB200_PEER_UNROLL=$3"
  NB200_PEER_GRID=$1 NB200_PEER_THREADS=$2 NB200_PEER_UNROLL=$3 timeout 120 $TR profiles/peer_probe.py 2>&1 | grep "^world"
  NB200_PEER_GRID=$1 NB200_PEER_THREADS=$2 NB200_PEER_UNROLL=$3 timeout 200 $TR bench.py --gpus $N --steps 60 --warmup 5 > $OUT/narrow_${N}_$1_$2_$3.json 2> $OUT/narrow_${N}_$1_$2_$3.err
  line $OUT/narrow_${N}_$1_$2_$3.json
done 2>&1 | tee $OUT/r02_peer_narrow_$N.txt
