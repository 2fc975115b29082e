#!/bin/bash
# sweep of the peer-update kernel's knobs on N GPUs
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578"
for cfg in "1 2" "1 4" "1 8" "2 2" "2 4" "4 2" "4 1"; do
  set -- $cfg
  NB200_PEER_UNROLL=$1 NB200_PEER_CTAS_PER_SM=$2 timeout 120 $TR profiles/peer_probe.py 2>&1 | grep "^world"
done
