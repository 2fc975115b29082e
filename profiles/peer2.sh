#!/bin/bash
# N-GPU visit: peer-memory update tests (N = 2 only), the update kernel alone (P2P and NVLS), then the bench line per update kind
set -u
N=${1:-2}
KINDS=${2:-"peer nvls nccl"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 500 python -m pytest tests/test_gpu_peer_update.py -x -q > $OUT/peer_pytest_$N.log 2>&1; echo "pytest rc=$?"
  grep -n "PEER_OK\|passed\|failed\|Error" $OUT/peer_pytest_$N.log | head -20
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
for NV in 0 1; do
  PROBE_NVLS=$NV timeout 120 $TR profiles/peer_probe.py 2>&1 | grep "^world\|Error\|error" | head -5 | tee -a $OUT/peer_probe_$N.txt
done
for K in $KINDS; do
  timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 5 --update $K > $OUT/${K}_bench_$N.json 2> $OUT/${K}_bench_$N.err; echo "$K bench rc=$?"
  grep -v "^\*\*\*\|UserWarning\|return func\|NCCL version" $OUT/${K}_bench_$N.err | tail -5
  python - $OUT/${K}_bench_$N.json <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], 'ms/step %.4f'%d['ms_per_step'], 'Mrays/s %.2f'%(d['value']/1e6), 'e2e %.2f'%(d['e2e']['value']/1e6), d.get('update'), 'loss', d['final_loss'])
P
done
