#!/bin/bash
# N-GPU visit of round 2: peer-update parity tests at world N, the bench line (narrow peer update, default) and alternatives,
# configs[4] strong scaling.  $1 = N; $2 = list of "grid threads unroll" alternatives separated by ';' (optional)
set -u
N=${1:-2}
ALTS=${2:-}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29583"
line() { python - "$1" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('   n=%d %s ms/step %.4f  Mrays/s %.2f  e2e %.2f  update %s  loss %.5f' % (d['n_gpus'], d['scaling'], d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d.get('update',{}).get('us'), d['final_loss']))
P
}
timeout 900 python -m pytest "tests/test_gpu_peer_update.py::test_ranks_match_nccl_allreduce_plus_adam[$N]" -x -q > $OUT/r02_peer_pytest_${N}gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02_peer_pytest_${N}gpu.log
tail -3 $OUT/r02_peer_pytest_${N}gpu.log
echo "== bench configs[1], default (narrow peer update)"
timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 5 > $OUT/r02_bench_${N}gpu.json 2> $OUT/r02_bench_${N}gpu.err; line $OUT/r02_bench_${N}gpu.json
IFS=';' read -ra A <<< "$ALTS"
for cfg in "${A[@]}"; do
  set -- $cfg
  echo "== NB200_PEER_GRID=$1 NB200_PEER_THREADS=$2 NB200_PEER_UNROLL=$3"
  NB200_PEER_GRID=$1 NB200_PEER_THREADS=$2 NB200_PEER_UNROLL=$3 timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 5 > $OUT/r02_bench_${N}gpu_$1_$2_$3.json 2> $OUT/r02_bench_${N}gpu_$1_$2_$3.err
  line $OUT/r02_bench_${N}gpu_$1_$2_$3.json
done
echo "== configs[4] strong scaling (1 M rays sharded over $N ranks)"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --config 4 > $OUT/r02_c4_strong_${N}gpu.json 2> $OUT/r02_c4_strong_${N}gpu.err; line $OUT/r02_c4_strong_${N}gpu.json
grep -v "^\*\*\*\|UserWarning\|return func\|NCCL version\|warnings.warn" $OUT/r02_c4_strong_${N}gpu.err | tail -3
