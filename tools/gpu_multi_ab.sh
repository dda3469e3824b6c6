#!/bin/bash
# N-GPU A/B of environment settings through torchrun.  Usage: gpu_multi_ab.sh N "ENV=VAL" "ENV=VAL" ...
mkdir -p gpurun_out
N=$1; shift
i=0
for kv in "$@"; do
i=$((i+1))
env $kv timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+i)) bench.py --gpus $N --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/mab_${N}_$i.log 2>&1; echo "$kv rc=$?"
grep -v "^W\|^\*\*\*\|NCCL version" gpurun_out/mab_${N}_$i.log | tail -n 1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('   ms', round(d['ms_per_step'], 4), 'Mtr/s', round(d['value'] / 1e6, 2), 'kernel', round(d['roofline']['kernel_ms_per_step'], 4), 'pool', d.get('nccl_user_buffers'))
except Exception as e: print('   parse failed', e)
"
done
