#!/bin/bash
# exchange-kernel variants alone on N GPUs.  Usage: [VARIANTS="0:64 1:64 ..."] tools/gpu_r2_xchg.sh N TAG      (bulk:ctas)
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
for v in ${VARIANTS:-0:64 1:64 0:8 1:8 1:16 1:32}; do
  PVAE_SYMM_BULK=${v%%:*} PVAE_SYMM_CTAS=${v##*:} timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/bench_exchange.py 6 24 2>&1 | grep -E '^\{|rror|Traceback|trap|timed out' | tee -a gpurun_out/xchg_n${N}_$TAG.log
done
