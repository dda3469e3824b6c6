#!/bin/bash
# N-GPU bench through torchrun exactly as the driver launches it.  Usage: gpu_multi.sh N [extra bench args]
mkdir -p gpurun_out
N=${1:-2}; shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 10 --no-cpu-baseline "$@" > gpurun_out/bench_n$N.log 2>&1; echo "bench N=$N rc=$?"; tail -n 2 gpurun_out/bench_n$N.log | cut -c 1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 10 --no-cpu-baseline --phase vae "$@" > gpurun_out/bench_vae_n$N.log 2>&1; echo "bench vae N=$N rc=$?"; tail -n 1 gpurun_out/bench_vae_n$N.log | cut -c 1-400
timeout 300 python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -n 1 gpurun_out/bench_ref.log | cut -c 1-600
