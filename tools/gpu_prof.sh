#!/bin/bash
# ncu --set full capture of selected GEMM launches of one world step. Usage: gpu_prof.sh TAG "<kernel regex>" SKIP COUNT
mkdir -p gpurun_out
TAG=${1:-p}; RE=${2:-pvae_gemm}; SKIP=${3:-27}; CNT=${4:-9}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_$TAG.ncu-rep
