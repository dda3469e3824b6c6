#!/bin/bash
# role timelines (clock64 stamps) of the nine GEMM launches of the 4th eager world step -> gpurun_out/trace_<idx>.npy
mkdir -p gpurun_out
for i in 27 28 29 30 31 32 33 34 35; do
PVAE_TRACE_IDX=$i timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/trace_$i.log 2>&1; echo "trace $i rc=$?"
done
