#!/bin/bash
# role timelines (clock64 stamps) of GEMM launches of the 4th eager world step -> gpurun_out/trace_<idx>.npy
# needs the debug build: nvcc ... -DPVAE_DEBUG_HOOKS -o physicsvae_b200/lib/libpvae_sm100_dbg.so
mkdir -p gpurun_out
export PVAE_LIB=$PWD/physicsvae_b200/lib/libpvae_sm100_dbg.so
for i in ${@:-24 25 26 27 28 29 30 31}; do
PVAE_TRACE_IDX=$i timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/trace_$i.log 2>&1; echo "trace $i rc=$?"
done
python tools/trace_report.py > gpurun_out/trace_report.log 2>&1
