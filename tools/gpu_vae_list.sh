#!/bin/bash
# launch list of VAE-phase steps (all pvae kernels) -> gpurun_out/launches_vae_$1.csv
mkdir -p gpurun_out
TAG=${1:-v}
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pvae -s 120 -c 80 --csv --log-file gpurun_out/launches_vae_$TAG.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --phase vae > gpurun_out/ncu_launch_vae_$TAG.log 2>&1; echo "ncu vae rc=$?"
PVAE_LOG_GEMM=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --phase vae 2>&1 | grep "pvae_gemm\]" | tail -28 > gpurun_out/gemm_log_vae_$TAG.log
python tools/launch_table.py gpurun_out/launches_vae_$TAG.csv | cut -c1-110 | tail -45
