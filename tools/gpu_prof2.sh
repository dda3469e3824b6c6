#!/bin/bash
# full ncu capture of one world step's GEMMs + launch list of one VAE step
mkdir -p gpurun_out
TAG=${1:-p}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pvae_gemm -s 27 -c 9 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pvae -s 130 -c 80 --csv --log-file gpurun_out/launches_vae_$TAG.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --phase vae > gpurun_out/ncu_launch_vae_$TAG.log 2>&1; echo "ncu vae launches rc=$?"
timeout 300 python -m pytest tests -m gpu -q -x -k "fused_adam" -p no:cacheprovider 2>&1 | tail -n 3
