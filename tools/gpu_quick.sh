#!/bin/bash
# Quick GPU visit: GPU tests (optional), bench world+vae, ncu launch list.  Usage: gpu_quick.sh TAG [notest]
mkdir -p gpurun_out
TAG=${1:-q}
if [ "$2" != "notest" ]; then
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/pytest_gpu_$TAG.log
fi
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -n 2 gpurun_out/bench_$TAG.log | cut -c 1-1800
timeout 600 python bench.py --steps 30 --warmup 5 --phase vae --no-cpu-baseline > gpurun_out/bench_vae_$TAG.log 2>&1; echo "bench vae rc=$?"; tail -n 1 gpurun_out/bench_vae_$TAG.log | cut -c 1-300
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pvae -s 40 -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
