#!/bin/bash
# What the driver does at round end, plus the profile captures: tests, smoke, bench (default args), reference arm, ncu launch list + full capture.
mkdir -p gpurun_out
TAG=${1:-final}
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke_$TAG.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.log 2>&1; echo "ref rc=$?"; tail -n 1 gpurun_out/bench_ref_$TAG.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_$TAG.log | cut -c1-2500
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench50_$TAG.log 2>&1; echo "bench50 rc=$?"
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --phase vae > gpurun_out/bench50_vae_$TAG.log 2>&1; echo "bench50 vae rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pvae -s 40 -c 30 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pvae_gemm -s 24 -c 8 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
