#!/bin/bash
# One GPU visit: tests, smoke, bench, ncu launch list + full capture of the GEMM kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" 
tail -n 25 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 --phase vae --no-cpu-baseline > gpurun_out/bench_vae_$TAG.log 2>&1; echo "bench vae rc=$?"; tail -n 2 gpurun_out/bench_vae_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pvae_gemm -s 24 -c 8 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -n 12
