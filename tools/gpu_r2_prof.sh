#!/bin/bash
# Round-2 profiling visit: full parity suite, ncu --set full of one world step (8 GEMM launches of the captured graph), metric list of
# one VAE step, the bench line.  Usage: tools/gpu_r2_prof.sh <tag>
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|^FAILED|^ERROR|full-size|bf16 trajectory" gpurun_out/pytest_gpu_$TAG.log | tail -n 30
BENCH="python bench.py --steps 2 --warmup 1 --only-phase --sustained-seconds 0 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pvae_gemm -s 16 -c 8 -f -o gpurun_out/prof_world_$TAG $BENCH --phase world > gpurun_out/ncu_world_$TAG.log 2>&1
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:pvae -s 48 -c 32 --csv --log-file gpurun_out/launches_vae_$TAG.csv $BENCH --phase vae > gpurun_out/ncu_vae_$TAG.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:pvae -s 26 -c 11 --csv --log-file gpurun_out/launches_world_$TAG.csv $BENCH --phase world > gpurun_out/ncu_world2_$TAG.log 2>&1
PVAE_LOG_GEMM=1 timeout 300 $BENCH --phase vae 2> gpurun_out/gemm_log_vae_$TAG.log > /dev/null
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_$TAG.log 2>&1
tail -c 300 gpurun_out/bench_$TAG.log
ls -la gpurun_out/ | tail -n 12
