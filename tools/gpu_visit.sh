#!/bin/bash
# One GPU visit.  Usage: gpu_visit.sh TAG [test|notest] [full|nofull] [extra env assignments...]
# test: pytest -m gpu;  bench world + vae;  ncu launch list (world);  full: ncu --set full of one world step's GEMMs.
mkdir -p gpurun_out
TAG=${1:-v}; T=${2:-test}; F=${3:-nofull}; shift 3
for kv in "$@"; do export "$kv"; done
if [ "$T" = "test" ]; then
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/pytest_gpu_$TAG.log
fi
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_$TAG.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('world ms', d['ms_per_step'], 'Mtr/s', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, 'frac', d['roofline']['frac'], 'clk', d['clocks'])
except Exception as e: print('bench parse failed', e)
"
timeout 600 python bench.py --steps 50 --warmup 10 --phase vae --no-cpu-baseline > gpurun_out/bench_vae_$TAG.log 2>&1; echo "bench vae rc=$?"; tail -n 1 gpurun_out/bench_vae_$TAG.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('vae ms', d['ms_per_step'], 'Mtr/s', d['value']/1e6, 'frac', d['roofline']['frac'])
except Exception as e: print('bench parse failed', e)
"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pvae -s 40 -c 30 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_table.py gpurun_out/launches_$TAG.csv 2>&1 | tail -n 24
if [ "$F" = "full" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pvae_gemm -s 24 -c 8 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_$TAG.ncu-rep
fi
