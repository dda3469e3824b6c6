#!/bin/bash
# A/B of engine knobs on ONE box: bench (50 steps) + ncu launch list per setting.  Usage: gpu_ab.sh "ENV=VAL" "ENV=VAL" ...
mkdir -p gpurun_out
i=0
for kv in "$@"; do
i=$((i+1))
for rep in 1 2; do
env $kv timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ab_${i}_$rep.log 2>&1
tail -n 1 gpurun_out/ab_${i}_$rep.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$kv rep $rep: ms', round(d['ms_per_step'], 4), 'kernel', round(d['roofline']['kernel_ms_per_step'], 4), 'e2e ms', round(d['e2e']['ms_per_step'], 3))"
done
env $kv timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pvae_gemm -s 24 -c 8 --csv --log-file gpurun_out/ab_$i.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ab_ncu_$i.log 2>&1
python tools/launch_table.py gpurun_out/ab_$i.csv | cut -c1-120
done
