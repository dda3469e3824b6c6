#!/bin/bash
# launch-list timings of one world step under different engine knobs.  Usage: gpu_knobs.sh "ENV=VAL ..." ["ENV=VAL ..."]...
mkdir -p gpurun_out
i=0
for kv in "$@"; do
i=$((i+1))
env $kv timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pvae_gemm -s 24 -c 8 --csv --log-file gpurun_out/knob_$i.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/knob_$i.log 2>&1
echo "== $kv"; python tools/launch_table.py gpurun_out/knob_$i.csv | cut -c1-100
done
