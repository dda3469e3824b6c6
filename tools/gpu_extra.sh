#!/bin/bash
# extra bench lines for the record: stock-PyTorch library baseline, the wide config (configs[4] dims at its per-GPU batch), configs[0] on one CPU thread
mkdir -p gpurun_out
T=${1:-x}
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/extra_${name}_$T.log 2>&1; echo "$name rc=$?"; tail -n 1 gpurun_out/extra_${name}_$T.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); r = d.get('roofline', {})
    print('   ms', round(d['ms_per_step'], 4), 'Mtr/s', round(d['value'] / 1e6, 3), 'TFLOP/s', round(r.get('achieved', 0), 1), 'frac', round(r.get('frac', 0), 3), 'whole-step TF/s', round(r.get('whole_step_tflops', 0), 1))
except Exception as e: print('   parse failed', e)
"; }
run torch_world --impl torch --steps 50 --warmup 10
run torch_vae --impl torch --steps 50 --warmup 10 --phase vae
run wide_world --config wide --batch 16384 --steps 50 --warmup 10 --no-cpu-baseline
run wide_vae --config wide --batch 16384 --steps 50 --warmup 10 --no-cpu-baseline --phase vae
run wide_world_64k --config wide --batch 65536 --steps 30 --warmup 5 --no-cpu-baseline
run torch_wide_world --impl torch --config wide --batch 16384 --steps 50 --warmup 10
run cfg0_cpu1 --impl reference --steps 20 --warmup 3 --cpu-threads 1 --cpu-sample 256
