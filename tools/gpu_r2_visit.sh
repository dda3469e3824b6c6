#!/bin/bash
# Round-2 GPU visit: parity tests, the bench line, lean-epilogue A/B, launch lists of both phases.  Usage: tools/gpu_r2_visit.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$TAG.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -n 25 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_$TAG.log 2>&1
tail -c 600 gpurun_out/bench_$TAG.log
for fast in 1 0; do
  for ph in world vae; do
    PVAE_FAST_EPI=$fast timeout 300 python bench.py --steps 100 --warmup 10 --phase $ph --only-phase --sustained-seconds 0 --no-cpu-baseline \
      > gpurun_out/ab_fast${fast}_${ph}_$TAG.log 2>&1
    python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/ab_fast${fast}_${ph}_$TAG.log") if l.startswith("{")][-1])
    print("fast=$fast $ph: %.4f ms/step, kernels %.4f ms, %.1f TFLOP/s" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["achieved"]))
except Exception as e:
    print("fast=$fast $ph: FAILED", e)
PY
  done
done
