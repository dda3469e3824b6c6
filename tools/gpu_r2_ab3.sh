#!/bin/bash
# A/B of the L2 operand prefetch distance (PVAE_PREFETCH: 0 = off, -1 = auto, n = units ahead) on one box, both phases.
TAG=${1:-ab3}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 300 -k "lean_epilogue or full_size" 2>&1 | tail -4
one() {
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --only-phase --sustained-seconds 0 --no-cpu-baseline "$@" > gpurun_out/ab_${label}_$TAG.log 2>&1
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/ab_${label}_$TAG.log") if l.startswith("{")][-1])
    print("%-28s %.4f ms/step  kernels %.4f ms  %s W" % ("$label", d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["clocks"]["power_w_max"]))
except Exception as e:
    print("$label FAILED", e)
PY
}
for rep in 1 2; do
for pf in 0 -1 1 2 3; do
one w_pf${pf}_$rep PVAE_PREFETCH=$pf -- --steps 200 --warmup 20
one v_pf${pf}_$rep PVAE_PREFETCH=$pf -- --steps 200 --warmup 20 --phase vae
done
done
