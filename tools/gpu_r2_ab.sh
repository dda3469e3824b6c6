#!/bin/bash
# quick A/B runs of the world / VAE step on one box: step count, PDL, zig-zag walk, lean epilogue.  Usage: tools/gpu_r2_ab.sh <tag>
TAG=${1:-ab}
mkdir -p gpurun_out
one() {  # label, env..., -- bench args
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
one w_100 X=1 -- --steps 100 --warmup 10
one w_200 X=1 -- --steps 200 --warmup 20
one w_1000 X=1 -- --steps 1000 --warmup 20
one w_200_again X=1 -- --steps 200 --warmup 20
one w_nopdl PVAE_PDL=0 -- --steps 200 --warmup 20
one w_nosnake PVAE_SNAKE=0 -- --steps 200 --warmup 20
one w_nofast PVAE_FAST_EPI=0 -- --steps 200 --warmup 20
one v_200 X=1 -- --steps 200 --warmup 20 --phase vae
one v_nofast PVAE_FAST_EPI=0 -- --steps 200 --warmup 20 --phase vae
one v_nopdl PVAE_PDL=0 -- --steps 200 --warmup 20 --phase vae
