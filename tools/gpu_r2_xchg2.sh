#!/bin/bash
# In-step effect of the exchange variants on N GPUs: world phase and the wide world phase, short runs.  Usage: tools/gpu_r2_xchg2.sh N TAG
N=${1:-2}; TAG=${2:-y}
mkdir -p gpurun_out
run() {   # label, bench args..., -- env...
  label=$1; shift
  args=(); while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline --only-phase --sustained-seconds 0 "${args[@]}" > gpurun_out/x2_n${N}_${label}_$TAG.log 2>&1
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/x2_n${N}_${label}_$TAG.log") if l.startswith("{")][-1])
    print("%-26s %.4f ms/step  %.1f M tr/s  exchange alone %s us  dp_check %s  overlapped %s" % ("$label", d["ms_per_step"], d["value"] / 1e6,
          (d["allreduce"].get("exchange") or {}).get("us_per_call"), d["dp_check"]["ok"], d["allreduce"]["overlapped_with_backward"]))
except Exception as e:
    print("$label FAILED", e)
    import subprocess; print(subprocess.run(["tail", "-n", "15", "gpurun_out/x2_n${N}_${label}_$TAG.log"], capture_output=True, text=True).stdout)
PY
}
IFS=';' read -ra CFGS <<< "${CFGS:-world;wide --config wide --batch 16384;vae --phase vae}"
for cfg in "${CFGS[@]}"; do
  set -- $cfg; name=$1; shift
  for m in ${MODES:-regs bulk bulk_ovl bulk_ovl4}; do
    case $m in
      regs) run ${name}_regs "$@" -- PVAE_SYMM_BULK=0 ;;
      bulk) run ${name}_bulk "$@" -- PVAE_SYMM_BULK=1 ;;
      bulk_ovl) run ${name}_bulk_ovl "$@" -- PVAE_SYMM_BULK=1 PVAE_OVERLAP=1 ;;
      bulk_ovl4) run ${name}_bulk_ovl4 "$@" -- PVAE_SYMM_BULK=1 PVAE_OVERLAP=1 PVAE_OVERLAP_SMS=4 ;;
      nccl) run ${name}_nccl "$@" -- PVAE_SYMM_AR=0 ;;
    esac
  done
done
