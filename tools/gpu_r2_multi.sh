#!/bin/bash
# Multi-GPU visit: bench.py under torchrun with the three gradient-exchange paths.  Usage: tools/gpu_r2_multi.sh <ngpus> <tag> [extra bench args]
N=${1:-2}; TAG=${2:-m}; shift; shift
mkdir -p gpurun_out
run() {   # name, env...
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline "${EXTRA[@]}" > gpurun_out/bench_n${N}_${name}_$TAG.log 2>&1
  echo "rc=$? $name"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n${N}_${name}_$TAG.log") if l.startswith("{")][-1])
    print("  $name: world %.4f ms/step (%.1f M tr/s)  vae %s  allreduce %s  dp_check %s" % (d["ms_per_step"], d["value"] / 1e6,
          ("%.4f ms" % d["phases"]["vae"]["ms_per_step"]) if d.get("phases", {}).get("vae") else "-", d.get("allreduce"), d.get("dp_check")))
    print("    e2e %.1f M tr/s (%.3f ms/step)  resident %.1f M tr/s  numa %s" % (d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e_resident"]["value"] / 1e6, d.get("numa")))
    for k, v in (d.get("configs") or {}).items():
        print("    %s: %.4f ms/step %.1f M tr/s" % (k, v["ms_per_step"], v["value"] / 1e6))
except Exception as e:
    print("  $name: FAILED", e)
    import subprocess; print(subprocess.run(["tail", "-n", "25", "gpurun_out/bench_n${N}_${name}_$TAG.log"], capture_output=True, text=True).stdout)
PY
}
EXTRA=("$@")
nvidia-smi topo -m > gpurun_out/topo_n${N}_$TAG.log 2>&1
(nproc; lscpu | grep -i -E "numa|socket|model name"; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null) >> gpurun_out/topo_n${N}_$TAG.log 2>&1
MODES=${MODES:-"symm nccl multimem"}
for m in $MODES; do
  case $m in
    symm) run symm PVAE_SYMM_AR=1 NCCL_DEBUG=WARN ;;
    nccl) run nccl PVAE_SYMM_AR=0 NCCL_DEBUG=WARN ;;
    overlap) run overlap PVAE_SYMM_AR=1 PVAE_OVERLAP=1 NCCL_DEBUG=WARN ;;
    multimem) run multimem PVAE_SYMM_AR=1 PVAE_SYMM_MULTIMEM=1 NCCL_DEBUG=WARN ;;
    nonuma) run nonuma PVAE_SYMM_AR=1 PVAE_NUMA_BIND=0 NCCL_DEBUG=WARN ;;
  esac
done
