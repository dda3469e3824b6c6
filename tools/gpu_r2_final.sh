#!/bin/bash
# What the driver runs at round end, in one visit: the GPU suite, smoke(), the default bench line, the reference arm.
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -n 12 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
( time timeout 900 python bench.py ) > gpurun_out/bench_$TAG.log 2>&1
grep -E "^real" gpurun_out/bench_$TAG.log
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_$TAG.log") if l.startswith("{")][-1])
r = d["roofline"]
print("value %.1f M tr/s  %.4f ms/step  frac %.3f (burst %.3f sustained %.3f)  e2e %.1f M  resident %.1f M  launches %s" % (
    d["value"] / 1e6, d["ms_per_step"], r["frac"], r.get("frac_burst", 0), r.get("frac_sustained", 0), d["e2e"]["value"] / 1e6,
    d["e2e_resident"]["value"] / 1e6, d["gpu_launches"]))
v = d["phases"]["vae"]
print("vae %.1f M tr/s %.4f ms  burst %.3f" % (v["value"] / 1e6, v["ms_per_step"], v["roofline"].get("frac_burst", 0)))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "clocks", d["clocks"])
PY
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_$TAG.log 2>&1
tail -c 700 gpurun_out/bench_ref_$TAG.log
