"""Latency / throughput of the inference API: PhysicsVAE.forward (encoder + decoder + world model + value branch,
rllib_model_torch.py:742-771) and the decoder-only pass-through of the runtime (envs/rllib_env_imitation.py:234-264), at the
batch sizes the runtime uses (1) up to a few thousand.  bf16 engine, default dims; eager calls and CUDA-graph replays.
Runs on the GPU box: python tools/bench_infer.py > gpurun_out/infer.log"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physicsvae_b200 import _abi
from tests import parity as P

cfg = P.DEFAULT
om, layers = P.oracle_model(cfg)
m = P.product_model(cfg, layers, om.state_dict(), precision="bf16", max_batch=4096)
eng = m.engine()
m.sync_weights()
ALL = _abi.PART_ENCODER | _abi.PART_DECODER | _abi.PART_WORLD | _abi.PART_VALUE
rows = []
for B in (1, 16, 256, 4096):
    obs = torch.randn(B, 2 * cfg["dsb"], device="cuda")
    z = torch.randn(B, cfg["z"], device="cuda")
    for name, call in (("forward (all four nets)", lambda: eng.forward(obs, ALL, noise=False)),
                       ("pass-through decoder (z given)", lambda: eng.forward(obs, _abi.PART_DECODER, z_in=z))):
        for _ in range(5):
            call()
        torch.cuda.synchronize()
        n0 = _abi.launch_count()
        call()
        launches = _abi.launch_count() - n0
        t0 = time.perf_counter()
        for _ in range(200):
            call()
        torch.cuda.synchronize()
        eager_us = (time.perf_counter() - t0) / 200 * 1e6
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                call()
        torch.cuda.current_stream().wait_stream(s)
        for _ in range(5):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_us = e0.elapsed_time(e1) / 200 * 1e3
        rows.append({"what": name, "batch": B, "kernel_launches": int(launches), "eager_us_per_call": round(eager_us, 1),
                     "graph_us_per_call": round(graph_us, 1), "rows_per_s_graph": round(B / (graph_us * 1e-6))})
        print(json.dumps(rows[-1]), flush=True)
# CPU reference at batch 1 (the runtime's setting): the oracle's forward on one thread
torch.set_num_threads(1)
x = torch.randn(1, 2 * cfg["dsb"])
om.latent_prior_noise = False
for _ in range(20):
    om.forward(x)
t0 = time.perf_counter()
for _ in range(200):
    om.forward(x)
print(json.dumps({"what": "oracle forward on 1 CPU thread", "batch": 1, "us_per_call": round((time.perf_counter() - t0) / 200 * 1e6, 1)}), flush=True)
