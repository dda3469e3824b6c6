"""Small cases for compute-sanitizer runs (compute-sanitizer --tool memcheck|synccheck|racecheck python tools/sanitize_case.py):
  1. world + VAE step in bf16 mode, ragged batch (general TMA-store epilogue, CTA pairs, PDL)
  2. world + VAE step in bf16 mode at batch 512, default dims (the lean epilogue: every hidden layer), deterministic mode on top
  3. swish nets (pre-activation kept, beta gradients), lookahead-2 rollout
  4. the small-batch cluster kernel (batch 1 and 3), stand-alone FC"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity as P
from oracle import pvae_oracle as orc
from physicsvae_b200 import rllib_model_torch as pm

for world in (True, False):
    o, p = P.step_pair(P.SMALL, 300, world, precision="bf16", out_std=0.3, cyc_coeff=0.05, n_rows=400, cursor=37)
    torch.cuda.synchronize()
    worst = max(P.rel_l2(p["grads"][k], g) for k, g in o["grads"].items())
    print("[sanitize] 1 world=%s loss %.6f (oracle %.6f) worst grad rel-L2 %.2e" % (world, p["loss"], o["loss"], worst), flush=True)
    assert abs(p["loss"] - o["loss"]) < 5e-3 * abs(o["loss"]) and worst < 0.2
for world in (True, False):
    o, p = P.step_pair(P.DEFAULT, 512, world, precision="bf16", out_std=0.3, cyc_coeff=0.05)
    eng = p["model"].engine()
    eng.set_deterministic(True)
    eng.set_cursor(0)
    (eng.world_step(512) if world else eng.vae_step(512, eps=o["eps"].cuda(), cyc_coeff=0.05))
    torch.cuda.synchronize()
    worst = max(P.rel_l2(p["grads"][k], g) for k, g in o["grads"].items())
    print("[sanitize] 2 world=%s loss %.6f (oracle %.6f) worst grad rel-L2 %.2e" % (world, p["loss"], o["loss"], worst), flush=True)
    assert abs(p["loss"] - o["loss"]) < 5e-3 * abs(o["loss"]) and worst < 0.2
for world in (True, False):
    o, p = P.step_pair(P.SMALL, 256, world, act="swish", precision="bf16", out_std=0.3, cyc_coeff=0.05)
    torch.cuda.synchronize()
    print("[sanitize] 3 swish world=%s loss %.6f (oracle %.6f)" % (world, p["loss"], o["loss"]), flush=True)
cfg, B, L, n = P.SMALL, 128, 2, 160
om, layers = P.oracle_model(cfg, seed=5, out_std=0.3)
X, Y = orc.build_transitions(orc.synthetic_episodes(3, 65, cfg["dsb"], cfg["da"], seed=11)["episodes"], num_samples=n, lookahead=L)
m = P.product_model(cfg, layers, om.state_dict(), precision="bf16", max_batch=B)
eng = m.engine()
m.sync_weights()
bufs = []
for t in range(L):
    eng.alloc_transitions(n)
    eng.ingest(torch.from_numpy(np.ascontiguousarray(X[:, t])).cuda(), torch.from_numpy(np.ascontiguousarray(Y[:, t])).float().cuda())
    bufs.append(eng.transitions)
for world in (True, False):
    m.set_learnable_task_encoder(not world); m.set_learnable_motor_decoder(not world); m.set_learnable_world_model(world)
    eng.set_cursor(9)
    loss = eng.rollout_step(B, world, bufs, n, noise=False, cyc_coeff=0.05)
    torch.cuda.synchronize()
    print("[sanitize] 3 rollout world=%s loss %.6f" % (world, float(loss[0])), flush=True)
for b in (1, 3):
    x = torch.randn(b, 2 * cfg["dsb"], device="cuda")
    lg, _ = m(input_dict={"obs": x, "obs_flat": x}, state=None, seq_lens=None)
    out, _ = m.forward_decoder(x[:, :cfg["dsb"]], torch.randn(b, cfg["z"], device="cuda"), [], None, 0)
    torch.cuda.synchronize()
    print("[sanitize] 4 small batch %d: logits %s pass-through %s" % (b, tuple(lg.shape), tuple(out.shape)), flush=True)
spec = [{"type": "fc", "hidden_size": 40, "activation": "elu", "init_weight": {"name": "normc", "std": 1.0}},
        {"type": "fc", "hidden_size": "output", "activation": "tanh", "init_weight": {"name": "normc", "std": 0.5}}]
fc = pm.FC(size_in=53, size_out=7, layers=spec).to("cuda:0")
print("[sanitize] 4 stand-alone FC", tuple(fc(torch.randn(2, 53, device="cuda")).shape, ), tuple(fc(torch.randn(40, 53, device="cuda")).shape))
torch.cuda.synchronize()
print("[sanitize] ok")
