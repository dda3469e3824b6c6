"""Per-layer clock breakdown of the small-batch cluster kernel (PVAE_SMALL_TRACE=1 python tools/small_trace.py), GPU box only."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity as P
from physicsvae_b200 import _abi
cfg = P.DEFAULT
om, layers = P.oracle_model(cfg)
m = P.product_model(cfg, layers, om.state_dict(), precision="bf16", max_batch=64)
eng = m.engine(); m.sync_weights()
obs = torch.randn(1, 2 * cfg["dsb"], device="cuda"); z = torch.randn(1, cfg["z"], device="cuda")
for i in range(4):
    eng.forward(obs, _abi.PART_DECODER, z_in=z); torch.cuda.synchronize(); print("---- call", i, flush=True)
