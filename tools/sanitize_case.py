"""One small world + VAE step in bf16 mode (TMA-store epilogue, CTA pairs, PDL) for compute-sanitizer runs:
  compute-sanitizer --tool memcheck|synccheck|racecheck python tools/sanitize_case.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity as P

for world in (True, False):
    o, p = P.step_pair(P.SMALL, 300, world, precision="bf16", out_std=0.3, cyc_coeff=0.05, n_rows=400, cursor=37)
    torch.cuda.synchronize()
    worst = max(P.rel_l2(p["grads"][k], g) for k, g in o["grads"].items())
    print("[sanitize] world=%s loss %.6f (oracle %.6f) worst grad rel-L2 %.2e" % (world, p["loss"], o["loss"], worst), flush=True)
    assert abs(p["loss"] - o["loss"]) < 5e-3 * abs(o["loss"]) and worst < 0.2
print("[sanitize] ok")
