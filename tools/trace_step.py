"""Role timeline (clock64 stamps of producer / MMA issuer / epilogue, per work unit) of ONE GEMM launch of a training step.
Needs the debug build:  nvcc ... -DPVAE_DEBUG_HOOKS -o physicsvae_b200/lib/libpvae_sm100_dbg.so physicsvae_b200/csrc/pvae_engine.cu
  PVAE_LIB=$PWD/physicsvae_b200/lib/libpvae_sm100_dbg.so PVAE_FAST_EPI=0 PVAE_TRACE_IDX=<k> python tools/trace_step.py [world|vae]
k counts the GEMM launches of the process: step s (0-based) of the world phase owns 8 s .. 8 s + 7, of the VAE phase 26 s .. 26 s + 25.
(the stamps live in the general epilogue: PVAE_FAST_EPI=0)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from physicsvae_b200 import _abi  # noqa: E402

phase = sys.argv[1] if len(sys.argv) > 1 else "world"
B = int(os.environ.get("TRACE_BATCH", "65536"))
cfg = bench.CONFIGS["default"]
tr = bench.make_trainer(cfg, B, "bf16", 0, 1, phase=phase)
tr._bind("train")
tr.model.sync_weights()
for i in range(4):
    tr.train_batch(B * (i % 4), B * (i % 4) + B)
torch.cuda.synchronize()
lib = _abi.load()
n = lib.pvae_debug_trace(None, 0, 0)
buf = (C.c_ulonglong * n)()
lib.pvae_debug_trace(buf, n, 0)
t = np.frombuffer(buf, dtype=np.uint64).reshape(160, 16, 8).astype(np.int64)
t = np.where(t == 0, np.nan, t.astype(np.float64))
L = t[0:148:2]
nun = int(np.sum(~np.isnan(L[0, :, 0])))
m = np.nanmean
print("== launch %s of a %s step: units traced per CTA pair %d" % (os.environ.get("PVAE_TRACE_IDX"), phase, nun))
if nun >= 4:
    hi = min(nun, 12)
    U = L[:, 2:hi]
    period = L[:, 3:hi, 0] - L[:, 2:hi - 1, 0]
    print("   MMA warp: period %.0f | wait tempty %.0f | wait first operands %.0f | issue k-blocks %.0f | rest %.0f" % (
        m(period), m(U[..., 1] - U[..., 0]), m(U[..., 2] - U[..., 1]), m(U[..., 3] - U[..., 2]), m(period) - m(U[..., 3] - U[..., 0])))
    print("   epilogue warp 0: wait acc %.0f | work %.0f | acc ready after MMA issue end %.0f" % (
        m(U[..., 5] - U[..., 4]), m(U[..., 6] - U[..., 5]), m(U[..., 5] - U[..., 3])))
    print("   producer: done with the unit's copies relative to MMA issue end %.0f" % m(U[..., 7] - U[..., 3]))
base = L[0, 0, 0]
for k in range(min(nun, 6)):
    print("     pair 0 unit %d: " % k + " ".join("%7.0f" % (x - base) for x in L[0, k]) + "   (mma_top tempty first_full last_issue | epi_wait acc_ready epi_done | prod_done)")
