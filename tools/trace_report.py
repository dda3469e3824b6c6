"""Summarise gpurun_out/trace_<idx>.npy files (tools/gpu_trace.sh): per-unit role timeline of a GEMM launch."""
import sys, numpy as np
names = {24: "fwd L0", 25: "fwd L1", 26: "fwd L2 MSE", 27: "wgrad L2", 28: "dgrad L1", 29: "wgrad L1", 30: "dgrad L0", 31: "wgrad L0"}
import glob, os, re
for f in glob.glob("gpurun_out/trace_*.npy"):
    i = int(re.search(r"trace_(\d+)", f).group(1))
    names.setdefault(i, "launch %d" % i)
for idx in sorted(names):
    try:
        t = np.load("gpurun_out/trace_%d.npy" % idx).reshape(160, 16, 8).astype(np.int64)
    except Exception as e:
        continue
    t = np.where(t == 0, np.nan, t.astype(np.float64))
    L = t[0:148:2]                       # leader CTAs (even rank)
    nun = int(np.sum(~np.isnan(L[0, :, 0])))
    print("== %d %s: units traced per CTA %d" % (idx, names[idx], nun))
    def m(a): return np.nanmean(a)
    U = L[:, 2:min(nun, 12)]
    period = L[:, 3:min(nun, 12), 0] - L[:, 2:min(nun, 12) - 1, 0]
    print("   MMA warp: period %.0f | wait tempty %.0f | wait first operands %.0f | issue k-blocks %.0f | rest %.0f" % (
        m(period), m(U[..., 1] - U[..., 0]), m(U[..., 2] - U[..., 1]), m(U[..., 3] - U[..., 2]), m(period) - m(U[..., 3] - U[..., 0])))
    print("   epilogue warp 0: wait acc %.0f | work %.0f | acc ready after MMA issue end %.0f | epi done -> next tempty seen %.0f" % (
        m(U[..., 5] - U[..., 4]), m(U[..., 6] - U[..., 5]), m(U[..., 5] - U[..., 3]), m(L[:, 4:min(nun, 12), 1] - L[:, 2:min(nun, 12) - 2, 6])))
    print("   producer: done with unit's copies relative to MMA issue end %.0f" % m(U[..., 7] - U[..., 3]))
    c = 0
    base = L[c, 0, 0]
    for k in range(min(nun, 6)):
        print("     cta0 unit %d: " % k + " ".join("%7.0f" % (x - base) for x in L[c, k]))
