"""Per-unit role timeline of selected GEMM launches of one world step (PVAE_DBG=32).  Runs on the GPU box.
Usage: PVAE_DBG=32 python tools/trace_units.py"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physicsvae_b200 import _abi
lib = _abi.load()
dev = torch.device("cuda:0")
N = 160 * 16 * 8
buf = (C.c_ulonglong * N)()

def run(name, M, N_, K, a_major, b_major, splits=1):
    A = torch.randn((K, M) if a_major else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N_) if b_major else (N_, K), device=dev).bfloat16()
    D = torch.zeros(M, N_, device=dev)
    for _ in range(3):
        _abi.check(lib.pvae_gemm_bf16(A.data_ptr(), a_major, B.data_ptr(), b_major, M, N_, K, 1, splits, D.data_ptr(), None))
    torch.cuda.synchronize()
    lib.pvae_debug_trace(None, 0, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _abi.check(lib.pvae_gemm_bf16(A.data_ptr(), a_major, B.data_ptr(), b_major, M, N_, K, 1, splits, D.data_ptr(), None))
    e1.record(); torch.cuda.synchronize()
    lib.pvae_debug_trace(buf, N, 0)
    t = np.frombuffer(buf, dtype=np.uint64).reshape(160, 16, 8).astype(np.int64)
    print("== %s M=%d N=%d K=%d majors %d%d: %.1f us" % (name, M, N_, K, a_major, b_major, e0.elapsed_time(e1) * 1e3))
    for cta in (0, 2, 74):
        base = t[cta, 0, 0]
        if base == 0: continue
        print(" cta %d (relative clocks; slots: mma_top tempty first_full last_issue | epi_wait acc_ready epi_done | prod_done)" % cta)
        for k in range(8):
            r = t[cta, k]
            if r[0] == 0: break
            print("   unit %2d: " % k + " ".join("%7d" % (x - base if x else -1) for x in r))
    # aggregate over leader CTAs: mean durations per unit (units 2..10)
    L = t[0:148:2, 2:11]
    ok = (L[..., 0] > 0) & (L[..., 6] > 0)
    def mean(a): return float(a[ok].mean()) if ok.any() else float('nan')
    print("  mean clocks/unit: period %.0f | tempty wait %.0f | first operands wait %.0f | issue span %.0f | epi: wait acc %.0f, work %.0f" % (
        mean(L[:, 1:, 0] - L[:, :-1, 0]) if L.shape[1] > 1 and ((L[:, 1:, 0] > 0) & (L[:, :-1, 0] > 0)).all() else float(np.nanmean(np.where((L[:, 1:, 0] > 0) & (L[:, :-1, 0] > 0), L[:, 1:, 0] - L[:, :-1, 0], np.nan))),
        mean(L[..., 1] - L[..., 0]), mean(L[..., 2] - L[..., 1]), mean(L[..., 3] - L[..., 2]), mean(L[..., 5] - L[..., 4]), mean(L[..., 6] - L[..., 5])))

B = 65536
run("fwd-L0-like (K=256)", B, 1024, 256, 0, 0)
run("fwd-L1-like (K=1024)", B, 1024, 1024, 0, 0)
run("dgrad-L1-like (K=197->256, B MN)", B, 1024, 256, 0, 1)
run("wgrad-L1-like", 1024, 1024, B, 1, 1, splits=9)
