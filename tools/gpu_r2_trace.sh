#!/bin/bash
# role timelines of selected GEMM launches (debug build; PVAE_FAST_EPI=0 for the general epilogue).  Usage: tools/gpu_r2_trace.sh world|vae idx [idx ...]
PH=$1; shift
export PVAE_LIB=$PWD/physicsvae_b200/lib/libpvae_sm100_dbg.so PVAE_FAST_EPI=${PVAE_FAST_EPI:-1}
for i in "$@"; do PVAE_TRACE_IDX=$i timeout 120 python tools/trace_step.py $PH 2>&1 | grep -E "^==|^   |pair 0 unit [0-3]"; done
