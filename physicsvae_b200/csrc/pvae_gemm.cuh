// pvae_gemm.cuh -- the one tensor-core kernel of the PhysicsVAE hot path.
//
// A persistent, warp-specialised sm_100a GEMM:  D[M,N] = epilogue( A[M,K] . B[N,K]^T )
//   * operands arrive in shared memory by TMA (cp.async.bulk.tensor, 128B swizzle),
//   * products are issued by one thread as tcgen05.mma (cta_group::1, 128 x bn x 16, bf16 -> fp32),
//   * accumulators live in TMEM (2 stages x 256 columns) and are drained by four epilogue warps
//     with tcgen05.ld while the next tile's main loop runs.
// Everything a Linear layer of the reference needs is expressed by parameters of this kernel:
//   forward   y = act(x W^T + b)            A = x  (K-major)   B = W shadow (K-major)
//   dgrad     dx = (dy W) * act'(y)         A = dy (K-major)   B = W shadow (MN-major view)
//   wgrad     dW^T = x^T dy                 A = x  (MN-major)  B = dy (MN-major), split over the batch
// (reference: rllib SlimFC = nn.Linear + activation, rllib_model_torch.py:248-253; autograd of it,
//  torch_models.py:142).  "Virtual concat" (torch.cat in rllib_model_torch.py:829,842 and
//  train_physics_vae.py:377) is done by giving operand A two K-segments with their own tensor maps.
// Precision modes: passes=1 -> plain bf16 operands; passes=3 -> bf16x3 split (hi*hi + hi*lo + lo*hi,
// operands stored as hi/lo planes) which reproduces fp32 results to ~1e-6 relative.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace pvae {

constexpr int BM = 128;                       // tile rows  (UMMA M)
constexpr int BK = 64;                        // k elements per pipeline stage (= one 128B swizzle span)
constexpr int MAX_BN = 256;                   // tile cols  (UMMA N), runtime value bn <= 256, multiple of 16
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KiB
constexpr int B_STAGE_BYTES = MAX_BN * BK * 2;    // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 8;                  // two warps per TMEM lane quarter, interleaved over 32-column chunks
constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;   // warps 0..3: TMA / MMA / TMEM-alloc / idle, warps 4..11: epilogue
constexpr int SLAB_BYTES = 32 * 128;          // per-epilogue-warp staging slab: 32 rows x 64 bf16 (TMA store / aux load box)
constexpr int OFF_STAGING = STAGES * STAGE_BYTES;
constexpr int OFF_BARS = OFF_STAGING + EPI_WARPS * SLAB_BYTES;
constexpr int OFF_BIAS = OFF_BARS + 256;
constexpr int BIAS_BYTES = EPI_WARPS * 32 * 4;      // per-epilogue-warp bias slice of the current 32 columns
constexpr int SMEM_BYTES = OFF_BIAS + BIAS_BYTES + 1024 /*alignment slack*/;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB of shared memory a CTA may use");

enum : int { EPI_STORE = 0, EPI_MSE = 1, EPI_DGRAD = 2, EPI_WGRAD = 3 };
enum : int { ACT_LINEAR = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_SIGMOID = 3, ACT_ELU = 4, ACT_SWISH = 5 };
enum : int { MAJOR_K = 0, MAJOR_MN = 1 };

struct EpiParams {
  int32_t type, act;
  int32_t m_valid, n_valid;        // rows / cols of D that exist
  const float* bias;               // [n_valid] fp32 or null
  // primary bf16 output (activation / gradient), hi plane at out, lo plane at out + out_ps
  __nv_bfloat16* out;  int64_t out_ld,  out_ps;  int32_t out_planes;  int32_t pad0;
  // secondary bf16 output (EPI_MSE: the prediction itself, e.g. the decoded action fed to the world model)
  __nv_bfloat16* out2; int64_t out2_ld, out2_ps; int32_t out2_planes; int32_t pad1;
  // fp32 output with arbitrary strides (EPI_WGRAD target, or fp32 copy of the EPI_STORE/EPI_MSE result)
  float* out_f32; int64_t f32_sm, f32_sn;
  // aux bf16 input: EPI_MSE target / EPI_DGRAD forward activation (for act')
  const __nv_bfloat16* aux; int64_t aux_ld, aux_ps; int32_t aux_planes; int32_t aux_dyn;
  int64_t aux_rows;                // rows of the aux tensor (host side: extent of its TMA descriptor)
  // addend bf16 input (EPI_DGRAD: g = acc + add)
  const __nv_bfloat16* add; int64_t add_ld, add_ps; int32_t add_planes; int32_t pad2;
  float scale;                     // EPI_MSE: d = scale * (pred - target)
  int32_t f32_atomic;              // EPI_WGRAD: 1 = red.global.add, 0 = plain store
  float* colsum;                   // [n_valid] fp32, atomically accumulated column sums of the primary output (bias grad)
  uint32_t* mask;                  // ReLU sign bits [rows][mask_ld] (1 bit per output): written by EPI_STORE, read by EPI_DGRAD
  int64_t mask_ld;                 // in 32-bit words, even
  double* loss;                    // EPI_MSE: sum of squared errors accumulated here
};

struct GemmParams {
  CUtensorMap tmA[2];              // operand A, one map per K-segment
  CUtensorMap tmB;                 // operand B
  int32_t a_major, b_major;        // MAJOR_K / MAJOR_MN
  int32_t kb[2];                   // 64-wide k-blocks per segment
  int32_t klen[2];                 // valid k elements per segment (to skip all-zero k16 slices)
  int32_t a_c0[2];                 // A: K-major -> first k element; MN-major -> first m element (inner coordinate)
  int32_t a_r0[2];                 // A: row coordinate offset (K-major: m rows, MN-major: k rows)
  int32_t a_dyn[2];                // A: add *row_cursor to the row coordinate
  int32_t b_k0[2];                 // B: k offset at the start of the segment (K-major: inner coord, MN-major: row coord)
  int32_t b_n0;                    // B: n offset
  int32_t b_dyn;                   // B: add *row_cursor to the k row coordinate (MN-major)
  int32_t passes;                  // 1 = bf16, 3 = bf16x3
  int32_t m_tiles, n_tiles, bn, splits;
  int32_t cluster;                 // 1, or 2: CTA pairs on adjacent M tiles share the B tile (each loads half, TMA multicast)
  const int32_t* row_cursor;       // device int (first row of the current mini-batch in the resident buffers) or null
  CUtensorMap tmOut;               // TMA-store epilogue: the primary bf16 output, box 64 cols x 32 rows (one warp's slab)
  CUtensorMap tmAux;               // TMA-store epilogue: the aux input (same box)
  EpiParams epi;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("pvae_gemm: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      uint64_t t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) mbar_timeout(bar, parity);   // 4 s
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
// One lane of a converged warp (warp-uniform control flow around it keeps TMA / MMA operands in uniform registers).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1, 128B swizzle).
//   K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused (1).
//   MN-major: 64-element (128 B) MN atoms x 8 k-rows; next 8 k-rows 1024 B on (SBO);
//             next MN atom one whole TMA box on = 64 k-rows * 128 B = 8192 B (LBO).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int major) {
  const uint64_t lbo = (major == MAJOR_K) ? 1ull : (8192ull >> 4);
  const uint64_t sbo = 1024ull >> 4;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}
// tcgen05 instruction descriptor, kind::f16, A/B = bf16, D = fp32, M = 128.
__device__ __forceinline__ uint32_t umma_idesc(int n, int a_major, int b_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// epilogue helpers
// ------------------------------------------------------------------------------------------------
template <int act> __device__ __forceinline__ float act_fwd(float v) {
  switch (act) {
    case ACT_RELU:    return fmaxf(v, 0.f);
    case ACT_TANH:    return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case ACT_ELU:     return v > 0.f ? v : expm1f(v);
    case ACT_SWISH:   return v / (1.f + __expf(-v));
    default:          return v;
  }
}
// derivative of the activation expressed through its OUTPUT y (what the forward pass stored)
template <int act> __device__ __forceinline__ float act_bwd_from_out(float y) {
  switch (act) {
    case ACT_RELU:    return y > 0.f ? 1.f : 0.f;
    case ACT_TANH:    return 1.f - y * y;
    case ACT_SIGMOID: return y * (1.f - y);
    case ACT_ELU:     return y > 0.f ? 1.f : y + 1.f;
    default:          return 1.f;
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo_part(float v) {  // v - bf16(v)
  return v - __bfloat162float(__float2bfloat16_rn(v));
}

// Load 32 consecutive bf16 (hi plane + optional lo plane) of one row as floats; columns >= nvalid read as 0.
__device__ __forceinline__ void load_row32(const __nv_bfloat16* base, int64_t ps, int planes, int nvalid, float (&v)[32]) {
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((ps & 7) == 0) && nvalid == 32;
  if (vec_ok) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(base) + g);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[g * 8 + 2 * j]     = __uint_as_float(w[j] << 16);
        v[g * 8 + 2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
      }
    }
    if (planes > 1) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(base + ps) + g);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[g * 8 + 2 * j]     += __uint_as_float(w[j] << 16);
          v[g * 8 + 2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
        }
      }
    }
  } else {
    const unsigned short* p0 = reinterpret_cast<const unsigned short*>(base);
    const unsigned short* p1 = reinterpret_cast<const unsigned short*>(base + ps);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float t = 0.f;
      if (i < nvalid) {
        t = __uint_as_float((uint32_t)__ldg(p0 + i) << 16);
        if (planes > 1) t += __uint_as_float((uint32_t)__ldg(p1 + i) << 16);
      }
      v[i] = t;
    }
  }
}
// Store 32 consecutive values of one row as bf16 hi (+lo) planes; 8-column groups that start at or past
// nvalid are skipped, columns past nvalid inside a written group are zero. Row base must be 16B aligned.
__device__ __forceinline__ void store_row32(__nv_bfloat16* base, int64_t ps, int planes, int nvalid, float (&v)[32]) {
  if (nvalid < 32) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (i < nvalid) ? v[i] : 0.f;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g * 8 < nvalid) {
      uint4 q;
      q.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]); q.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
      q.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]); q.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
      *(reinterpret_cast<uint4*>(base) + g) = q;
      if (planes > 1) {
        uint4 l;
        l.x = pack_bf16x2(bf16_lo_part(v[g * 8 + 0]), bf16_lo_part(v[g * 8 + 1]));
        l.y = pack_bf16x2(bf16_lo_part(v[g * 8 + 2]), bf16_lo_part(v[g * 8 + 3]));
        l.z = pack_bf16x2(bf16_lo_part(v[g * 8 + 4]), bf16_lo_part(v[g * 8 + 5]));
        l.w = pack_bf16x2(bf16_lo_part(v[g * 8 + 6]), bf16_lo_part(v[g * 8 + 7]));
        *(reinterpret_cast<uint4*>(base + ps) + g) = l;
      }
    }
  }
}
__device__ __forceinline__ void store_f32_row32(float* dst, int64_t sn, int nvalid, const float (&v)[32]) {
  if (sn == 1 && nvalid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
    for (int g = 0; g < 8; ++g) reinterpret_cast<float4*>(dst)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    return;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < nvalid) dst[(int64_t)i * sn] = v[i];
}
// 64 values of one row (a TMA-epilogue sub-tile) to an fp32 output with column stride sn
__device__ __forceinline__ void store_f32_row64(float* dst, int64_t sn, int nvalid, const float (&v)[64]) {
  const float (&v0)[32] = *reinterpret_cast<const float (*)[32]>(&v[0]);
  const float (&v1)[32] = *reinterpret_cast<const float (*)[32]>(&v[32]);
  store_f32_row32(dst, sn, nvalid > 32 ? 32 : nvalid, v0);
  if (nvalid > 32) store_f32_row32(dst + 32 * sn, sn, nvalid - 32, v1);
}
// Transposing butterfly: on return lane j holds sum over the 32 lanes of v[j].
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------------
// epilogue of one 32-column chunk of one output row (compile-time epilogue type and activation keep
// the instruction footprint of every instantiation small enough to stay in the instruction cache)
// ------------------------------------------------------------------------------------------------
template <int EPI, int ACT>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& e, float (&v)[32], const float* bias_s, int row, bool row_ok,
                                               int col0, int nv, int row0, int lane, double& loss_local) {
  if (EPI == EPI_WGRAD) {
    if (row_ok) {
      float* dst = e.out_f32 + (int64_t)row * e.f32_sm + (int64_t)col0 * e.f32_sn;
      const int64_t sn = e.f32_sn;
      if (e.f32_atomic) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nv) atomicAdd(dst + (int64_t)i * sn, v[i]);
      } else {
        store_f32_row32(dst, sn, nv, v);
      }
    }
    return;
  }
  if (EPI != EPI_DGRAD) {
    if (e.bias) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 b = *reinterpret_cast<const float4*>(bias_s + g * 4);
        v[g * 4 + 0] += b.x; v[g * 4 + 1] += b.y; v[g * 4 + 2] += b.z; v[g * 4 + 3] += b.w;
      }
    }
  }
  if (EPI == EPI_STORE) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = act_fwd<ACT>(v[i]);
    if (row_ok) {
      if (e.out_f32) store_f32_row32(e.out_f32 + (int64_t)row * e.f32_sm + (int64_t)col0 * e.f32_sn, e.f32_sn, nv, v);
      if (e.out) store_row32(e.out + (int64_t)row * e.out_ld + col0, e.out_ps, e.out_planes, nv, v);
    }
  } else if (EPI == EPI_MSE) {
    float t[32];
    if (row_ok) {
      const int64_t arow = (int64_t)row + (e.aux_dyn ? row0 : 0);
      load_row32(e.aux + arow * e.aux_ld + col0, e.aux_ps, e.aux_planes, nv, t);
      if (e.out_f32) store_f32_row32(e.out_f32 + (int64_t)row * e.f32_sm + (int64_t)col0 * e.f32_sn, e.f32_sn, nv, v);
    }
    float sq = 0.f;
    const float scale = e.scale;
    float d[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float di = (row_ok && i < nv) ? (v[i] - t[i]) : 0.f;
      sq += di * di;
      d[i] = scale * di;
    }
    loss_local += (double)sq;
    if (row_ok) {
      if (e.out2) store_row32(e.out2 + (int64_t)row * e.out2_ld + col0, e.out2_ps, e.out2_planes, nv, v);
      if (e.out) store_row32(e.out + (int64_t)row * e.out_ld + col0, e.out_ps, e.out_planes, nv, d);
    }
    if (e.colsum) {
      const float s = warp_colsum32(d, lane);
      if (lane < nv) atomicAdd(e.colsum + col0 + lane, s);
    }
  } else {  // EPI_DGRAD
    if (row_ok) {
      if (e.add) {
        float a[32];
        load_row32(e.add + (int64_t)row * e.add_ld + col0, e.add_ps, e.add_planes, nv, a);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += a[i];
      }
      if (ACT != ACT_LINEAR) {
        float y[32];
        load_row32(e.aux + (int64_t)row * e.aux_ld + col0, e.aux_ps, e.aux_planes, nv, y);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= act_bwd_from_out<ACT>(y[i]);
      }
      if (e.out_f32) store_f32_row32(e.out_f32 + (int64_t)row * e.f32_sm + (int64_t)col0 * e.f32_sn, e.f32_sn, nv, v);
      if (e.out) store_row32(e.out + (int64_t)row * e.out_ld + col0, e.out_ps, e.out_planes, nv, v);
    }
    if (e.colsum) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (row_ok && i < nv) ? v[i] : 0.f;
      const float s = warp_colsum32(v, lane);
      if (lane < nv) atomicAdd(e.colsum + col0 + lane, s);
    }
  }
}

// position inside the (pass, segment, k-block) iteration space of one output tile
struct KIter {
  int pass, seg, r;
  __device__ __forceinline__ void init(int it, int kb0, int kb_total) {
    pass = it / kb_total;
    const int rem = it - pass * kb_total;
    seg = rem >= kb0 ? 1 : 0;
    r = seg ? rem - kb0 : rem;
  }
  __device__ __forceinline__ void next(int kb0, int kb1) {
    ++r;
    if (r == (seg ? kb1 : kb0)) {
      r = 0;
      if (seg == 0 && kb1 > 0) seg = 1;
      else { seg = 0; ++pass; }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// TMAEPI: the primary bf16 output (one precision plane) leaves through shared memory and cp.async.bulk.tensor stores, the
// aux operand (forward activation for act', MSE target) arrives the same way; the epilogue warps then touch only TMEM,
// shared memory and registers.  Without it (bf16x3 planes, fp32-only outputs) rows are read / written directly.
template <int EPI, int ACT, bool TMAEPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) pvae_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 128B swizzle atoms need 1024 B alignment
  const uint32_t bar_base = smem_base + OFF_BARS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
  auto auxfull_bar = [&](int w) { return bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES + w); };   // one per epilogue warp
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES + EPI_WARPS);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));    // generic pointer to the aligned base
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));
  float* bias_all = reinterpret_cast<float*>(smem_gen + OFF_BIAS);
  // aux operand through TMA: the MSE target, or the forward activation for act' (ReLU uses the sign-bit mask instead)
  constexpr bool HAS_AUX = TMAEPI && (EPI == EPI_MSE || (EPI == EPI_DGRAD && ACT != ACT_LINEAR && ACT != ACT_RELU));
  constexpr bool USE_MASK = TMAEPI && EPI == EPI_DGRAD && ACT == ACT_RELU;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    if (p.kb[1] > 0) tma_prefetch_desc(&p.tmA[1]);
    tma_prefetch_desc(&p.tmB);
    if (TMAEPI) tma_prefetch_desc(&p.tmOut);
    if (HAS_AUX) tma_prefetch_desc(&p.tmAux);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), p.cluster); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS); }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(auxfull_bar(w), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();        // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int csize = p.cluster;
  const int crank = csize > 1 ? (int)cluster_ctarank() : 0;
  const int unit0 = blockIdx.x / csize, unit_stride = gridDim.x / csize;

  const int row0 = p.row_cursor ? *p.row_cursor : 0;
  const int kb_total = p.kb[0] + p.kb[1];
  const int iters_total = p.passes * kb_total;
  // units of one CTA (pair): tile-minor / split-major, so that the CTAs running at the same time share K ranges through L2
  const int tile_units = ((p.m_tiles + csize - 1) / csize) * p.n_tiles;
  const int total_units = tile_units * p.splits;
  const int bn = p.bn;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // The whole warp walks the loop (warp-uniform control flow); one elected lane issues the copies.
    const int b_boxes = (bn + 63) >> 6;
    const uint32_t stage_tx = A_STAGE_BYTES + (p.b_major == MAJOR_K ? (uint32_t)bn * (BK * 2) : (uint32_t)b_boxes * 8192u);
    const int a_major = p.a_major, b_major = p.b_major;
    const int hrows = bn >> 1;
    int stage = 0; uint32_t phase = 0;
    for (int u = unit0; u < total_units; u += unit_stride) {
      const int split = u / tile_units;
      const int tile = u - split * tile_units;
      const int n_tile = tile % p.n_tiles;
      const int m_tile = (tile / p.n_tiles) * csize + crank;
      const int it_begin = (int)(((int64_t)iters_total * split) / p.splits);
      const int it_end = (int)(((int64_t)iters_total * (split + 1)) / p.splits);
      KIter ki; ki.init(it_begin, p.kb[0], kb_total);
      for (int it = it_begin; it < it_end; ++it, ki.next(p.kb[0], p.kb[1])) {
        const int pass = ki.pass, seg = ki.seg, r = ki.r;
        const int a_plane = (pass == 2) ? 1 : 0;
        const int b_plane = (pass == 1) ? 1 : 0;
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint32_t sb = sa + A_STAGE_BYTES;
        const uint32_t fb = full_bar(stage);
        const int a_row = p.a_r0[seg] + (p.a_dyn[seg] ? row0 : 0);
        const int b_k = p.b_k0[seg] + r * BK;
        const int b_n = p.b_n0 + n_tile * bn;
        const CUtensorMap* tmA = &p.tmA[seg];
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(fb, stage_tx);
          if (a_major == MAJOR_K) {
            tma_load_3d(sa, tmA, fb, p.a_c0[seg] + r * BK, a_row + m_tile * BM, a_plane);
          } else {
            tma_load_3d(sa,         tmA, fb, p.a_c0[seg] + m_tile * BM,      a_row + r * BK, a_plane);
            tma_load_3d(sa + 8192u, tmA, fb, p.a_c0[seg] + m_tile * BM + 64, a_row + r * BK, a_plane);
          }
          if (csize == 1) {
            if (b_major == MAJOR_K) {
              tma_load_3d(sb, &p.tmB, fb, b_k, b_n, b_plane);
            } else {
              const int b_row = b_k + (p.b_dyn ? row0 : 0);
              for (int j = 0; j < b_boxes; ++j) tma_load_3d(sb + 8192u * j, &p.tmB, fb, b_n + 64 * j, b_row, b_plane);
            }
          } else {
            // CTA pair: this CTA fetches its half of the B tile and multicasts it into both CTAs' shared memory
            if (b_major == MAJOR_K) {           // K-major: rows [crank * bn/2, +bn/2) of the tile (box rows = bn/2)
              tma_load_3d_mc(sb + (uint32_t)(crank * hrows) * 128u, &p.tmB, fb, b_k, b_n + crank * hrows, b_plane, (uint16_t)3);
            } else {
              const int b_row = b_k + (p.b_dyn ? row0 : 0);
              for (int j = crank; j < b_boxes; j += 2)
                tma_load_3d_mc(sb + 8192u * j, &p.tmB, fb, b_n + 64 * j, b_row, b_plane, (uint16_t)3);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t idesc = umma_idesc(bn, p.a_major, p.b_major);
    const uint32_t a_kstep = (p.a_major == MAJOR_K) ? 2u : 128u;   // 16 k elements, in 16 B units: 32 B or 16 rows * 128 B
    const uint32_t b_kstep = (p.b_major == MAJOR_K) ? 2u : 128u;
    const uint64_t adesc0 = umma_desc(smem_base, p.a_major);
    const uint64_t bdesc0 = umma_desc(smem_base + A_STAGE_BYTES, p.b_major);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = unit0; u < total_units; u += unit_stride) {
      const int split = u / tile_units;
      const int it_begin = (int)(((int64_t)iters_total * split) / p.splits);
      const int it_end = (int)(((int64_t)iters_total * (split + 1)) / p.splits);
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)acc * MAX_BN;
      KIter ki; ki.init(it_begin, p.kb[0], kb_total);
      uint32_t accum = 0u;
      for (int it = it_begin; it < it_end; ++it, ki.next(p.kb[0], p.kb[1])) {
        int k16 = (p.klen[ki.seg] - ki.r * BK + 15) >> 4;
        k16 = k16 > 4 ? 4 : (k16 < 1 ? 1 : k16);
        const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)(stage * STAGE_BYTES) >> 4);
        const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)(stage * STAGE_BYTES) >> 4);
        const uint32_t eb = empty_bar(stage);
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          umma_bf16(tmem_d, adesc, bdesc, idesc, accum);
          if (k16 > 1) umma_bf16(tmem_d, adesc + (uint64_t)a_kstep, bdesc + (uint64_t)b_kstep, idesc, 1u);
          if (k16 > 2) umma_bf16(tmem_d, adesc + (uint64_t)(2 * a_kstep), bdesc + (uint64_t)(2 * b_kstep), idesc, 1u);
          if (k16 > 3) umma_bf16(tmem_d, adesc + (uint64_t)(3 * a_kstep), bdesc + (uint64_t)(3 * b_kstep), idesc, 1u);
          if (csize == 1) umma_commit(eb);                 // frees the smem slot once these MMAs have read it
          else umma_commit_mc(eb, (uint16_t)3);            // ... in both CTAs of the pair (both producers write both)
        }
        __syncwarp();
        accum = 1u;
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(tfull_bar(acc));        // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const EpiParams& e = p.epi;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read (warp id % 4)
    const int half = (warp - 4) >> 2;             // which 32-column half of a 64-column sub-tile / which chunk parity
    float* bias_s = bias_all + (warp - 4) * 32;
    const int m_valid = e.m_valid, n_valid = e.n_valid;
    const float* bias = (EPI == EPI_STORE || EPI == EPI_MSE) ? e.bias : nullptr;
    int acc = 0; uint32_t acc_phase = 0;
    double loss_local = 0.0;
    uint32_t aux_phase = 0;
    const uint32_t slab = smem_base + OFF_STAGING + (warp - 4) * SLAB_BYTES;      // this warp's private staging slab
    uint8_t* slab_gen = smem_gen + OFF_STAGING + (warp - 4) * SLAB_BYTES;
    for (int u = unit0; u < total_units; u += unit_stride) {
      const int tile = u % tile_units;
      const int n_tile = tile % p.n_tiles;
      const int m_tile = (tile / p.n_tiles) * csize + crank;
      const int row_in_tile = q * 32 + lane;
      const int row = m_tile * BM + row_in_tile;
      const bool row_ok = row < m_valid;
      // operands of this warp's (at most two) sub-tiles that do not depend on the accumulator: fetched while the main loop runs
      float pre_bias[4] = {0.f, 0.f, 0.f, 0.f};
      uint2 pre_mask[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
      if (TMAEPI) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int cs = n_tile * bn + (half + 2 * j) * 64;
          if ((half + 2 * j) * 64 < bn && cs < n_valid) {
            if (bias) {
              pre_bias[2 * j] = (cs + lane < n_valid) ? __ldg(bias + cs + lane) : 0.f;
              pre_bias[2 * j + 1] = (cs + lane + 32 < n_valid) ? __ldg(bias + cs + lane + 32) : 0.f;
            }
            if (USE_MASK && row_ok) pre_mask[j] = __ldg(reinterpret_cast<const uint2*>(e.mask + (int64_t)row * e.mask_ld + (cs >> 5)));
          }
        }
      }
      if (HAS_AUX) {
        // fetch the aux slab of this warp's first sub-tile while the main loop of the tile is still running
        const int col0s = n_tile * bn + half * 64;
        if (col0s < n_valid && half * 64 < bn) {
          tma_store_wait_read<0>();               // the slab's previous store has been read out (groups are per thread)
          __syncwarp();
          if (elect_one()) {
            mbar_expect_tx(auxfull_bar(warp - 4), SLAB_BYTES);
            tma_load_3d(slab, &p.tmAux, auxfull_bar(warp - 4), col0s, (e.aux_dyn ? row0 : 0) + m_tile * BM + q * 32, 0);
          }
          __syncwarp();
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (!TMAEPI) {
        const int nchunks = (bn + 31) >> 5;
        for (int c = half; c < nchunks; c += 2) {
          const int col0 = n_tile * bn + c * 32;
          if (col0 >= n_valid) break;             // warp-uniform
          int nv = n_valid - col0; nv = nv > 32 ? 32 : nv;
          const int tile_nv = bn - c * 32;        // columns of this chunk that belong to this tile
          if (tile_nv < nv) nv = tile_nv;
          if (bias) {
            __syncwarp();
            bias_s[lane] = (lane < nv) ? __ldg(bias + col0 + lane) : 0.f;
            __syncwarp();
          }
          uint32_t raw[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * MAX_BN + c * 32), raw);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          epilogue_chunk<EPI, ACT>(e, v, bias_s, row, row_ok, col0, nv, row0, lane, loss_local);
        }
      } else {
        // Each warp owns the 32 rows of its TMEM lane quarter and every other 64-column sub-tile; it stages its 32 x 64 bf16
        // slab in private shared memory and stores it with one TMA instruction -- no cross-warp synchronisation.
        const int nsub = (bn + 63) >> 6;
        const int sw = lane & 7;
        uint8_t* srow = slab_gen + lane * 128;
        for (int s = half; s < nsub; s += 2) {
          const int col0s = n_tile * bn + s * 64;
          if (col0s >= n_valid) break;            // warp-uniform
          int nv = n_valid - col0s; nv = nv > 64 ? 64 : nv;
          const int tile_nv = bn - s * 64;        // columns of this sub-tile that belong to this tile
          if (tile_nv < nv) nv = tile_nv;
          const int sj = (s - half) >> 1;         // 0 or 1: which of this warp's sub-tiles (bn <= 256)
          const uint2 mbits = sj ? pre_mask[1] : pre_mask[0];
          const float bias_lo = sj ? pre_bias[2] : pre_bias[0];   // lane -> bias of columns lane and lane + 32 of the sub-tile
          const float bias_hi = sj ? pre_bias[3] : pre_bias[1];
          if (HAS_AUX && s != half) {             // later sub-tiles of the tile: the aux load is exposed
            tma_store_wait_read<0>();
            __syncwarp();
            if (elect_one()) {
              mbar_expect_tx(auxfull_bar(warp - 4), SLAB_BYTES);
              tma_load_3d(slab, &p.tmAux, auxfull_bar(warp - 4), col0s, (e.aux_dyn ? row0 : 0) + m_tile * BM + q * 32, 0);
            }
            __syncwarp();
          }
          uint32_t raw[64];
          {
            uint32_t (&r0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[0]);
            uint32_t (&r1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[32]);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * MAX_BN + s * 64);
            tmem_ld32(taddr, r0);
            tmem_ld32(taddr + 32, r1);
            tmem_ld_wait();
          }
          float v[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(raw[i]);
          if (bias) {
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              __syncwarp();
              bias_s[lane] = hb ? bias_hi : bias_lo;
              __syncwarp();
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b = *reinterpret_cast<const float4*>(bias_s + g * 4);
                v[hb * 32 + g * 4 + 0] += b.x; v[hb * 32 + g * 4 + 1] += b.y; v[hb * 32 + g * 4 + 2] += b.z; v[hb * 32 + g * 4 + 3] += b.w;
              }
            }
          }
          if (EPI == EPI_STORE) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = act_fwd<ACT>(v[i]);
            if (nv < 64) {                        // ragged last sub-tile: keep the padding columns of the mask / slab zero
#pragma unroll
              for (int i = 0; i < 64; ++i) v[i] = (i < nv) ? v[i] : 0.f;
            }
            if (row_ok) {
              if (e.out_f32) store_f32_row64(e.out_f32 + (int64_t)row * e.f32_sm + (int64_t)col0s * e.f32_sn, e.f32_sn, nv, v);
              if (ACT == ACT_RELU && e.mask) {
                uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
                for (int i = 0; i < 32; ++i) { m0 |= (v[i] > 0.f ? 1u : 0u) << i; m1 |= (v[32 + i] > 0.f ? 1u : 0u) << i; }
                *reinterpret_cast<uint2*>(e.mask + (int64_t)row * e.mask_ld + (col0s >> 5)) = make_uint2(m0, m1);
              }
            }
          } else {
            float y[64];
            if (HAS_AUX) {
              mbar_wait(auxfull_bar(warp - 4), aux_phase);
              aux_phase ^= 1u;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 qv = *reinterpret_cast<const uint4*>(srow + ((j ^ sw) << 4));
                const uint32_t w[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  y[j * 8 + 2 * t]     = __uint_as_float(w[t] << 16);
                  y[j * 8 + 2 * t + 1] = __uint_as_float(w[t] & 0xFFFF0000u);
                }
              }
              __syncwarp();                       // every lane has read the aux slab before it is overwritten below
            }
            if (EPI == EPI_MSE) {
              if (row_ok) {
                if (e.out_f32) store_f32_row64(e.out_f32 + (int64_t)row * e.f32_sm + (int64_t)col0s * e.f32_sn, e.f32_sn, nv, v);
                if (e.out2) {
                  float (&v0)[32] = *reinterpret_cast<float (*)[32]>(&v[0]);
                  float (&v1)[32] = *reinterpret_cast<float (*)[32]>(&v[32]);
                  store_row32(e.out2 + (int64_t)row * e.out2_ld + col0s, e.out2_ps, e.out2_planes, nv > 32 ? 32 : nv, v0);
                  if (nv > 32) store_row32(e.out2 + (int64_t)row * e.out2_ld + col0s + 32, e.out2_ps, e.out2_planes, nv - 32, v1);
                }
              }
              float sq = 0.f;
              const float scale = e.scale;
#pragma unroll
              for (int i = 0; i < 64; ++i) {
                const float di = (row_ok && i < nv) ? (v[i] - y[i]) : 0.f;
                sq += di * di;
                v[i] = scale * di;
              }
              loss_local += (double)sq;
            } else {  // EPI_DGRAD
              if (e.add && row_ok) {
                float a[32];
                load_row32(e.add + (int64_t)row * e.add_ld + col0s, e.add_ps, e.add_planes, nv > 32 ? 32 : nv, a);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += a[i];
                if (nv > 32) {
                  load_row32(e.add + (int64_t)row * e.add_ld + col0s + 32, e.add_ps, e.add_planes, nv - 32, a);
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[32 + i] += a[i];
                }
              }
#pragma unroll
              for (int i = 0; i < 64; ++i) {
                float g = v[i];
                if (USE_MASK) g = ((i < 32 ? mbits.x >> i : mbits.y >> (i - 32)) & 1u) ? g : 0.f;   // mask is 0 for rows >= m_valid
                else if (ACT != ACT_LINEAR) g *= act_bwd_from_out<ACT>(y[i]);
                v[i] = g;
              }
              if (nv < 64 || (!USE_MASK && !row_ok)) {   // keep what the bias-gradient column sums must not see at zero
#pragma unroll
                for (int i = 0; i < 64; ++i) v[i] = (row_ok && i < nv) ? v[i] : 0.f;
              }
              if (row_ok && e.out_f32) store_f32_row64(e.out_f32 + (int64_t)row * e.f32_sm + (int64_t)col0s * e.f32_sn, e.f32_sn, nv, v);
            }
          }
          if (!HAS_AUX) {                         // (with aux the slab was already claimed before the aux load)
            tma_store_wait_read<0>();             // bulk groups are per thread: only the electing lane ever has pending ones
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 qv;
            qv.x = pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]); qv.y = pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]);
            qv.z = pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]); qv.w = pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]);
            *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = qv;
          }
          fence_proxy_async();                    // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (m_tile * BM + q * 32 < m_valid) {   // warp-uniform
            if (elect_one()) {
              tma_store_3d(&p.tmOut, slab, col0s, m_tile * BM + q * 32, 0);
              tma_store_commit();
            }
            __syncwarp();
          }
          if (EPI != EPI_STORE && e.colsum) {
            // bias gradient: column sums of the staged (bf16-rounded) slab; lane -> columns lane, lane + 32
#pragma unroll
            for (int hcol = 0; hcol < 2; ++hcol) {
              const int col = lane + 32 * hcol;
              if (col < nv) {
                float sum = 0.f;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                  const unsigned short hv = *reinterpret_cast<const unsigned short*>(slab_gen + r * 128 + ((((col >> 3) ^ (r & 7)) << 4) | ((col & 7) << 1)));
                  sum += __uint_as_float((uint32_t)hv << 16);
                }
                atomicAdd(e.colsum + col0s + col, sum);
              }
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
    if (TMAEPI) tma_store_wait_read<0>();
    if (EPI == EPI_MSE && e.loss) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) loss_local += __shfl_xor_sync(0xffffffffu, loss_local, off);
      if (lane == 0 && loss_local != 0.0) atomicAdd(e.loss, loss_local);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();            // no CTA leaves while its peer may still multicast into it / signal its barriers
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

typedef void (*GemmKernelFn)(const GemmParams);
// host-side dispatch table: [epilogue type][activation][tma epilogue]
template <int EPI, bool T> struct KernelRow {
  static GemmKernelFn get(int act) {
    switch (act) {
      case ACT_RELU:    return pvae_gemm_kernel<EPI, ACT_RELU, T>;
      case ACT_TANH:    return pvae_gemm_kernel<EPI, ACT_TANH, T>;
      case ACT_SIGMOID: return pvae_gemm_kernel<EPI, ACT_SIGMOID, T>;
      case ACT_ELU:     return pvae_gemm_kernel<EPI, ACT_ELU, T>;
      case ACT_SWISH:   return pvae_gemm_kernel<EPI, ACT_SWISH, T>;
      default:          return pvae_gemm_kernel<EPI, ACT_LINEAR, T>;
    }
  }
};
static inline GemmKernelFn select_kernel(int epi, int act, bool tma) {
  switch (epi) {
    case EPI_STORE: return tma ? KernelRow<EPI_STORE, true>::get(act) : KernelRow<EPI_STORE, false>::get(act);
    case EPI_DGRAD: return tma ? KernelRow<EPI_DGRAD, true>::get(act) : KernelRow<EPI_DGRAD, false>::get(act);
    case EPI_MSE:   return tma ? pvae_gemm_kernel<EPI_MSE, ACT_LINEAR, true> : pvae_gemm_kernel<EPI_MSE, ACT_LINEAR, false>;
    default:        return pvae_gemm_kernel<EPI_WGRAD, ACT_LINEAR, false>;
  }
}

}  // namespace pvae
