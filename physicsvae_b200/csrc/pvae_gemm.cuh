// pvae_gemm.cuh -- the one tensor-core kernel of the PhysicsVAE hot path.
//
// A persistent, warp-specialised sm_100a GEMM:  D[M,N] = epilogue( A[M,K] . B[N,K]^T )
//   * operands arrive in shared memory by TMA (cp.async.bulk.tensor, 128B swizzle),
//   * products are issued by one thread as tcgen05.mma: CG = 2 (the normal case) pairs two CTAs of a cluster on a
//     256 x bn x 16 instruction (cta_group::2) -- each CTA stages its own 128 rows of A and HALF of the B tile, the
//     leader CTA issues for both; CG = 1 (single M tile) is the plain 128 x bn x 16 form,
//   * accumulators live in TMEM (2 stages x 256 columns) and are drained by eight epilogue warps
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
#include <type_traits>

namespace pvae {

constexpr int BM = 128;                       // tile rows  (UMMA M)
constexpr int BK = 64;                        // k elements per pipeline stage (= one 128B swizzle span)
constexpr int MAX_BN = 256;                   // tile cols  (UMMA N), runtime value bn <= 256, multiple of 16
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KiB
// Pipeline geometry per MMA mode: CG = 2 stages half a B tile per CTA (32 KiB slots, 6 deep), CG = 1 a whole one (48 KiB, 4 deep).
template <int CG> struct Geo {
  static constexpr int STAGES = CG == 2 ? 6 : 4;
  static constexpr int B_STAGE_BYTES = MAX_BN * BK * 2 / CG;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
};
constexpr int MAX_STAGES = 8;
constexpr int PIPE_BYTES = 196608;                // 6 x 32 KiB = 4 x 48 KiB
static_assert(Geo<1>::STAGES * Geo<1>::STAGE_BYTES == PIPE_BYTES && Geo<2>::STAGES * Geo<2>::STAGE_BYTES == PIPE_BYTES, "pipeline size");
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 16;                 // four warps per TMEM lane quarter, interleaved over 32-column chunks
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;    // warp 0: TMA producer, warp 1: MMA issuer + TMEM allocation, warps 2..17: epilogue (576 threads -> 112 registers each)
constexpr int SLAB_BYTES = 32 * 64;           // per-epilogue-warp staging slab: 32 rows x 32 bf16 (TMA store / aux load box, 64B swizzle)
constexpr int OFF_STAGING = PIPE_BYTES;
constexpr int OFF_BARS = OFF_STAGING + EPI_WARPS * SLAB_BYTES;
constexpr int OFF_BIAS = OFF_BARS + 512;
constexpr int BIAS_BYTES = EPI_WARPS * 32 * 4;      // per-epilogue-warp bias slice of the current 32 columns
constexpr int SMEM_BYTES = OFF_BIAS + BIAS_BYTES + 512 /*alignment slack: the dynamic window starts 1 KiB aligned in practice; checked*/;
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
  uint32_t* mask;                  // ReLU sign bits [32-column word][mask_ld rows] (1 bit per output): written by EPI_STORE, read by EPI_DGRAD
  int64_t mask_ld;                 // rows per 32-column word plane (the workspace batch capacity)
  double* loss;                    // EPI_MSE: sum of squared errors accumulated here
  const float* act_param;          // swish: beta of x * sigmoid(beta x) (device scalar, rllib's Swish parameter); null = 1
  float* act_grad;                 // EPI_DGRAD + swish: d loss / d beta accumulated here (null: not wanted)
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
  // M segments (MN-major A only, wgrad of a layer whose input is a virtual concat): M tiles [0, m_seg_tiles) come from
  // tmA[0], the rest from tmA[1]; a segment-1 tile's rows land in the output at row m_seg_out0 + local row.  0 = off.
  int32_t m_seg_tiles, m_seg_rows[2], m_seg_out0;
  int32_t m_gap0, m_gap;           // rows [m_gap0, m_gap0 + m_gap) of D are an alignment gap of the A operand: not stored, later rows move up
  int32_t passes;                  // 1 = bf16, 3 = bf16x3
  int32_t m_tiles, n_tiles, bn, splits;
  int32_t dbg;                     // PVAE_DBG bit mask: skip parts of the TMA epilogue (timing experiments only, results are wrong)
  int32_t reverse;                 // walk the batch dimension from its far end (see UnitWalk)
  int32_t cs_mma;                  // bias-gradient column sums: 1 = mma.sync (ones . slab), 0 = lanes add columns, -1 = by K depth
  int32_t pf_dist;                 // (debug build only) L2 prefetch of the streamed operand(s): units ahead of the one being loaded (0 = off)
  int32_t pf_b;                    // prefetch B as well (weight gradients stream both operands; forward / dgrad weights live in L2)
  int32_t cg;                      // 1, or 2: CTA pairs on adjacent M tiles run one 256-row tcgen05.mma.cta_group::2 (kernel template CG)
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
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.  The spinning part lives out of
// line so that the pipeline loops (one iteration per 64-wide k-block) stay a few dozen instructions long.
// SITE only separates the call sites (producer / MMA / epilogue) in profiles: stall samples land in the instance of the waiter.
enum : int { W_EMPTY = 0, W_FULL = 1, W_TEMPTY = 2, W_TFULL = 3, W_AUX = 4 };
template <int SITE> __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      uint64_t t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) {   // 4 s
        printf("pvae_gemm: mbarrier wait %d timed out (block %d thread %d bar 0x%x parity %u)\n", SITE, (int)blockIdx.x, (int)threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
template <int SITE> __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow<SITE>(bar, parity);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// Tile load.  CG = 2: the copy may signal an mbarrier of the peer CTA (`bar` is then a shared::cluster address of the
// leader's barrier, obtained with mapa) -- both CTAs of a pair report their bytes to the leader's "full" barrier.
template <int CG>
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  if (CG == 2)
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// L2 prefetch of one box: no shared memory, no barrier -- the later cp.async.bulk.tensor of the same box then pays L2 latency
// instead of DRAM latency, which the 5-6 pipeline stages (160-192 KiB in flight per CTA) cannot hide on their own.
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2) : "memory");
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
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier of (possibly) the peer CTA.  Default semantics (release at CTA scope) on purpose: what the arrival
// publishes is the completion of this warp's tcgen05.ld reads (ordered by tcgen05.wait::ld + fence::before_thread_sync), not
// generic-proxy memory, and a cluster-scope release compiles to MEMBAR.ALL.GPU + ERRBAR (40 % of a thin layer's epilogue).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// TMEM allocation: CG = 2 is a collective of the same warp of both CTAs of the pair.
template <int CG> __device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  if (CG == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
}
template <int CG> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]; descriptors are passed as their low words (start address | LBO) plus the shared high word.
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  if (CG == 2)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// "all MMAs issued so far by this thread have completed" -> mbarrier arrive.  CG = 2: delivered to the barrier at the same
// offset in both CTAs of the pair.
template <int CG> __device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (CG == 2)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1, 128B swizzle), split into the
// low word (start address >> 4 | LBO << 16) and the high word (SBO | version | swizzle), which is the same for all operands.
//   K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused (1).
//   MN-major: 64-element (128 B) MN atoms x 8 k-rows; next 8 k-rows 1024 B on (SBO);
//             next MN atom one whole TMA box on = 64 k-rows * 128 B = 8192 B (LBO).
constexpr uint32_t UMMA_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, int major) {
  const uint32_t lbo = (major == MAJOR_K) ? 1u : (8192u >> 4);
  return ((saddr & 0x3FFFFu) >> 4) | (lbo << 16);
}
// tcgen05 instruction descriptor, kind::f16, A/B = bf16, D = fp32, M = 128 per CTA (256 for a CTA pair).
template <int CG> __device__ __forceinline__ uint32_t umma_idesc(int n, int a_major, int b_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// epilogue helpers
// ------------------------------------------------------------------------------------------------
template <int act> __device__ __forceinline__ float act_fwd(float v, float beta = 1.f) {
  switch (act) {
    case ACT_RELU:    return fmaxf(v, 0.f);
    case ACT_TANH:    return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case ACT_ELU:     return v > 0.f ? v : expm1f(v);
    case ACT_SWISH:   return v / (1.f + __expf(-beta * v));
    default:          return v;
  }
}
// d/d beta of x * sigmoid(beta x) = x^2 s (1 - s), s = sigmoid(beta x)  (x: the stored pre-activation)
__device__ __forceinline__ float swish_dbeta(float x, float beta) {
  const float sg = 1.f / (1.f + __expf(-beta * x));
  return x * x * sg * (1.f - sg);
}
// derivative of the activation expressed through what the forward pass stored: its OUTPUT y -- except for swish, whose
// derivative is not a function of the output: the forward pass keeps the PRE-activation x of swish layers (in the layer's
// gradient buffer, which the backward pass overwrites in place), and y here is that x:  d/dx x.s(x) = s(x) (1 + x (1 - s(x)))
template <int act> __device__ __forceinline__ float act_bwd_from_out(float y, float beta = 1.f) {
  switch (act) {
    case ACT_SWISH:   { const float sg = 1.f / (1.f + __expf(-beta * y)); return sg * (1.f + beta * y * (1.f - sg)); }
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
// relu fused into the conversion (negative -> +0)
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// two fp32 additions in one instruction (FADD2 on sm_100): (a0, a1) += (b0, b1)
__device__ __forceinline__ void fadd2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tadd.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
// per 16-bit half: 0xffff if the bf16 value is > 0, else 0
__device__ __forceinline__ uint32_t bf16x2_gt0(uint32_t a) {
  uint32_t d;
  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(0u));
  return d;
}
// ReLU sign-bit mask word of one 32-column chunk of one row.  Column c = 4 s + 2 h + e (s = 0..7, h, e = 0 / 1) lives in bit
// 8 (2 e + h) + 7 - s: the word shifted left by s has the bits of columns 4 s .. 4 s + 3 in the sign positions of bytes
// 0 (h=0,e=0), 1 (h=1,e=0), 2 (h=0,e=1), 3 (h=1,e=1), from where one prmt (sign-replicating mode) per bf16 pair makes the AND mask.
__device__ __forceinline__ constexpr uint32_t relu_mask_bit(int c) { return 1u << (8 * (2 * (c & 1) + ((c >> 1) & 1)) + 7 - (c >> 2)); }
// Column sums of a 32 x 32 bf16 slab (32 rows of 64 B, 64B swizzle) on the legacy tensor path: ones[16 x 32] . slab, i.e.
// per 8-column group one ldmatrix.x4.trans (lane l supplies the address of its own row's 16-byte piece) feeding two
// mma.sync m16n8k16.  Returns the sum of column `lane`.  `piece_addr(j)` = shared address of piece j of this lane's row.
template <typename F> __device__ __forceinline__ float slab_colsum32(F piece_addr, int lane) {
  float d[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t b0, b1, b2, b3;
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(piece_addr(j)));
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    const uint32_t one2 = 0x3F803F80u;            // bf16x2 (1, 1)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %4, %4, %4}, {%5, %6}, {%0, %1, %2, %3};"
                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3) : "r"(one2), "r"(b0), "r"(b1));
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %4, %4, %4}, {%5, %6}, {%0, %1, %2, %3};"
                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3) : "r"(one2), "r"(b2), "r"(b3));
    d[j][0] = c0; d[j][1] = c1;                   // columns 8 j + 2 (lane % 4) + {0, 1}, the same in every row group
  }
  // lane l keeps column (l / 8) * 8 + (l % 4) * 2 + ((l / 4) & 1): with l % 4 = t and (l / 4) & 1 = e that is d[l / 8][e]
  const int j = lane >> 3, e = (lane >> 2) & 1;
  float r = 0.f;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const float pick = e ? d[jj][1] : d[jj][0];
    r = (jj == j) ? pick : r;
  }
  return r;
}
// which column of a chunk slab_colsum32 leaves in `lane`
__device__ __forceinline__ int slab_colsum_col(int lane) { return ((lane >> 3) << 3) + ((lane & 3) << 1) + ((lane >> 2) & 1); }

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
                                               int col0, int nv, int row0, int lane, double& loss_local, float beta, float& dbeta_local) {
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
    if (ACT == ACT_SWISH && e.out2 && row_ok)      // swish: keep the pre-activation for the backward pass
      store_row32(e.out2 + (int64_t)row * e.out2_ld + col0, e.out2_ps, e.out2_planes, nv, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = act_fwd<ACT>(v[i], beta);
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
        if (ACT == ACT_SWISH && e.act_grad) {
#pragma unroll
          for (int i = 0; i < 32; ++i) dbeta_local += v[i] * swish_dbeta(y[i], beta);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= act_bwd_from_out<ACT>(y[i], beta);
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

// A run of consecutive k-blocks of one output tile that stay inside one (precision pass, K segment): the unit of the
// pipeline loops' outer level, so that their inner level is a plain counted loop with constant coordinate increments.
struct KRun {
  int pass, seg, r0, n;        // r0: first k-block inside the segment; n: k-blocks in the run
  bool ends_seg;               // the run contains the segment's last (possibly ragged) k-block
  __device__ __forceinline__ void init(int it, int it_end, int kb0, int kb1, int passes) {
    const int kb_total = kb0 + kb1;
    pass = passes == 1 ? 0 : it / kb_total;       // (no division in the single-pass bf16 mode)
    const int rem = it - pass * kb_total;
    seg = rem >= kb0 ? 1 : 0;
    r0 = seg ? rem - kb0 : rem;
    const int left = (seg ? kb1 : kb0) - r0;
    n = left < it_end - it ? left : it_end - it;
    ends_seg = (n == left);
  }
};

// PVAE_DBG bit 5: per-unit clock64 stamps of the three roles of every CTA (first TRACE_UNITS units), read back through
// pvae_debug_trace().  slots: 0 MMA loop top, 1 tempty acquired, 2 first operands landed, 3 last k-block issued,
// 4 epilogue before tfull wait, 5 accumulator ready, 6 epilogue done, 7 producer issued the unit's last copy.
// The hooks (role timeline + the epilogue-skipping PVAE_DBG switches) are compiled in only with -DPVAE_DEBUG_HOOKS: their tests,
// the trace index division and the stamps cost the hot epilogue loop ~15 % of its instructions.
#ifdef PVAE_DEBUG_HOOKS
constexpr bool DEBUG_HOOKS = true;
#else
constexpr bool DEBUG_HOOKS = false;
#endif
constexpr int TRACE_UNITS = 16, TRACE_CTAS = 160;
__device__ unsigned long long g_trace[TRACE_CTAS * TRACE_UNITS * 8];
__device__ __forceinline__ void trace_stamp(int dbg, int k, int slot) {
  if (DEBUG_HOOKS && (dbg & 32) && k < TRACE_UNITS && blockIdx.x < TRACE_CTAS) g_trace[((int)blockIdx.x * TRACE_UNITS + k) * 8 + slot] = clock64();
}

// The work units of one CTA (pair): u = unit0, unit0 + stride, ... numbered split-major / M-tile / N-tile-minor.  The
// (split, m, n) triple is stepped as a mixed-radix counter: integer divisions (~150 dependent cycles each, 64-bit ones
// several hundred) happen once per kernel instead of once per unit in every role -- for a 5-k-block tile they cost the
// MMA-issuing warp more time than the tile's tensor work.
// `reverse` walks the batch dimension (M pairs of a row-streaming GEMM, K splits of a weight gradient) from the far end: a kernel
// that starts where its producer / the previous reader of the same activation stopped finds that end still in the 126 MB L2.
struct UnitWalk {
  int split, mp, nt;
  int d_split, d_mp, d_nt, n_tiles, m_pairs, it_base, it_rem, n_splits, rev;
  __device__ __forceinline__ void init(int u0, int stride, int n_tiles_, int m_pairs_, int iters_total, int splits, int reverse) {
    n_tiles = n_tiles_; m_pairs = m_pairs_; n_splits = splits; rev = reverse;
    const int tile_units = n_tiles * m_pairs;
    split = u0 / tile_units;
    int t = u0 - split * tile_units;
    mp = t / n_tiles; nt = t - mp * n_tiles;
    d_split = stride / tile_units;
    t = stride - d_split * tile_units;
    d_mp = t / n_tiles; d_nt = t - d_mp * n_tiles;
    it_base = iters_total / splits; it_rem = iters_total - it_base * splits;
  }
  __device__ __forceinline__ void next() {
    nt += d_nt;
    int c = nt >= n_tiles ? 1 : 0;
    nt -= c ? n_tiles : 0;
    mp += d_mp + c;
    c = mp >= m_pairs ? 1 : 0;
    mp -= c ? m_pairs : 0;
    split += d_split + c;
  }
  __device__ __forceinline__ int m_pair() const { return (rev && n_splits == 1) ? m_pairs - 1 - mp : mp; }
  __device__ __forceinline__ int k_split() const { return (rev && n_splits > 1) ? n_splits - 1 - split : split; }
  // k-blocks of the current split: the first it_rem splits get one more than the others
  __device__ __forceinline__ int it_begin() const { const int sp = k_split(); return sp * it_base + (sp < it_rem ? sp : it_rem); }
  __device__ __forceinline__ int it_end() const { const int sp = k_split(); return it_begin() + it_base + (sp < it_rem ? 1 : 0); }
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// TMAEPI: the primary bf16 output (one precision plane) leaves through shared memory and cp.async.bulk.tensor stores, the
// aux operand (forward activation for act', MSE target) arrives the same way; the epilogue warps then touch only TMEM,
// shared memory and registers.  Without it (bf16x3 planes, fp32-only outputs) rows are read / written directly.
// CG: CTAs per MMA instruction (see the top of the file); the launch uses clusters of CG CTAs.
// FAST (TMAEPI, ReLU, EPI_STORE / EPI_DGRAD only): the host guarantees that every row and every 32-column chunk of every tile
// exists (M a multiple of the M tile of the CTA pair, N a multiple of 32), that bias / mask are present and that there is no
// addend, fp32 copy, M segment or gap -- the epilogue then runs without any of the per-chunk tests (see the FAST block below).
template <int EPI, int ACT, bool TMAEPI, int CG, bool FAST = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) pvae_gemm_kernel(const __grid_constant__ GemmParams p) {
  static_assert(!FAST || (TMAEPI && ACT == ACT_RELU && (EPI == EPI_STORE || EPI == EPI_DGRAD)), "FAST epilogue: ReLU store / dgrad through TMA only");
  // The lean epilogue trades one pipeline stage (5 instead of 6 -- 160 KiB of operands in flight per CTA is still twice what the
  // L2 -> SM feed rate times the TMA latency needs) for a second staging slab per epilogue warp, so that a warp waits for the TMA
  // store issued TWO chunks ago, not for the one it has just issued (the stores queue behind the operand loads in the SM's single
  // TMA unit).  Barrier block and bias slices sit at the same offsets in both layouts.
  constexpr int STAGES = FAST ? 5 : Geo<CG>::STAGES;
  constexpr int STAGE_BYTES = Geo<CG>::STAGE_BYTES;
  constexpr int WARP_SLAB = FAST ? 2 * SLAB_BYTES : SLAB_BYTES;
  constexpr int OFF_STG = FAST ? OFF_BARS - EPI_WARPS * WARP_SLAB : OFF_STAGING;
  static_assert(STAGES * STAGE_BYTES <= OFF_STG, "pipeline overlaps the staging slabs");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 128B swizzle atoms need 1024 B alignment
  if (smem_base - smem_u32(smem_raw) > 512u) {                        // the carve-out assumes at most 512 B of alignment slack
    if (threadIdx.x == 0) printf("pvae_gemm: dynamic shared memory base 0x%x leaves too little room after 1 KiB alignment\n", smem_u32(smem_raw));
    __trap();
  }
  const uint32_t bar_base = smem_base + OFF_BARS;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + ACC_STAGES + s); };
  auto auxfull_bar = [&](int w) { return bar_base + 8u * (2 * MAX_STAGES + 2 * ACC_STAGES + w); };   // one per epilogue warp
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 2 * ACC_STAGES + EPI_WARPS);
  static_assert(8 * (2 * MAX_STAGES + 2 * ACC_STAGES + EPI_WARPS + 1) <= 512, "barrier block");
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
    // full: one arrive.expect_tx by the (leader's) producer + the bytes of both CTAs; empty / tfull: one tcgen05.commit
    // (delivered to both CTAs for CG = 2); tempty (used in the leader only): every epilogue warp of the pair.
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), CG * EPI_WARPS); }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(auxfull_bar(w), 1);
    fence_barrier_init();
  }
  if (warp == 1) { __syncwarp(); tmem_alloc<CG>(tmem_slot, TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  if (CG > 1) cluster_sync_all();               // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  constexpr int csize = CG;
  const int crank = CG > 1 ? (int)cluster_ctarank() : 0;
  const int unit0 = blockIdx.x / csize, unit_stride = gridDim.x / csize;

  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, cluster sync) touched
  // no global memory, so it may run while the previous kernel of the stream is still draining -- a CTA of this grid becomes
  // resident the moment the CTA of the previous grid on its SM exits.  launch_dependents lets the NEXT grid do the same with
  // us; wait blocks until the previous grid has completed and its writes are visible.  Nothing below this line may move up.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int row0 = p.row_cursor ? *p.row_cursor : 0;
  const int kb_total = p.kb[0] + p.kb[1];
  const int iters_total = p.passes * kb_total;
  // units of one CTA (pair): tile-minor / split-major, so that the CTAs running at the same time share K ranges through L2
  const int m_pairs = (p.m_tiles + csize - 1) / csize;
  const int tile_units = m_pairs * p.n_tiles;
  const int total_units = tile_units * p.splits;
  const int bn = p.bn;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // The whole warp walks the loops (warp-uniform control flow); one elected lane issues the copies.  Each CTA stages
    // its own 128 rows of A and its 1/CG share of the B tile; for CG = 2 both CTAs' bytes are reported to the leader's
    // "full" barrier, whose single arrival (with the byte count of the pair) comes from the leader's producer.
    const bool a_mn = p.a_major == MAJOR_MN, b_mn = p.b_major == MAJOR_MN;
    const int bn_loc = bn / CG;                           // B rows (n) staged by this CTA
    const int b_boxes = b_mn ? (bn_loc + 63) >> 6 : 1;
    const uint32_t stage_tx = (uint32_t)CG * (A_STAGE_BYTES + (b_mn ? (uint32_t)b_boxes * 8192u : (uint32_t)bn_loc * (BK * 2)));
    const bool leader = (CG == 1) || crank == 0;
    const uint32_t full0 = (CG == 2) ? mapa_u32(full_bar(0), 0u) : full_bar(0);
    int stage = 0; uint32_t phase = 0;
    UnitWalk w; w.init(unit0, unit_stride, p.n_tiles, m_pairs, iters_total, p.splits, p.reverse);
    // Operand boxes of the unit `wp` is at, as L2 prefetches (same coordinate arithmetic as the copy loop below).
    auto prefetch_unit = [&](const UnitWalk& wp) {
      const int m_tile = wp.m_pair() * csize + crank;
      const int it_end = wp.it_end();
      int it = wp.it_begin();
      const int b_n = p.b_n0 + wp.nt * bn + crank * bn_loc;
      while (it < it_end) {
        KRun run; run.init(it, it_end, p.kb[0], p.kb[1], p.passes);
        it += run.n;
        const int seg = run.seg;
        const int a_plane = (run.pass == 2) ? 1 : 0;
        const int b_plane = (run.pass == 1) ? 1 : 0;
        const int aseg = (p.m_seg_tiles && m_tile >= p.m_seg_tiles) ? 1 : seg;
        const int m_loc = (p.m_seg_tiles && m_tile >= p.m_seg_tiles) ? m_tile - p.m_seg_tiles : m_tile;
        const CUtensorMap* tmA = &p.tmA[aseg];
        const int a_row = p.a_r0[aseg] + (p.a_dyn[aseg] ? row0 : 0);
        int ac0 = a_mn ? p.a_c0[aseg] + m_loc * BM : p.a_c0[aseg] + run.r0 * BK;
        int ac1 = a_mn ? a_row + run.r0 * BK : a_row + m_loc * BM;
        int bc0 = b_mn ? b_n : p.b_k0[seg] + run.r0 * BK;
        int bc1 = b_mn ? p.b_k0[seg] + run.r0 * BK + (p.b_dyn ? row0 : 0) : b_n;
        for (int j = 0; j < run.n; ++j) {
          if (elect_one()) {
            tma_prefetch_3d(tmA, ac0, ac1, a_plane);
            if (a_mn) tma_prefetch_3d(tmA, ac0 + 64, ac1, a_plane);
            if (p.pf_b) {
              tma_prefetch_3d(&p.tmB, bc0, bc1, b_plane);
              for (int x = 1; x < b_boxes; ++x) tma_prefetch_3d(&p.tmB, bc0 + 64 * x, bc1, b_plane);
            }
          }
          __syncwarp();
          if (a_mn) ac1 += BK; else ac0 += BK;
          if (b_mn) bc1 += BK; else bc0 += BK;
        }
      }
    };
    // (experiment, compiled only into the -DPVAE_DEBUG_HOOKS build: measured 10-15 % slower, profiles/r02_bench.md)
    UnitWalk wpf = w;                                      // runs pf_dist units ahead of w
    int upf = unit0;
    if (DEBUG_HOOKS && p.pf_dist > 0) {
      for (int d = 0; d < p.pf_dist; ++d) {
        if (d > 0 && upf < total_units) prefetch_unit(wpf);   // units 1 .. pf_dist-1: nobody prefetches them later
        upf += unit_stride; wpf.next();
      }
    }
    for (int u = unit0; u < total_units; u += unit_stride, w.next()) {
      const int n_tile = w.nt;
      const int m_tile = w.m_pair() * csize + crank;
      const int it_end = w.it_end();
      int it = w.it_begin();
      const int b_n = p.b_n0 + n_tile * bn + crank * bn_loc;
      if (DEBUG_HOOKS && p.pf_dist > 0) {
        if (upf < total_units) prefetch_unit(wpf);
        upf += unit_stride; wpf.next();
      }
      while (it < it_end) {
        KRun run; run.init(it, it_end, p.kb[0], p.kb[1], p.passes);
        it += run.n;
        const int seg = run.seg;
        const int a_plane = (run.pass == 2) ? 1 : 0;
        const int b_plane = (run.pass == 1) ? 1 : 0;
        const int aseg = (p.m_seg_tiles && m_tile >= p.m_seg_tiles) ? 1 : seg;     // M segment (wgrad) or K segment
        const int m_loc = (p.m_seg_tiles && m_tile >= p.m_seg_tiles) ? m_tile - p.m_seg_tiles : m_tile;
        const CUtensorMap* tmA = &p.tmA[aseg];
        const int a_row = p.a_r0[aseg] + (p.a_dyn[aseg] ? row0 : 0);
        // coordinates of the run's first k-block; a k-block further on moves the K coordinate by BK
        int ac0 = a_mn ? p.a_c0[aseg] + m_loc * BM : p.a_c0[aseg] + run.r0 * BK;
        int ac1 = a_mn ? a_row + run.r0 * BK : a_row + m_loc * BM;
        int bc0 = b_mn ? b_n : p.b_k0[seg] + run.r0 * BK;
        int bc1 = b_mn ? p.b_k0[seg] + run.r0 * BK + (p.b_dyn ? row0 : 0) : b_n;
        for (int j = 0; j < run.n; ++j) {
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          mbar_wait<W_EMPTY>(empty_bar(stage), phase ^ 1u);
          if (elect_one()) {
            if (leader) mbar_expect_tx(full_bar(stage), stage_tx);
            const uint32_t fb = full0 + 8u * stage;
            tma_load_3d<CG>(sa, tmA, fb, ac0, ac1, a_plane);
            if (a_mn) tma_load_3d<CG>(sa + 8192u, tmA, fb, ac0 + 64, ac1, a_plane);
            tma_load_3d<CG>(sb, &p.tmB, fb, bc0, bc1, b_plane);
            for (int x = 1; x < b_boxes; ++x) tma_load_3d<CG>(sb + 8192u * x, &p.tmB, fb, bc0 + 64 * x, bc1, b_plane);
          }
          __syncwarp();
          if (a_mn) ac1 += BK; else ac0 += BK;
          if (b_mn) bc1 += BK; else bc0 += BK;
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (DEBUG_HOOKS && lane == 0) trace_stamp(p.dbg, (u - unit0) / unit_stride, 7);
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA of the pair only) ================================
    if (CG == 1 || crank == 0) {
      const uint32_t idesc = umma_idesc<CG>(bn, p.a_major, p.b_major);
      const uint32_t a_kstep = (p.a_major == MAJOR_K) ? 2u : 128u;   // 16 k elements, in 16 B units: 32 B or 16 rows * 128 B
      const uint32_t b_kstep = (p.b_major == MAJOR_K) ? 2u : 128u;
      const uint32_t a_lo0 = umma_desc_lo(smem_base, p.a_major);
      const uint32_t b_lo0 = umma_desc_lo(smem_base + A_STAGE_BYTES, p.b_major);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      // one k-block: wait for its operands, issue k16 (1..4) instructions of K = 16, free the slot when they have read it
      auto kblock = [&](uint32_t tmem_d, uint32_t accum, int k16) {
        mbar_wait<W_FULL>(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t so = (uint32_t)(stage * STAGE_BYTES) >> 4;
          const uint32_t a = a_lo0 + so, b = b_lo0 + so;
          umma_bf16<CG>(tmem_d, a, b, UMMA_DESC_HI, idesc, accum);
          if (k16 > 1) umma_bf16<CG>(tmem_d, a + a_kstep, b + b_kstep, UMMA_DESC_HI, idesc, 1u);
          if (k16 > 2) umma_bf16<CG>(tmem_d, a + 2 * a_kstep, b + 2 * b_kstep, UMMA_DESC_HI, idesc, 1u);
          if (k16 > 3) umma_bf16<CG>(tmem_d, a + 3 * a_kstep, b + 3 * b_kstep, UMMA_DESC_HI, idesc, 1u);
          umma_commit<CG>(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      };
      UnitWalk w; w.init(unit0, unit_stride, p.n_tiles, m_pairs, iters_total, p.splits, p.reverse);
      for (int u = unit0; u < total_units; u += unit_stride, w.next()) {
        const int it_end = w.it_end();
        int it = w.it_begin();
        const int tk = (DEBUG_HOOKS && (p.dbg & 32)) ? (u - unit0) / unit_stride : 0;
        if (lane == 0) trace_stamp(p.dbg, tk, 0);
        mbar_wait<W_TEMPTY>(tempty_bar(acc), acc_phase ^ 1u);          // every epilogue warp (of both CTAs) has drained this stage
        tc_fence_after();
        if (lane == 0) trace_stamp(p.dbg, tk, 1);
        bool first_kb = true;
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * MAX_BN;
        uint32_t accum = 0u;
        while (it < it_end) {
          KRun run; run.init(it, it_end, p.kb[0], p.kb[1], p.passes);
          it += run.n;
          // all-zero k16 slices at the ragged end of a segment are skipped
          int tail = 4;
          if (run.ends_seg) {
            tail = (p.klen[run.seg] - (p.kb[run.seg] - 1) * BK + 15) >> 4;
            tail = tail > 4 ? 4 : (tail < 1 ? 1 : tail);
          }
          const int n_full = tail < 4 ? run.n - 1 : run.n;
          if (DEBUG_HOOKS && (p.dbg & 32) && first_kb) {          // (trace only) when did the unit's first operands land?
            mbar_wait<W_FULL>(full_bar(stage), phase);
            if (lane == 0) trace_stamp(p.dbg, tk, 2);
            first_kb = false;
          }
          for (int j = 0; j < n_full; ++j) { kblock(tmem_d, accum, 4); accum = 1u; }
          if (tail < 4) { kblock(tmem_d, accum, tail); accum = 1u; }
        }
        if (lane == 0) trace_stamp(p.dbg, tk, 3);
        if (elect_one()) umma_commit<CG>(tfull_bar(acc));    // accumulator complete -> epilogue (of both CTAs)
        __syncwarp();
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if constexpr (FAST) {
    // ================================ epilogue, lean form (16 warps) ================================
    // Same division of labour as the general epilogue below (warp w: TMEM lane quarter w % 4, 32-column chunks cgrp, cgrp + 4 of
    // the tile, private 2 KiB slab, one TMA store per chunk), minus everything the host has ruled out: no row / column edge
    // tests, no addend, no fp32 copy, no aux operand.  Two things are done differently because this path is what bounds every
    // layer with K <= 512 (the 16 warps need longer for a 128 x 256 tile than the tensor pipe):
    //   * the ReLU-mask words / bias values of the NEXT unit are fetched while the current one is processed (a load issued
    //     right before the accumulator wait is not hidden when the accumulator is already there);
    //   * bias-gradient column sums read the staged slab as 32-bit words -- lanes 0-15 rows 0-15, lanes 16-31 rows 16-31,
    //     two columns each, opposite row parity per half so that one LDS never hits a bank twice -- and add with FADD2.
    const EpiParams& e = p.epi;
    const int q = warp & 3;
    const int ew = warp - 2;
    const int cgrp = ew >> 2;
    float* bias_s = bias_all + ew * 32;
    const int n_valid = e.n_valid;
    int acc = 0; uint32_t acc_phase = 0;
    constexpr int CS_TILES = 4;
    float cs_acc[4 * CS_TILES];                    // [n tile][chunk J][column 2 w, 2 w + 1]
#pragma unroll
    for (int k = 0; k < 4 * CS_TILES; ++k) cs_acc[k] = 0.f;
    const bool do_cs = EPI == EPI_DGRAD && e.colsum != nullptr;
    const uint32_t tempty0 = (CG == 2) ? mapa_u32(tempty_bar(0), 0u) : tempty_bar(0);
    const uint32_t slab0 = smem_base + OFF_STG + ew * WARP_SLAB;           // two 2 KiB slabs per warp, used alternately
    uint8_t* slab0_gen = smem_gen + OFF_STG + ew * WARP_SLAB;
    int sbuf = 0;
    const int sw = (lane >> 1) & 3;
    const int hh = lane >> 4, ww = lane & 15;     // column-sum role: row half, word (= column pair) of the 64-byte slab row
    UnitWalk w; w.init(unit0, unit_stride, p.n_tiles, m_pairs, iters_total, p.splits, p.reverse);
    // operands of a unit's (at most two) chunks that do not depend on the accumulator
    auto fetch = [&](const UnitWalk& uw, float (&pb)[2], uint32_t (&pm)[2]) {
      const int nt = uw.nt;
      const int r = (uw.m_pair() * csize + crank) * BM + q * 32 + lane;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int cs = nt * bn + (cgrp + 4 * j) * 32;
        pb[j] = 0.f; pm[j] = 0u;
        if ((cgrp + 4 * j) * 32 < bn && cs < n_valid) {
          if (EPI == EPI_STORE) pb[j] = __ldg(e.bias + cs + lane);
          else pm[j] = __ldg(e.mask + (int64_t)(cs >> 5) * e.mask_ld + r);
        }
      }
    };
    float pre_bias[2]; uint32_t pre_mask[2];
    if (unit0 < total_units) fetch(w, pre_bias, pre_mask);
    for (int u = unit0; u < total_units; u += unit_stride, w.next()) {
      const int n_tile = w.nt;
      const int m_tile = w.m_pair() * csize + crank;
      const int row = m_tile * BM + q * 32 + lane;
      float nxt_bias[2] = {0.f, 0.f}; uint32_t nxt_mask[2] = {0u, 0u};
      if (u + unit_stride < total_units) { UnitWalk wn = w; wn.next(); fetch(wn, nxt_bias, nxt_mask); }
      const int tk = (DEBUG_HOOKS && (p.dbg & 32)) ? (u - unit0) / unit_stride : 0;
      if (DEBUG_HOOKS && ew == 0 && lane == 0) trace_stamp(p.dbg, tk, 4);
      mbar_wait<W_TFULL>(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (DEBUG_HOOKS && ew == 0 && lane == 0) trace_stamp(p.dbg, tk, 5);
      const uint32_t tmem_unit = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * MAX_BN + cgrp * 32);
      const int col_unit = n_tile * bn + cgrp * 32;
      auto chunk = [&](auto jc) {
        constexpr int J = decltype(jc)::value;
        const int col0 = col_unit + 128 * J;
        if ((cgrp + 4 * J) * 32 >= bn || col0 >= n_valid) return;          // warp-uniform
        float v[32];
        {
          uint32_t raw[32];
          tmem_ld32(tmem_unit + 128u * J, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        }
        uint32_t pk[16];
        if (EPI == EPI_STORE) {
          __syncwarp();
          bias_s[lane] = pre_bias[J];             // lane -> bias of column lane of the chunk
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + g * 4);
            fadd2(v[g * 4 + 0], v[g * 4 + 1], b4.x, b4.y);
            fadd2(v[g * 4 + 2], v[g * 4 + 3], b4.z, b4.w);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2_relu(v[2 * i], v[2 * i + 1]);
          uint32_t m = 0u;
#pragma unroll
          for (int i = 0; i < 16; ++i) m |= bf16x2_gt0(pk[i]) & (relu_mask_bit(2 * i) | relu_mask_bit(2 * i + 1));
          e.mask[(int64_t)(col0 >> 5) * e.mask_ld + row] = m;
        } else {
          const uint32_t mbits = pre_mask[J];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
#pragma unroll
          for (int s = 0; s < 8; ++s) {           // ReLU': see relu_mask_bit()
            const uint32_t t = mbits << s;
            uint32_t m0, m1;
            asm("prmt.b32 %0, %1, %1, 0xAA88;" : "=r"(m0) : "r"(t));
            asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m1) : "r"(t));
            pk[2 * s] &= m0;
            pk[2 * s + 1] &= m1;
          }
        }
        const uint32_t slab = slab0 + (uint32_t)sbuf * SLAB_BYTES;
        uint8_t* slab_gen = slab0_gen + sbuf * SLAB_BYTES;
        uint8_t* srow = slab_gen + lane * 64;
        sbuf ^= 1;
        tma_store_wait_read<1>();                 // the store issued from THIS slab two chunks ago has been read out (groups are per thread, in order)
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(srow + ((g ^ sw) << 4)) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
          tma_store_3d(&p.tmOut, slab, col0, m_tile * BM + q * 32, 0);
          tma_store_commit();
        }
        __syncwarp();
        if (do_cs) {
          // rows 2 i, 2 i + 1 of this lane's half share the swizzle term i & 3; half 0 reads the even row first, half 1 the odd one
          const uint8_t* cs_base = slab_gen + (hh << 10) + ((ww & 3) << 2);
          uint32_t wa[8], wb[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int po = (((ww >> 2) ^ (i & 3)) << 4);
            wa[i] = *reinterpret_cast<const uint32_t*>(cs_base + i * 128 + (hh << 6) + po);
            wb[i] = *reinterpret_cast<const uint32_t*>(cs_base + i * 128 + ((hh ^ 1) << 6) + po);
          }
          float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            fadd2(s0, s1, __uint_as_float(wa[i] << 16), __uint_as_float(wa[i] & 0xFFFF0000u));
            fadd2(t0, t1, __uint_as_float(wb[i] << 16), __uint_as_float(wb[i] & 0xFFFF0000u));
          }
          fadd2(s0, s1, t0, t1);
          s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          if (n_tile < CS_TILES) {
#pragma unroll
            for (int t = 0; t < CS_TILES; ++t) {
              cs_acc[4 * t + 2 * J]     += (n_tile == t) ? s0 : 0.f;
              cs_acc[4 * t + 2 * J + 1] += (n_tile == t) ? s1 : 0.f;
            }
          } else if (hh == 0) {
            atomicAdd(e.colsum + col0 + 2 * ww, s0);
            atomicAdd(e.colsum + col0 + 2 * ww + 1, s1);
          }
          __syncwarp();
        }
      };
      chunk(std::integral_constant<int, 0>{});
      chunk(std::integral_constant<int, 1>{});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty0 + 8u * acc);
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      if (DEBUG_HOOKS && ew == 0 && lane == 0) trace_stamp(p.dbg, tk, 6);
      pre_bias[0] = nxt_bias[0]; pre_bias[1] = nxt_bias[1];
      pre_mask[0] = nxt_mask[0]; pre_mask[1] = nxt_mask[1];
    }
    tma_store_wait_read<0>();
    if (do_cs) {
      // flush: per (n tile, chunk) 32 column sums in natural order; the four lane-quarter warps of a column group combine
      // through their idle slabs, one warp per column group issues the reds
      __syncwarp();
      float* mine = reinterpret_cast<float*>(slab0_gen);
      if (hh == 0) {
#pragma unroll
        for (int k = 0; k < 2 * CS_TILES; ++k)
          *reinterpret_cast<float2*>(mine + k * 32 + 2 * ww) = make_float2(cs_acc[2 * k], cs_acc[2 * k + 1]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      if ((ew & 3) == 0) {
#pragma unroll
        for (int k = 0; k < 2 * CS_TILES; ++k) {
          float t = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            t += reinterpret_cast<const float*>(smem_gen + OFF_STG + (cgrp * 4 + qq) * WARP_SLAB)[k * 32 + lane];
          const int cl = (cgrp + 4 * (k & 1)) * 32 + lane;
          const int col = (k >> 1) * bn + cl;
          if (cl < bn && (k >> 1) < p.n_tiles && col < n_valid) atomicAdd(e.colsum + col, t);
        }
      }
    }
  } else {
    // ================================ epilogue (16 warps) ================================
    // Warp w owns the 32 rows of TMEM lane quarter w % 4 (a hardware rule) and the 32-column chunks c = (w - 2) / 4 + 4 j of
    // the tile: four warps per scheduler hide each other's TMEM / shared-memory / fence latencies.
    const EpiParams& e = p.epi;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read (warp id % 4)
    const int ew = warp - 2;                      // epilogue warp index 0..15
    const int cgrp = ew >> 2;                     // column group: chunks cgrp, cgrp + 4 (its four warps cover the four quarters)
    float* bias_s = bias_all + ew * 32;
    const int m_valid = e.m_valid, n_valid = e.n_valid;
    const float* bias = (EPI == EPI_STORE || EPI == EPI_MSE) ? e.bias : nullptr;
    int acc = 0; uint32_t acc_phase = 0;
    double loss_local = 0.0;
    const float beta = (ACT == ACT_SWISH && e.act_param) ? __ldg(e.act_param) : 1.f;
    float dbeta_local = 0.f;
    uint32_t aux_phase = 0;
    const bool cs_mma = p.cs_mma > 0;              // see the column-sum code below
    constexpr int CS_TILES = 4;                   // N tiles whose bias-gradient column sums are kept in registers
    float cs_acc[2 * CS_TILES];
#pragma unroll
    for (int k = 0; k < 2 * CS_TILES; ++k) cs_acc[k] = 0.f;
    const uint32_t tempty0 = (CG == 2) ? mapa_u32(tempty_bar(0), 0u) : tempty_bar(0);
    const uint32_t slab = smem_base + OFF_STAGING + ew * SLAB_BYTES;      // this warp's private staging slab
    uint8_t* slab_gen = smem_gen + OFF_STAGING + ew * SLAB_BYTES;
    // 32 rows x 64 B, 64B swizzle: the 16-byte piece j of row r sits at piece j ^ ((r >> 1) & 3)
    uint8_t* srow = slab_gen + lane * 64;
    const int sw = (lane >> 1) & 3;
    UnitWalk w; w.init(unit0, unit_stride, p.n_tiles, m_pairs, iters_total, p.splits, p.reverse);
    for (int u = unit0; u < total_units; u += unit_stride, w.next()) {
      const int n_tile = w.nt;
      const int m_tile = w.m_pair() * csize + crank;
      const int row_in_tile = q * 32 + lane;
      int row = m_tile * BM + row_in_tile;
      bool row_ok = row < m_valid;
      if (p.m_seg_tiles) {                         // output row of an M-segmented (virtual concat) operand
        const int s1 = m_tile >= p.m_seg_tiles ? 1 : 0;
        const int local = row - (s1 ? p.m_seg_tiles * BM : 0);
        row_ok = local < p.m_seg_rows[s1];
        row = (s1 ? p.m_seg_out0 : 0) + local;
      }
      if (p.m_gap && row >= p.m_gap0) {            // (s_t | gap | a_t) input rows of the world model's layer-0 weight gradient
        row_ok = row_ok && row >= p.m_gap0 + p.m_gap;
        row -= p.m_gap;
      }
      // operands of this warp's (at most two) chunks that do not depend on the accumulator: fetched while the main loop runs
      float pre_bias[2] = {0.f, 0.f};
      uint32_t pre_mask[2] = {0u, 0u};
      if (TMAEPI) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int cs = n_tile * bn + (cgrp + 4 * j) * 32;
          if ((cgrp + 4 * j) * 32 < bn && cs < n_valid) {
            if (bias) pre_bias[j] = (cs + lane < n_valid) ? __ldg(bias + cs + lane) : 0.f;
            if (USE_MASK && row_ok) pre_mask[j] = __ldg(e.mask + (int64_t)(cs >> 5) * e.mask_ld + row);
          }
        }
      }
      if (HAS_AUX) {
        // fetch the aux slab of this warp's first chunk while the main loop of the tile is still running
        const int col0s = n_tile * bn + cgrp * 32;
        if (col0s < n_valid && cgrp * 32 < bn) {
          tma_store_wait_read<0>();               // the slab's previous store has been read out (groups are per thread)
          __syncwarp();
          if (elect_one()) {
            mbar_expect_tx(auxfull_bar(ew), SLAB_BYTES);
            tma_load_3d<1>(slab, &p.tmAux, auxfull_bar(ew), col0s, (e.aux_dyn ? row0 : 0) + m_tile * BM + q * 32, 0);
          }
          __syncwarp();
        }
      }
      const int tk = (DEBUG_HOOKS && (p.dbg & 32)) ? (u - unit0) / unit_stride : 0;
      if (ew == 0 && lane == 0) trace_stamp(p.dbg, tk, 4);
      mbar_wait<W_TFULL>(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (ew == 0 && lane == 0) trace_stamp(p.dbg, tk, 5);
      const int nchunks = (bn + 31) >> 5;
      if (!TMAEPI) {
        for (int c = cgrp; c < nchunks; c += 4) {
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
          epilogue_chunk<EPI, ACT>(e, v, bias_s, row, row_ok, col0, nv, row0, lane, loss_local, beta, dbeta_local);
        }
      } else {
        // Each warp stages its 32 x 32 bf16 chunk in a private 2 KiB slab and stores it with one TMA instruction -- no
        // cross-warp synchronisation.  The (at most two) chunks of a warp are two compile-time instances of one body, so that
        // their prefetched operands and bias-gradient accumulators are plain registers.
        const uint32_t tmem_unit = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * MAX_BN + cgrp * 32);
        const int col_unit = n_tile * bn + cgrp * 32;
        const int arow0 = (e.aux_dyn ? row0 : 0) + m_tile * BM + q * 32;      // first row of this warp's slab in the aux / output tensor
        auto chunk = [&](auto jc) {
          constexpr int J = decltype(jc)::value;
          const int c = cgrp + 4 * J;
          const int col0 = col_unit + 128 * J;
          if (c >= nchunks || col0 >= n_valid) return;              // warp-uniform
          int nv = n_valid - col0; nv = nv > 32 ? 32 : nv;
          const int tile_nv = bn - c * 32;        // columns of this chunk that belong to this tile
          if (tile_nv < nv) nv = tile_nv;
          const uint32_t mbits = pre_mask[J];
          if (HAS_AUX && J != 0) {                // second chunk of the tile: the aux load is exposed
            tma_store_wait_read<0>();
            __syncwarp();
            if (elect_one()) {
              mbar_expect_tx(auxfull_bar(ew), SLAB_BYTES);
              tma_load_3d<1>(slab, &p.tmAux, auxfull_bar(ew), col0, arow0, 0);
            }
            __syncwarp();
          }
          float v[32];
          {
            uint32_t raw[32];
            if (!(DEBUG_HOOKS && (p.dbg & 16))) {
              tmem_ld32(tmem_unit + 128u * J, raw);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) raw[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          }
          if (bias && !(DEBUG_HOOKS && (p.dbg & 1))) {
            __syncwarp();
            bias_s[lane] = pre_bias[J];           // lane -> bias of column lane of the chunk
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + g * 4);
              fadd2(v[g * 4 + 0], v[g * 4 + 1], b4.x, b4.y);
              fadd2(v[g * 4 + 2], v[g * 4 + 3], b4.z, b4.w);
            }
          }
          uint32_t pk[16];                        // the chunk's row as packed bf16 pairs
          if (EPI == EPI_STORE) {
            if (ACT != ACT_RELU || e.out_f32 != nullptr || nv < 32) {   // (warp-uniform) general path
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = act_fwd<ACT>(v[i], beta);
              if (nv < 32) {                      // ragged last chunk: keep the padding columns of the mask / slab zero
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (i < nv) ? v[i] : 0.f;
              }
              if (row_ok && e.out_f32) store_f32_row32(e.out_f32 + (int64_t)row * e.f32_sm + (int64_t)col0 * e.f32_sn, e.f32_sn, nv, v);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = (ACT == ACT_RELU) ? pack_bf16x2_relu(v[2 * i], v[2 * i + 1]) : pack_bf16x2(v[2 * i], v[2 * i + 1]);
            if (ACT == ACT_RELU && e.mask && row_ok && !(DEBUG_HOOKS && (p.dbg & 2))) {
              // sign-bit mask from the packed results: one packed compare + one (a & imm) | m per pair; layout: relu_mask_bit()
              uint32_t m = 0u;
#pragma unroll
              for (int i = 0; i < 16; ++i) m |= bf16x2_gt0(pk[i]) & (relu_mask_bit(2 * i) | relu_mask_bit(2 * i + 1));
              e.mask[(int64_t)(col0 >> 5) * e.mask_ld + row] = m;
            }
          } else {
            // aux values (MSE target / forward activation) are consumed piece by piece to keep the register footprint small
            auto aux_piece = [&](int g, float (&y)[8]) {
              const uint4 qv = *reinterpret_cast<const uint4*>(srow + ((g ^ sw) << 4));
              const uint32_t w4[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                y[2 * t]     = __uint_as_float(w4[t] << 16);
                y[2 * t + 1] = __uint_as_float(w4[t] & 0xFFFF0000u);
              }
            };
            if (HAS_AUX) {
              mbar_wait<W_AUX>(auxfull_bar(ew), aux_phase);
              aux_phase ^= 1u;
            }
            if (EPI == EPI_MSE) {
              if (row_ok) {
                if (e.out_f32) store_f32_row32(e.out_f32 + (int64_t)row * e.f32_sm + (int64_t)col0 * e.f32_sn, e.f32_sn, nv, v);
                if (e.out2) store_row32(e.out2 + (int64_t)row * e.out2_ld + col0, e.out2_ps, e.out2_planes, nv, v);
              }
              float sq = 0.f;
              const float scale = e.scale;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float y[8];
                aux_piece(g, y);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                  const int i = g * 8 + t;
                  const float di = (row_ok && i < nv) ? (v[i] - y[t]) : 0.f;
                  sq += di * di;
                  v[i] = scale * di;
                }
              }
              loss_local += (double)sq;
            } else {  // EPI_DGRAD
              if (e.add && row_ok) {
                float a[32];
                load_row32(e.add + (int64_t)row * e.add_ld + col0, e.add_ps, e.add_planes, nv, a);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += a[i];
              }
              if (USE_MASK && e.out_f32 != nullptr) {   // (an fp32 copy wants the masked values themselves)
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (mbits & relu_mask_bit(i)) ? v[i] : 0.f;   // mask is 0 for rows >= m_valid
              } else if (!USE_MASK && ACT != ACT_LINEAR) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  float y[8];
                  aux_piece(g, y);
                  if (ACT == ACT_SWISH && e.act_grad) {      // (rows / columns outside the tensor hold zero pre-activations: no contribution)
#pragma unroll
                    for (int t = 0; t < 8; ++t) dbeta_local += ((row_ok && g * 8 + t < nv) ? v[g * 8 + t] : 0.f) * swish_dbeta(y[t], beta);
                  }
#pragma unroll
                  for (int t = 0; t < 8; ++t) v[g * 8 + t] *= act_bwd_from_out<ACT>(y[t], beta);
                }
              }
              if (nv < 32 || (!USE_MASK && !row_ok)) {   // keep what the bias-gradient column sums must not see at zero
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (row_ok && i < nv) ? v[i] : 0.f;
              }
              if (row_ok && e.out_f32) store_f32_row32(e.out_f32 + (int64_t)row * e.f32_sm + (int64_t)col0 * e.f32_sn, e.f32_sn, nv, v);
            }
            if (HAS_AUX) __syncwarp();            // every lane has read the aux slab before it is overwritten below
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            if (USE_MASK && !(DEBUG_HOOKS && (p.dbg & 2))) {
              // ReLU': the mask word shifted left by s carries the bits of pairs 2 s / 2 s + 1 in the byte sign positions
              // (relu_mask_bit()); prmt in sign-replicating mode turns two of them into the AND mask of a bf16 pair
#pragma unroll
              for (int s = 0; s < 8; ++s) {
                const uint32_t t = mbits << s;
                uint32_t m0, m1;
                asm("prmt.b32 %0, %1, %1, 0xAA88;" : "=r"(m0) : "r"(t));
                asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m1) : "r"(t));
                pk[2 * s] &= m0;
                pk[2 * s + 1] &= m1;
              }
            }
          }
          if (DEBUG_HOOKS && (p.dbg & 4)) return;
          if (!HAS_AUX) {                         // (with aux the slab was already claimed before the aux load)
            tma_store_wait_read<0>();             // bulk groups are per thread: only the electing lane ever has pending ones
            __syncwarp();
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(srow + ((g ^ sw) << 4)) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          fence_proxy_async();                    // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (m_tile * BM + q * 32 < m_valid) {   // warp-uniform
            if (elect_one()) {
              tma_store_3d(&p.tmOut, slab, col0, m_tile * BM + q * 32, 0);
              tma_store_commit();
            }
            __syncwarp();
          }
          if (EPI != EPI_STORE && e.colsum && !(DEBUG_HOOKS && (p.dbg & 8))) {
            // bias gradient: column sums of the staged (bf16-rounded) slab; lane -> column slab_colsum_col(lane).  The sums stay
            // in registers (one accumulator per (N tile, chunk) this warp can meet) until the end of the kernel: per-chunk
            // red.global.add to the same few cache lines from every CTA serialises in L2 (65536 warp-wide reds onto 32 lines for a
            // 1024-wide layer).  (columns >= nv of the slab are zero, rows >= m_valid too)
            // Lanes add up columns.  The alternative (PVAE_CS_MMA=1: ones . slab on mma.sync) stalls the tcgen05 stream that shares
            // the tensor pipe: measured 1024-deep dgrad 98 -> 118 us, 197-deep dgrad 54 -> 57 us.
            float sum = 0.f;
            if (cs_mma) {
              sum = slab_colsum32([&](int jp) { return slab + (uint32_t)lane * 64u + (uint32_t)((jp ^ sw) << 4); }, lane);
            } else {
              const int col = slab_colsum_col(lane);
              const int piece = col >> 3, within = (col & 7) << 1;
              float part[4] = {0.f, 0.f, 0.f, 0.f};     // four independent chains instead of one 32-long dependent one
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const unsigned short hv = *reinterpret_cast<const unsigned short*>(slab_gen + r * 64 + (((piece ^ ((r >> 1) & 3)) << 4) | within));
                part[r & 3] += __uint_as_float((uint32_t)hv << 16);
              }
              sum = (part[0] + part[1]) + (part[2] + part[3]);
            }
            if (n_tile < CS_TILES) {
#pragma unroll
              for (int t = 0; t < CS_TILES; ++t) cs_acc[2 * t + J] += (n_tile == t) ? sum : 0.f;
            } else if (slab_colsum_col(lane) < nv) {
              atomicAdd(e.colsum + col0 + slab_colsum_col(lane), sum);
            }
            __syncwarp();
          }
        };
        chunk(std::integral_constant<int, 0>{});
        chunk(std::integral_constant<int, 1>{});
      }
      if (ew == 0 && lane == 0) trace_stamp(p.dbg, tk, 6);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                            // the accumulator stage may be overwritten: tell the (leader's) MMA warp
        if (CG == 2) mbar_arrive_cluster(tempty0 + 8u * acc);
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
    if (TMAEPI) tma_store_wait_read<0>();
    if (TMAEPI && EPI != EPI_STORE && e.colsum) {
      // flush the bias-gradient sums: the four lane-quarter warps of a column group combine through their (now idle)
      // staging slabs, then one warp per column group issues the reds -- 32 warp-wide reds per CTA instead of 32 per tile.
      __syncwarp();
      float* mine = reinterpret_cast<float*>(slab_gen);
#pragma unroll
      for (int k = 0; k < 2 * CS_TILES; ++k) mine[k * 32 + lane] = cs_acc[k];
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      if ((ew & 3) == 0) {
#pragma unroll
        for (int k = 0; k < 2 * CS_TILES; ++k) {
          float t = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            t += reinterpret_cast<const float*>(smem_gen + OFF_STAGING + (cgrp * 4 + qq) * SLAB_BYTES)[k * 32 + lane];
          const int c = cgrp + 4 * (k & 1);
          const int cl = c * 32 + slab_colsum_col(lane);
          const int col = (k >> 1) * bn + cl;
          if (cl < bn && (k >> 1) < p.n_tiles && col < n_valid) atomicAdd(e.colsum + col, t);
        }
      }
    }
    if (EPI == EPI_MSE && e.loss) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) loss_local += __shfl_xor_sync(0xffffffffu, loss_local, off);
      if (lane == 0 && loss_local != 0.0) atomicAdd(e.loss, loss_local);
    }
    if (EPI == EPI_DGRAD && ACT == ACT_SWISH && e.act_grad) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) dbeta_local += __shfl_xor_sync(0xffffffffu, dbeta_local, off);
      if (lane == 0 && dbeta_local != 0.f) atomicAdd(e.act_grad, dbeta_local);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG > 1) cluster_sync_all();               // no CTA leaves while its peer may still write its TMEM / signal its barriers
  if (warp == 1) { tc_fence_after(); tmem_dealloc<CG>(tmem_base, TMEM_COLS); }
}

typedef void (*GemmKernelFn)(const GemmParams);
// host-side dispatch table: [epilogue type][activation][tma epilogue][CTAs per MMA]
template <int EPI, bool T, int CG> struct KernelRow {
  static GemmKernelFn get(int act) {
    switch (act) {
      case ACT_RELU:    return pvae_gemm_kernel<EPI, ACT_RELU, T, CG>;
      case ACT_TANH:    return pvae_gemm_kernel<EPI, ACT_TANH, T, CG>;
      case ACT_SIGMOID: return pvae_gemm_kernel<EPI, ACT_SIGMOID, T, CG>;
      case ACT_ELU:     return pvae_gemm_kernel<EPI, ACT_ELU, T, CG>;
      case ACT_SWISH:   return pvae_gemm_kernel<EPI, ACT_SWISH, T, CG>;
      default:          return pvae_gemm_kernel<EPI, ACT_LINEAR, T, CG>;
    }
  }
};
template <int CG> static inline GemmKernelFn select_kernel_cg(int epi, int act, bool tma) {
  switch (epi) {
    case EPI_STORE: return tma ? KernelRow<EPI_STORE, true, CG>::get(act) : KernelRow<EPI_STORE, false, CG>::get(act);
    case EPI_DGRAD: return tma ? KernelRow<EPI_DGRAD, true, CG>::get(act) : KernelRow<EPI_DGRAD, false, CG>::get(act);
    case EPI_MSE:   return tma ? pvae_gemm_kernel<EPI_MSE, ACT_LINEAR, true, CG> : pvae_gemm_kernel<EPI_MSE, ACT_LINEAR, false, CG>;
    default:        return pvae_gemm_kernel<EPI_WGRAD, ACT_LINEAR, false, CG>;
  }
}
static inline GemmKernelFn select_kernel(int epi, int act, bool tma, int cg, bool fast = false) {
  if (fast && tma && act == ACT_RELU && cg == 2) {       // lean epilogue (see the FAST block of the kernel)
    if (epi == EPI_STORE) return pvae_gemm_kernel<EPI_STORE, ACT_RELU, true, 2, true>;
    if (epi == EPI_DGRAD) return pvae_gemm_kernel<EPI_DGRAD, ACT_RELU, true, 2, true>;
  }
  return cg == 2 ? select_kernel_cg<2>(epi, act, tma) : select_kernel_cg<1>(epi, act, tma);
}

}  // namespace pvae
