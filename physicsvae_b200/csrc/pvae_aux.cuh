// pvae_aux.cuh -- the small HBM-bound kernels around the tensor-core GEMM:
//   ingest          raw transition arrays (f64/f32) -> resident bf16 hi(/lo) planes, vectorised + coalesced
//   sync_weights    fp32 nn.Linear masters -> K-padded bf16 hi(/lo) shadow operands
//   reparam_fwd     z = mu + eps * exp(0.5 logvar), KL partial sums          (rllib_model_torch.py:734-740,
//                                                                              train_physics_vae.py:384-389)
//   reparam_bwd     d(mu|logvar) from dz and the KL term, bias-grad column sums
//   finalize_loss   the weighted sum of train_physics_vae.py:430-434
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace pvae {

__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }

// ---- ingest: one thread per (row, 2 columns) pair; rows are contiguous so warps read/write whole lines ------
// source column of destination column c when the first w0 source columns stay at [0, w0) and the remaining ones start at
// the 16-byte aligned destination column p0 (the (s_t | s_{t+1}) split of a transition row); -1 = padding
__device__ __forceinline__ int split_src_col(int c, int w0, int p0, int width) {
  if (c < p0) return c < w0 ? c : -1;
  const int j = w0 + (c - p0);
  return j < width ? j : -1;
}
// src2 (optional): a second fp32 array whose w2 columns land at destination columns [c2, c2 + w2) (a_t behind s_t)
template <typename SrcT>
__global__ void ingest_kernel(const SrcT* __restrict__ src, int64_t src_ld, int width, int w0, int p0,
                              const float* __restrict__ src2, int64_t src2_ld, int w2, int c2, __nv_bfloat16* __restrict__ dst,
                              int64_t dst_ld, int64_t dst_ps, int planes, int64_t n_rows) {
  const int64_t pairs_per_row = dst_ld >> 1;   // dst_ld is a multiple of 8
  const int64_t total = n_rows * pairs_per_row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / pairs_per_row;
    const int c = (int)(i - r * pairs_per_row) * 2;
    const int s0 = split_src_col(c, w0, p0, width), s1 = split_src_col(c + 1, w0, p0, width);
    float v0 = (s0 >= 0) ? (float)src[r * src_ld + s0] : 0.f;
    float v1 = (s1 >= 0) ? (float)src[r * src_ld + s1] : 0.f;
    if (src2) {
      if (c >= c2 && c < c2 + w2) v0 = src2[r * src2_ld + (c - c2)];
      if (c + 1 >= c2 && c + 1 < c2 + w2) v1 = src2[r * src2_ld + (c + 1 - c2)];
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    *reinterpret_cast<__nv_bfloat162*>(dst + r * dst_ld + c) = h;
    if (planes > 1) {
      __nv_bfloat162 l = __floats2bfloat162_rn(v0 - __low2float(h), v1 - __high2float(h));
      *reinterpret_cast<__nv_bfloat162*>(dst + dst_ps + r * dst_ld + c) = l;
    }
  }
}

// ---- dataset build on the device (load_dataset_for_PhysicsVAE, train_physics_vae.py:133-156): every state of every episode is
// uploaded ONCE ([n_states][dsb]); transition r is (s = states[first[r]], a = actions[first[r]], s' = states[first[r] + 1]).
// One thread per (row, 2 destination columns) of the resident x rows (s | 0.. | a | 0.. | s' | 0..) / y rows (a | 0..).
template <typename T> __device__ __forceinline__ float to_f32(T v) { return (float)v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename SrcT, typename ActT = float>
__global__ void ingest_episodes_kernel(const SrcT* __restrict__ states, const ActT* __restrict__ actions,
                                       const int64_t* __restrict__ first, int dsb, int da, int a_col, int s2_col,
                                       __nv_bfloat16* __restrict__ dst, int64_t dst_ld, int64_t dst_ps, int planes, int64_t n_rows) {
  const int64_t pairs_per_row = dst_ld >> 1;
  const int64_t total = n_rows * pairs_per_row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / pairs_per_row;
    const int c0 = (int)(i - r * pairs_per_row) * 2;
    const int64_t st = first[r];
    float v[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = c0 + k;
      float x = 0.f;
      if (s2_col >= 0) {                       // x row
        if (c < dsb) x = to_f32(states[st * dsb + c]);
        else if (c >= a_col && c < a_col + da) x = to_f32(actions[st * da + (c - a_col)]);
        else if (c >= s2_col && c < s2_col + dsb) x = to_f32(states[(st + 1) * dsb + (c - s2_col)]);
      } else if (c < da) {                     // y row
        x = to_f32(actions[st * da + c]);
      }
      v[k] = x;
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
    *reinterpret_cast<__nv_bfloat162*>(dst + r * dst_ld + c0) = h;
    if (planes > 1) {
      __nv_bfloat162 l = __floats2bfloat162_rn(v[0] - __low2float(h), v[1] - __high2float(h));
      *reinterpret_cast<__nv_bfloat162*>(dst + dst_ps + r * dst_ld + c0) = l;
    }
  }
}

// ---- shadow weights: Wsh[plane][out][Kpad]; columns [0,K0pad) <- W[:, 0:k0], [K0pad, K0pad+K1pad) <- W[:, k0:k0+k1]
__global__ void sync_weights_kernel(const float* __restrict__ W, int out, int in, int k0, int k1, int K0pad, int Kpad,
                                    __nv_bfloat16* __restrict__ Wsh, int64_t ps, int planes) {
  const int64_t total = (int64_t)out * Kpad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / Kpad);
    const int c = (int)(i - (int64_t)o * Kpad);
    int src_c = -1;
    if (c < K0pad) { if (c < k0) src_c = c; }
    else { const int j = c - K0pad; if (j < k1) src_c = k0 + j; }
    const float v = (src_c >= 0) ? W[(int64_t)o * in + src_c] : 0.f;
    const __nv_bfloat16 h = f2bf(v);
    Wsh[i] = h;
    if (planes > 1) Wsh[ps + i] = f2bf(v - __bfloat162float(h));
  }
}

// ---- fused Adam + shadow refresh for one Linear layer ----------------------------------------------------------------------
// torch.optim.Adam semantics (torch_models.py:119-122: lr, betas (0.9, 0.999), eps 1e-8, weight_decay L2, no amsgrad):
//   g += wd * p;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// and, in the same pass, the bf16 (hi / lo) shadow operand of the updated weight in its padded two-segment layout.
// Elements [0, out*in) are the weight, [out*in, out*in + out) the bias (flat layout of one layer: W | b).
__global__ void adam_layer_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                                  float* __restrict__ v, const float* __restrict__ step_dev, float lr, float b1, float b2,
                                  float eps, float wd, int out, int in, int k0, int K0pad, int Kpad,
                                  __nv_bfloat16* __restrict__ Wsh, int64_t ps, int planes) {
  const float t = *step_dev;
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  const int64_t nw = (int64_t)out * in, total = nw + out;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float p = param[i];
    float g = grad[i];
    if (wd != 0.f) g += wd * p;
    const float mi = b1 * m[i] + (1.f - b1) * g;
    const float vi = b2 * v[i] + (1.f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    p -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    param[i] = p;
    if (i < nw) {
      const int o = (int)(i / in);
      const int c = (int)(i - (int64_t)o * in);
      const int sc = c < k0 ? c : K0pad + (c - k0);
      const int64_t si = (int64_t)o * Kpad + sc;
      const __nv_bfloat16 h = f2bf(p);
      Wsh[si] = h;
      if (planes > 1) Wsh[ps + si] = f2bf(p - __bfloat162float(h));
    }
  }
}
__global__ void add_scalar_kernel(float* x, float d) { if (threadIdx.x == 0 && blockIdx.x == 0) *x += d; }

// ---- fused Adam + shadow refresh for a whole net: ONE launch over the flat [W0 | b0 | W1 | b1 | ...] buffers ------------------
// (three per-layer launches + the step-counter kernel cost 26 us per world step, this one ~8: the work is 1.5 M elements).
// The step counter lives on the device (a captured graph replays the launch): every block reads it on entry, the block that
// finishes last writes t + 1 back -- by then every other block has finished, hence read it.
struct AdamLayer {
  int64_t off;                 // first flat element of the layer (its weight; the bias follows at off + out * in)
  int32_t out, in, k0, K0pad, Kpad, active;
  __nv_bfloat16* Wsh;
  int64_t ps;
};
struct AdamNet {
  AdamLayer L[8];
  int32_t n_layers;
  int64_t total;
};
__device__ __forceinline__ void adam_element(const AdamNet& net, int64_t i, float& p, float g, float& mi, float& vi, float b1, float b2,
                                             float eps, float wd, float step_size, float bc2_sqrt, int planes, bool& active) {
  int l = 0;
#pragma unroll 1
  while (l + 1 < net.n_layers && i >= net.L[l + 1].off) ++l;
  const AdamLayer& L = net.L[l];
  active = L.active != 0;
  if (!active) return;
  if (wd != 0.f) g += wd * p;
  mi = b1 * mi + (1.f - b1) * g;
  vi = b2 * vi + (1.f - b2) * g * g;
  p -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  const int64_t local = i - L.off;
  if (local < (int64_t)L.out * L.in) {
    const int o = (int)(local / L.in);
    const int c = (int)(local - (int64_t)o * L.in);
    const int sc = c < L.k0 ? c : L.K0pad + (c - L.k0);
    const int64_t si = (int64_t)o * L.Kpad + sc;
    const __nv_bfloat16 h = f2bf(p);
    L.Wsh[si] = h;
    if (planes > 1) L.Wsh[L.ps + si] = f2bf(p - __bfloat162float(h));
  }
}
__global__ void adam_net_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                float* __restrict__ step_dev, unsigned int* __restrict__ done_blocks, float lr, float b1, float b2,
                                float eps, float wd, const __grid_constant__ AdamNet net, int planes) {
  const float t = *step_dev + 1.f;
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  const int64_t groups = (net.total + 3) >> 2;
  for (int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gi < groups; gi += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i0 = gi << 2;
    if (i0 + 3 < net.total) {
      float4 p4 = *reinterpret_cast<float4*>(param + i0);
      const float4 g4 = *reinterpret_cast<const float4*>(grad + i0);
      float4 m4 = *reinterpret_cast<float4*>(m + i0);
      float4 v4 = *reinterpret_cast<float4*>(v + i0);
      bool a0, a1, a2, a3;
      adam_element(net, i0 + 0, p4.x, g4.x, m4.x, v4.x, b1, b2, eps, wd, step_size, bc2_sqrt, planes, a0);
      adam_element(net, i0 + 1, p4.y, g4.y, m4.y, v4.y, b1, b2, eps, wd, step_size, bc2_sqrt, planes, a1);
      adam_element(net, i0 + 2, p4.z, g4.z, m4.z, v4.z, b1, b2, eps, wd, step_size, bc2_sqrt, planes, a2);
      adam_element(net, i0 + 3, p4.w, g4.w, m4.w, v4.w, b1, b2, eps, wd, step_size, bc2_sqrt, planes, a3);
      if (a0 | a1 | a2 | a3) {           // (inactive elements come back unchanged)
        *reinterpret_cast<float4*>(param + i0) = p4;
        *reinterpret_cast<float4*>(m + i0) = m4;
        *reinterpret_cast<float4*>(v + i0) = v4;
      }
    } else {
      for (int64_t i = i0; i < net.total; ++i) {
        float p = param[i], mi = m[i], vi = v[i];
        bool a;
        adam_element(net, i, p, grad[i], mi, vi, b1, b2, eps, wd, step_size, bc2_sqrt, planes, a);
        if (a) { param[i] = p; m[i] = mi; v[i] = vi; }
      }
    }
  }
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(done_blocks, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *step_dev = t;
    *done_blocks = 0u;
  }
}

// ---- Philox4x32-10 + Box-Muller: counter-based N(0,1) stream for the (seed, offset) noise mode ----------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t offset, uint64_t idx) {
  uint32_t c[4] = {(uint32_t)(idx >> 1), (uint32_t)(idx >> 33), (uint32_t)offset, (uint32_t)(offset >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u1 = ((float)c[0] + 0.5f) * 2.3283064365386963e-10f;   // (0,1)
  const float u2 = ((float)c[1] + 0.5f) * 2.3283064365386963e-10f;
  const float rad = sqrtf(-2.f * __logf(u1));
  float s, co;
  __sincosf(6.283185307179586f * u2, &s, &co);
  return (idx & 1) ? rad * s : rad * co;
}

// ---- reparameterise + KL.  ml: [B][w] fp32, w = 2z (mu | logvar) with a prior, z without.  One thread per (row, latent)
__global__ void reparam_fwd_kernel(const float* __restrict__ ml, const float* __restrict__ eps_in, float* __restrict__ eps_out,
                                   int has_lv, int noise, uint64_t seed, uint64_t offset, int B, int z,
                                   __nv_bfloat16* __restrict__ zb, int64_t z_ld, int64_t z_ps, int planes,
                                   float* __restrict__ z_f32, float* __restrict__ mu_out, float* __restrict__ lv_out,
                                   double* __restrict__ kl_acc, const unsigned long long* __restrict__ noise_ctr) {
  const int64_t total = (int64_t)B * z;
  if (noise_ctr) offset += *noise_ctr;           // device-side step counter of the Philox stream (graph replays draw fresh noise)
  const int w = has_lv ? 2 * z : z;
  float kl = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / z), d = (int)(i - (int64_t)r * z);
    const float mu = ml[(int64_t)r * w + d];
    const float lv = has_lv ? ml[(int64_t)r * w + z + d] : 0.f;
    float zt = mu;
    if (noise) {
      const float e = eps_in ? eps_in[i] : philox_normal(seed, offset, (uint64_t)i);
      if (eps_out) eps_out[i] = e;
      zt = mu + e * expf(0.5f * lv);
    }
    const __nv_bfloat16 h = f2bf(zt);
    zb[(int64_t)r * z_ld + d] = h;
    if (planes > 1) zb[z_ps + (int64_t)r * z_ld + d] = f2bf(zt - __bfloat162float(h));
    if (z_f32) z_f32[i] = zt;
    if (mu_out) mu_out[i] = mu;
    if (lv_out) lv_out[i] = lv;
    kl += -0.5f * (1.f + lv - mu * mu - expf(lv));
  }
  if (kl_acc) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, off);
    __shared__ float part[32];
    const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) part[wi] = kl;
    __syncthreads();
    if (wi == 0) {
      float t = (l < (int)(blockDim.x >> 5)) ? part[l] : 0.f;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
      if (l == 0) atomicAdd(kl_acc, (double)t);
    }
  }
}

// ---- backward of reparameterise + KL.  dml[B][w] = (dmu | dlogvar);  colsum -> bias grad of the encoder's last layer
//   dmu = dz + kl_scale*mu ; dlogvar = dz*eps*0.5*exp(0.5 lv) + kl_scale*0.5*(exp(lv)-1),  kl_scale = kl_coeff / B
// without a prior (has_lv == 0) the encoder output IS z: dml = dz.
// blockDim.x == w * rows_per_block, thread t -> column t % w
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ ml, const float* __restrict__ eps,
                                   int has_lv, int noise, float kl_scale, int B, int z, __nv_bfloat16* __restrict__ dml,
                                   int64_t ld, int64_t ps, int planes, float* __restrict__ colsum) {
  const int w2 = has_lv ? 2 * z : z;
  const int col = threadIdx.x % w2;
  const int rsub = threadIdx.x / w2;
  const int rows_per_block = blockDim.x / w2;
  float acc = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + rsub; r < B; r += (int64_t)gridDim.x * rows_per_block) {
    const int d = (col < z) ? col : col - z;
    const float g = dz[r * z + d];
    float v;
    if (!has_lv) {
      v = g;
    } else {
      const float mu = ml[r * w2 + d];
      const float lv = ml[r * w2 + z + d];
      if (col < z) v = g + kl_scale * mu;
      else v = (noise ? g * eps[r * z + d] * 0.5f * expf(0.5f * lv) : 0.f) + kl_scale * 0.5f * (expf(lv) - 1.f);
    }
    const __nv_bfloat16 h = f2bf(v);
    dml[r * ld + col] = h;
    if (planes > 1) dml[ps + r * ld + col] = f2bf(v - __bfloat162float(h));
    acc += v;
  }
  if (colsum) atomicAdd(colsum + col, acc);
}

// ---- fp32 [B][w] -> bf16 planes (decoder / world inputs handed in by the inference API) -------------------------
__global__ void f32_to_planes_kernel(const float* __restrict__ src, int64_t src_ld, int width, int w0, int p0,
                                     __nv_bfloat16* __restrict__ dst, int64_t dst_ld, int64_t dst_ps, int planes, int64_t n_rows) {
  const int64_t total = n_rows * dst_ld;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dst_ld;
    const int c = (int)(i - r * dst_ld);
    const int sc = split_src_col(c, w0, p0, width);
    const float v = (sc >= 0) ? src[r * src_ld + sc] : 0.f;
    const __nv_bfloat16 h = f2bf(v);
    dst[i] = h;
    if (planes > 1) dst[dst_ps + i] = f2bf(v - __bfloat162float(h));
  }
}

// ---- deterministic bias gradients (pvae_set_deterministic): column sums of a bf16 (hi + lo) gradient tensor in a FIXED order ------------
// pass 1: block (column group of 32, row chunk c of DET_CHUNKS) -- thread (column, row lane of 8) walks its rows in order, the 8 row
// lanes are combined in order -> partial[c][column];  pass 2: the DET_CHUNKS partials are added in order and accumulated into the
// bias gradient.  No atomics: the result does not depend on scheduling.
constexpr int DET_CHUNKS = 64;
__global__ void colsum_det_partial_kernel(const __nv_bfloat16* __restrict__ g, int64_t ld, int64_t ps, int planes, int rows, int cols,
                                          float* __restrict__ partial) {
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;                       // 0..7
  const int per = (rows + DET_CHUNKS - 1) / DET_CHUNKS;
  const int r0 = blockIdx.y * per, r1 = min(r0 + per, rows);
  float acc = 0.f;
  if (col < cols)
    for (int r = r0 + rl; r < r1; r += 8) {
      float v = __bfloat162float(g[(int64_t)r * ld + col]);
      if (planes > 1) v += __bfloat162float(g[ps + (int64_t)r * ld + col]);
      acc += v;
    }
  __shared__ float sh[8][32];
  sh[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x];
    partial[(int64_t)blockIdx.y * cols + col] = t;
  }
}
__global__ void colsum_det_final_kernel(const float* __restrict__ partial, int cols, float* __restrict__ colsum) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= cols) return;
  float t = 0.f;
  for (int c = 0; c < DET_CHUNKS; ++c) t += partial[(int64_t)c * cols + col];
  colsum[col] += t;
}

// ---- data-parallel gradient exchange over peer memory (NVLink 5 / NVSwitch) ------------------------------------------------------
// One kernel = barrier + reduce-scatter + all-gather of a contiguous fp32 range that lives at the same offset of a symmetric
// (peer-mapped) allocation on every rank -- the [gradients | loss slots] range of the model's gradient pool:
//   1. every CTA b announces itself to CTA b of every peer (flag in the PEER's copy of the pool) and waits for theirs: once a
//      peer's kernel is running, that peer's weight-gradient kernels have completed (stream order);
//   2. rank r owns slice r of the range: it loads the slice from every rank (plain loads on the peer pointers, or ONE
//      multimem.ld_reduce on the multicast address when the pool is multicast-bound: the switch adds), scales by 1 / world and
//      stores the result into every rank's copy in place (peer stores, or ONE multimem.st) -- a slice is read and written by its
//      owner only, so in-place is race-free, and because one rank computes each element all replicas end up bit-identical;
//   3. the same barrier again: every rank's copy is complete before Adam reads it, and nobody's next step clears its
//      gradients while a peer still reads them.
// Sequence numbers live in the pool (per CTA, bumped by the kernel itself), so a captured CUDA graph replays it.
constexpr int AR_CTAS = 64, AR_THREADS = 512, AR_MAX_RANKS = 8;
struct SymmArgs {
  float* peer[AR_MAX_RANKS];      // base of the symmetric allocation on every rank (peer-mapped addresses)
  float* mc;                      // multicast address of the same allocation, or null
  int32_t rank, world;
  int64_t off, count;             // the range [off, off + count) in fp32 elements; off and count are multiples of 4
  int64_t flags_off;              // where the flag block starts (fp32 elements): [AR_CTAS][AR_MAX_RANKS] arrival flags, then [AR_CTAS] sequence counters
  float scale;
};
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void symm_barrier(const SymmArgs& a, uint32_t* my_flags, uint32_t seq) {
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    uint32_t* theirs = reinterpret_cast<uint32_t*>(a.peer[threadIdx.x] + a.flags_off) + blockIdx.x * AR_MAX_RANKS + a.rank;
    st_release_sys(theirs, seq);
    const uint32_t* mine = my_flags + blockIdx.x * AR_MAX_RANKS + threadIdx.x;
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {
      if ((++spins & 0x3FFFu) == 0) {                 // bounded: a rank that never arrives must not hang the GPU
        uint64_t t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0) t0 = t;
        else if (t - t0 > 20000000000ull) {            // 20 s
          printf("pvae symm_allreduce: rank %d CTA %d still waits for rank %d (seq %u)\n", a.rank, (int)blockIdx.x, (int)threadIdx.x, seq);
          __trap();
        }
      }
    }
  }
  __syncthreads();
}
__global__ void __launch_bounds__(AR_THREADS) symm_allreduce_kernel(const SymmArgs a) {
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(a.peer[a.rank] + a.flags_off);
  uint32_t* ctr = my_flags + AR_CTAS * AR_MAX_RANKS + blockIdx.x;
  const uint32_t seq = *ctr;
  symm_barrier(a, my_flags, seq + 1);
  const int64_t n4 = a.count >> 2;
  const int64_t per = (n4 + a.world - 1) / a.world;
  const int64_t lo = (int64_t)a.rank * per, hi = lo + per < n4 ? lo + per : n4;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = a.off + (i << 2);
    float4 s;
    if (a.mc) {
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(s.x), "=f"(s.y), "=f"(s.z), "=f"(s.w) : "l"(a.mc + e) : "memory");
    } else {
      s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < AR_MAX_RANKS; ++p) {
        if (p < a.world) {                              // fixed order 0 .. world-1: the sum does not depend on who computes it
          float4 v;
          asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a.peer[p] + e) : "memory");
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
      }
    }
    s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
    if (a.mc) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.mc + e), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
    } else {
#pragma unroll
      for (int p = 0; p < AR_MAX_RANKS; ++p)
        if (p < a.world)
          asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.peer[p] + e), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
    }
  }
  __threadfence_system();
  symm_barrier(a, my_flags, seq + 2);
  if (threadIdx.x == 0) *ctr = seq + 2;
}

// The same exchange with the data moved by the bulk-copy engine (cp.async.bulk) instead of per-thread loads: one thread of a
// CTA keeps ARB_STAGES x 32 KiB of peer reads in flight (a stage = the same chunk of the slice from every rank, landing in
// shared memory, completion on an mbarrier), the 256 threads add the `world` copies in fixed rank order, and the averaged chunk
// leaves as one bulk store per rank.  Bytes in flight per SM no longer depend on registers: 160 KiB instead of the 64 KiB that
// 512 threads x 8 outstanding 16-byte loads hold -- what an exchange squeezed onto a few SMs beside the backward GEMMs needs.
// Barriers, flag block, slice ownership and summation order are those of symm_allreduce_kernel (results bit-identical to it).
constexpr int ARB_THREADS = 256, ARB_STAGES = 5, ARB_STAGE_BYTES = 32768, ARB_OUT_SLOTS = 3, ARB_OUT_BYTES = 16384;
constexpr int ARB_SMEM_BYTES = ARB_STAGES * ARB_STAGE_BYTES + ARB_OUT_SLOTS * ARB_OUT_BYTES + 64;
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(dst)), "r"(src_smem), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(ARB_THREADS) symm_allreduce_bulk_kernel(const SymmArgs a) {
  extern __shared__ __align__(128) uint8_t arb_smem[];
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(a.peer[a.rank] + a.flags_off);
  uint32_t* ctr = my_flags + AR_CTAS * AR_MAX_RANKS + blockIdx.x;
  const uint32_t seq = *ctr;
  const uint32_t smem0 = smem_u32(arb_smem);
  const uint32_t out0 = smem0 + ARB_STAGES * ARB_STAGE_BYTES;
  const uint32_t bar0 = out0 + ARB_OUT_SLOTS * ARB_OUT_BYTES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ARB_STAGES; ++s) mbar_init(bar0 + 8u * s, 1);
    fence_barrier_init();
  }
  symm_barrier(a, my_flags, seq + 1);            // (its __syncthreads also publish the barrier initialisation)
  asm volatile("fence.proxy.async;" ::: "memory");   // peer data observed through the flags (generic proxy) -> bulk reads (async proxy)
  const int64_t n4 = a.count >> 2;
  const int64_t per = (n4 + a.world - 1) / a.world;
  const int64_t lo = (int64_t)a.rank * per, hi = lo + per < n4 ? lo + per : n4;
  // a chunk = chunk4 float4 of the slice, from every rank: world * chunk4 * 16 B <= one stage; the sum of a chunk <= one out slot
  int chunk4 = (ARB_STAGE_BYTES / 16) / a.world;
  if (chunk4 > ARB_OUT_BYTES / 16) chunk4 = ARB_OUT_BYTES / 16;
  chunk4 &= ~63;
  const int64_t n_chunks = hi > lo ? (hi - lo + chunk4 - 1) / chunk4 : 0;
  const int64_t my_chunks = n_chunks > (int64_t)blockIdx.x ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](int64_t k) {                  // thread 0: the k-th chunk of this CTA into stage k % ARB_STAGES
    const int64_t c = (int64_t)blockIdx.x + k * gridDim.x;
    const int64_t first = lo + c * chunk4;
    const uint32_t n = (uint32_t)((hi - first < chunk4 ? hi - first : chunk4) * 16);
    const int st = (int)(k % ARB_STAGES);
    const uint32_t bar = bar0 + 8u * st;
    mbar_expect_tx(bar, n * (uint32_t)a.world);
    for (int p = 0; p < a.world; ++p)
      bulk_load(smem0 + st * ARB_STAGE_BYTES + p * (chunk4 * 16), a.peer[p] + a.off + (first << 2), n, bar);
  };
  if (threadIdx.x == 0)
    for (int64_t k = 0; k < my_chunks && k < ARB_STAGES; ++k) issue(k);
  for (int64_t k = 0; k < my_chunks; ++k) {
    const int64_t c = (int64_t)blockIdx.x + k * gridDim.x;
    const int64_t first = lo + c * chunk4;
    const int n4c = (int)(hi - first < chunk4 ? hi - first : chunk4);
    const int st = (int)(k % ARB_STAGES);
    const int os = (int)(k % ARB_OUT_SLOTS);
    mbar_wait<W_FULL>(bar0 + 8u * st, (uint32_t)((k / ARB_STAGES) & 1));
    const float4* in = reinterpret_cast<const float4*>(arb_smem + st * ARB_STAGE_BYTES);
    float4* out = reinterpret_cast<float4*>(arb_smem + ARB_STAGES * ARB_STAGE_BYTES + os * ARB_OUT_BYTES);
    for (int i = threadIdx.x; i < n4c; i += ARB_THREADS) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < a.world; ++p) {          // fixed order 0 .. world-1, as in symm_allreduce_kernel
        const float4 v = in[p * chunk4 + i];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      t.x *= a.scale; t.y *= a.scale; t.z *= a.scale; t.w *= a.scale;
      out[i] = t;
    }
    fence_proxy_async();                           // this thread's shared-memory writes -> visible to the bulk stores
    __syncthreads();                               // the out slot is complete, the input stage has been read by everyone
    if (threadIdx.x == 0) {
      for (int p = 0; p < a.world; ++p)
        bulk_store(a.peer[p] + a.off + (first << 2), out0 + os * ARB_OUT_BYTES, (uint32_t)n4c * 16u);
      tma_store_commit();
      // three out slots, one __syncthreads per chunk: the slot that chunk k + 2 fills was last read by the stores of chunk k - 1,
      // which this wait (only the newest group may be pending) retires before thread 0 reaches the __syncthreads of chunk k + 1
      tma_store_wait_read<1>();
      if (k + ARB_STAGES < my_chunks) issue(k + ARB_STAGES);
    }
  }
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // the stores have been performed, not just read out of shared memory
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __threadfence_system();
  symm_barrier(a, my_flags, seq + 2);
  if (threadIdx.x == 0) *ctr = seq + 2;
}

// ---- loss bookkeeping ---------------------------------------------------------------------------------------------
// acc: [0] sum sq a, [1] sum kl, [2] sum sq s (world), [3] sum sq cyc
// (the accumulators are cleared here, for the next step: one memset node less per step)
__global__ void finalize_loss_kernel(double* __restrict__ acc, float* __restrict__ loss, int B, int da, int dsb,
                                     float a_c, float kl_c, float s_c, float cyc_c, unsigned long long* __restrict__ noise_ctr,
                                     unsigned long long noise_stride, int steps = 1) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (noise_ctr) *noise_ctr += noise_stride;
    // `steps` > 1: the accumulators hold the sums over the steps of an autoregressive rollout, the loss terms are their means
    // over the steps (train_physics_vae.py:423-428)
    const double inv = 1.0 / (double)steps;
    const float la = (float)(acc[0] * inv / ((double)B * da));
    const float lk = (float)(acc[1] * inv / (double)B);
    const float ls = (float)(acc[2] * inv / ((double)B * dsb));
    const float lc = (float)(acc[3] * inv / ((double)B * dsb));
    loss[1] = la; loss[2] = lk; loss[3] = ls; loss[4] = lc;
    loss[0] = a_c * la + kl_c * lk + s_c * ls + cyc_c * lc;
    acc[0] = 0.0; acc[1] = 0.0; acc[2] = 0.0; acc[3] = 0.0;
  }
}

__global__ void set_cursor_kernel(int32_t* cur, int32_t v) { if (threadIdx.x == 0) *cur = v; }
__global__ void set_u64_kernel(unsigned long long* p, unsigned long long v) { if (threadIdx.x == 0) *p = v; }
// cursor += delta; past the end it wraps to (cursor mod delta): a rank that started at row s < delta (its slice of the first global
// mini-batch) is back on row s, a single rank on row 0
__global__ void advance_cursor_kernel(int32_t* cur, int32_t delta, int32_t batch, int32_t limit) {
  if (threadIdx.x == 0) {
    int32_t c = *cur + delta;
    if (c + batch > limit) c = delta > 0 ? c % delta : 0;
    *cur = c;
  }
}

}  // namespace pvae
