// pvae_small.cuh -- latency path of the inference API: a whole FC stack (rllib_model_torch.FC.forward,
// rllib_model_torch.py:274-275) for a handful of rows in ONE kernel.
//
// The runtime consumer of the model calls it with batch 1 (envs/rllib_env_imitation.py:234-264: decoder pass-through per control
// step); a tcgen05 tile is 128 rows tall, so for <= 16 rows the tensor-core path is pure launch latency (15 launches, 82 us at
// batch 1 in round 1).  Here one thread-block CLUSTER of 8 CTAs runs the net layer by layer:
//   * every CTA keeps the full activation vector of every row in its shared memory;
//   * CTA r computes a slice of the layer's output neurons -- a warp per neuron, lanes stride the K dimension with 16-byte loads
//     of the bf16 shadow weights (1.3 MB for the decoder: L2-resident between control steps), fp32 accumulate, shuffle reduction --
//     into its own shared memory;
//   * one cluster barrier per layer, after which every CTA pulls the other seven slices out of its peers' shared memory
//     (distributed shared memory, 16-byte ld.shared::cluster).
// Weights are read exactly once per call, split eight ways; nothing but the final output touches global memory.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace pvae {

constexpr int SF_CLUSTER = 8, SF_THREADS = 256, SF_MAX_ROWS = 16, SF_SMALL_ROWS = 4, SF_MAX_LAYERS = 8;

struct SmallLayer {
  const __nv_bfloat16* W;      // shadow operand [plane][out][kpad]
  int64_t ps;                  // plane stride (elements)
  const float* bias;
  const float* act_param;      // swish: beta (device scalar) or null = 1
  int32_t out, kpad, act, pad;
};
struct SmallNet {
  SmallLayer L[SF_MAX_LAYERS];
  int32_t n_layers, planes;
  int32_t k0, k1, K0pad;       // layer-0 input segments in shadow-column order: [0, k0) <- in0, [K0pad, K0pad + k1) <- in1
  int32_t width;               // shared-memory row stride (floats): max over layers of kpad / out, multiple of 32
  int32_t trace;               // PVAE_SMALL_TRACE=1: CTA 0 prints its per-layer clock breakdown (debugging aid)
};

__device__ __forceinline__ float small_act(int act, float v, float beta = 1.f) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return tanhf(v);
    case 3: return 1.f / (1.f + __expf(-v));
    case 4: return v > 0.f ? v : expm1f(v);
    case 5: return v / (1.f + __expf(-beta * v));
    default: return v;
  }
}
__device__ __forceinline__ uint32_t small_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void small_st_cluster(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void small_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ float4 small_ld_cluster_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// BT: rows computed (a power of two >= the batch; missing rows are zero rows of shared memory)
//
// Exchange protocol (double buffer X / Y per CTA, one cluster barrier per layer): layer l reads the full input from X, writes its own
// slice of the output into its local Y; barrier; every CTA PULLS the seven remote slices out of the peers' Y with 16-byte
// ld.shared::cluster (scalar remote stores were 4x slower: distributed shared memory moves ~20 B/clk per CTA and pays per
// transaction); the roles of X and Y swap.  A peer may still be pulling from my Y while I already write my slice of layer l + 1 into X --
// different buffers -- and I cannot reach layer l + 2 (which writes Y again) before every peer has passed the barrier of layer l + 1,
// i.e. finished pulling.
template <int BT>
__global__ void __cluster_dims__(SF_CLUSTER, 1, 1) __launch_bounds__(SF_THREADS, 1)
small_fc_kernel(const __grid_constant__ SmallNet net, const float* __restrict__ in0, int64_t in0_ld, const float* __restrict__ in1, int64_t in1_ld, int B,
                float* __restrict__ out, int64_t out_ld) {
  extern __shared__ float sf_smem[];
  const int width = net.width;                   // multiple of 32
  float* buf[2] = {sf_smem, sf_smem + BT * width};
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // stage the input rows in shadow-column order (every CTA its own copy); in bf16 mode round like the tensor-core path does
  for (int i = threadIdx.x; i < BT * width; i += SF_THREADS) {
    const int b = i / width, c = i - b * width;
    float v = 0.f;
    if (b < B) {
      if (c < net.k0) v = in0[(int64_t)b * in0_ld + c];
      else if (c >= net.K0pad && c < net.K0pad + net.k1) v = in1[(int64_t)b * in1_ld + (c - net.K0pad)];
      if (net.planes == 1) v = __bfloat162float(__float2bfloat16_rn(v));
    }
    buf[0][i] = v;
    buf[1][i] = 0.f;
  }
  // PVAE_SMALL_TRACE=1: thread 0 of CTA 0 collects clock stamps (start, staged, per layer: computed / barrier passed / pulled) and
  // prints them once at the end (every __syncthreads below stays warp-uniform: the trace only adds stamps)
  const bool tr = net.trace && crank == 0 && threadIdx.x == 0;
  long long stamp[2 + 3 * SF_MAX_LAYERS];
  int ns = 0;
  if (tr) stamp[ns++] = clock64();
  small_cluster_sync();          // (also: every CTA of the cluster is running before anyone reads its shared memory)
  if (tr) stamp[ns++] = clock64();
  int cur = 0;
  for (int l = 0; l < net.n_layers; ++l) {
    const SmallLayer& L = net.L[l];
    const bool last = l == net.n_layers - 1;
    // slices of 32-neuron granularity so that a slice is whole 16-byte pieces for the pull
    const int per = (((L.out + SF_CLUSTER - 1) / SF_CLUSTER) + 31) & ~31;
    const int n_lo = min((int)crank * per, L.out), n_hi = min(n_lo + per, L.out);
    const float* x = buf[cur];
    float* y = buf[cur ^ 1];
    // A warp works on NPW neurons at a time: their weight loads and bias loads are in flight together, the shuffle reductions of
    // the NPW sums interleave, and lane j finishes neuron j (bias, activation, store) -- measured at batch 1, the loop is pure
    // latency (two warps per scheduler): processing neurons one after the other cost ~1000 clocks per neuron.
    constexpr int NPW = BT <= 2 ? 8 : 4, NW = SF_THREADS / 32;
    for (int n0 = n_lo + warp; n0 < n_hi; n0 += NW * NPW) {
      float acc[NPW][BT];
#pragma unroll
      for (int j = 0; j < NPW; ++j)
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[j][b] = 0.f;
      const int n_mine = n0 + (lane < NPW ? lane : 0) * NW;                 // the neuron this lane finishes
      const float bias = (L.bias && lane < NPW && n_mine < n_hi) ? __ldg(L.bias + n_mine) : 0.f;
      const float beta = (L.act == 5 && L.act_param) ? __ldg(L.act_param) : 1.f;
      for (int k = lane * 8; k < L.kpad; k += 256) {        // kpad is a multiple of 64: whole 16-byte pieces
        uint4 q[NPW], q2[NPW];
#pragma unroll
        for (int j = 0; j < NPW; ++j) {
          const int n = n0 + j * NW;
          const __nv_bfloat16* wrow = L.W + (int64_t)(n < n_hi ? n : n_lo) * L.kpad + k;
          q[j] = __ldg(reinterpret_cast<const uint4*>(wrow));
          if (net.planes > 1) q2[j] = __ldg(reinterpret_cast<const uint4*>(wrow + L.ps));
        }
        float xv[BT][8];
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const float4 x0 = *reinterpret_cast<const float4*>(x + b * width + k);
          const float4 x1 = *reinterpret_cast<const float4*>(x + b * width + k + 4);
          xv[b][0] = x0.x; xv[b][1] = x0.y; xv[b][2] = x0.z; xv[b][3] = x0.w; xv[b][4] = x1.x; xv[b][5] = x1.y; xv[b][6] = x1.z; xv[b][7] = x1.w;
        }
#pragma unroll
        for (int j = 0; j < NPW; ++j) {
          float w[8];
          const uint32_t qw[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) { w[2 * t] = __uint_as_float(qw[t] << 16); w[2 * t + 1] = __uint_as_float(qw[t] & 0xFFFF0000u); }
          if (net.planes > 1) {
            const uint32_t q2w[4] = {q2[j].x, q2[j].y, q2[j].z, q2[j].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) { w[2 * t] += __uint_as_float(q2w[t] << 16); w[2 * t + 1] += __uint_as_float(q2w[t] & 0xFFFF0000u); }
          }
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            float a0 = acc[j][b], a1 = 0.f;               // two chains per (neuron, row)
#pragma unroll
            for (int t = 0; t < 4; ++t) { a0 = fmaf(w[t], xv[b][t], a0); a1 = fmaf(w[4 + t], xv[b][4 + t], a1); }
            acc[j][b] = a0 + a1;
          }
        }
      }
      // butterfly reductions of all NPW x BT sums (independent chains), then lane j keeps neuron j
      float mine[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b) mine[b] = 0.f;
#pragma unroll
      for (int j = 0; j < NPW; ++j) {
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          float a = acc[j][b];
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
          mine[b] = (lane == j) ? a : mine[b];
        }
      }
      if (lane < NPW && n_mine < n_hi) {
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          float v = small_act(L.act, mine[b] + bias, beta);
          if (last) {
            if (b < B) out[(int64_t)b * out_ld + n_mine] = v;
          } else {
            if (net.planes == 1) v = __bfloat162float(__float2bfloat16_rn(v));
            y[b * width + n_mine] = v;
          }
        }
      }
    }
    if (net.trace) __syncthreads();                  // (uniform: net.trace is a kernel argument)
    if (tr) stamp[ns++] = clock64();
    if (!last) {
      small_cluster_sync();                          // every CTA's slice of layer l is in its local y
      if (tr) stamp[ns++] = clock64();
      // pull the remote slices (16-byte pieces) and clear the tail [out, next kpad) that the next layer reads as K padding
      const int pieces = per >> 2;                   // float4 pieces per (slice, row)
      const uint32_t y_s = (uint32_t)__cvta_generic_to_shared(y);
      for (int i = threadIdx.x; i < (SF_CLUSTER - 1) * BT * pieces; i += SF_THREADS) {
        const int pc = i % pieces;
        const int b = (i / pieces) % BT;
        const int pr = (int)(crank + 1 + i / (pieces * BT)) % SF_CLUSTER;      // start with the next rank: spread the traffic
        const int c0 = pr * per + pc * 4;
        if (c0 < L.out) {
          const uint32_t off = (uint32_t)(b * width + c0) * 4u;
          const float4 v = small_ld_cluster_v4(small_mapa(y_s + off, (uint32_t)pr));
          *reinterpret_cast<float4*>(y + b * width + c0) = v;
        }
      }
      __syncthreads();
      const int next_k = net.L[l + 1].kpad;
      for (int i = threadIdx.x; i < BT * (next_k - L.out); i += SF_THREADS) {
        const int b = i / (next_k - L.out), c = L.out + i % (next_k - L.out);
        y[b * width + c] = 0.f;
      }
      __syncthreads();
      if (tr) stamp[ns++] = clock64();
      cur ^= 1;
    }
  }
  small_cluster_sync();          // no CTA exits while a peer may still read its shared memory
  if (tr) {
    stamp[ns++] = clock64();
    printf("[small_fc] clocks since start:");
    for (int i = 1; i < ns; ++i) printf(" %lld", stamp[i] - stamp[0]);
    printf("  (staged | per layer: computed, barrier, pulled | ... | last computed, exit barrier)\n");
  }
}

}  // namespace pvae
