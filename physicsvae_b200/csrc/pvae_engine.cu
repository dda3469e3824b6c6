// pvae_engine.cu -- host side of libpvae_sm100.so: the C ABI declared in include/pvae_sm100.h.
//
// The engine owns no tensors besides bf16 shadow weights and a few scalars.  Every step entry point turns the
// reference's Python control flow (TrainModel.compute_loss, train_physics_vae.py:361-435; PhysicsVAE.forward*,
// rllib_model_torch.py:742-853; autograd of both, torch_models.py:142) into a fixed sequence of launches of the single
// tcgen05 GEMM kernel in pvae_gemm.cuh with different operand views / epilogues, plus the few elementwise kernels of
// pvae_aux.cuh.  Nothing here synchronises or allocates, so a step can be captured into a CUDA graph.
#include "../../include/pvae_sm100.h"
#include "pvae_gemm.cuh"
#include "pvae_aux.cuh"
#include "pvae_small.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace pvae {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(PVAE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKR(call)                  \
  do {                             \
    int r_ = (call);               \
    if (r_ != PVAE_OK) return r_;  \
  } while (0)

static inline int rup(int v, int m) { return (v + m - 1) / m * m; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- driver entry point for tensor-map encoding (resolved at run time: the library must load on a box without libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int resolve_driver() {
  if (g_encode) return PVAE_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
    return fail(PVAE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver (%s)", cudaGetErrorString(e));
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return PVAE_OK;
}

// A 2-D bf16 matrix (one or two precision planes) somewhere in device memory.
struct View {
  const __nv_bfloat16* base = nullptr;
  int64_t ld = 0;       // row stride, elements (multiple of 8)
  int64_t ps = 0;       // plane stride, elements
  int planes = 1;
  int width = 0;        // valid columns
  int64_t rows = 0;     // valid rows
  int dyn = 0;          // add the device-side row cursor to row coordinates
};

static int encode_map(CUtensorMap* m, const View& v, int box_cols, int box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  if ((reinterpret_cast<uintptr_t>(v.base) & 15) != 0) return fail(PVAE_ERR_INVALID, "operand base %p is not 16-byte aligned", (const void*)v.base);
  if ((v.ld & 7) != 0) return fail(PVAE_ERR_INVALID, "operand row stride %lld is not a multiple of 8 elements", (long long)v.ld);
  if (v.planes > 1 && (v.ps & 7) != 0) return fail(PVAE_ERR_INVALID, "operand plane stride %lld is not a multiple of 8 elements", (long long)v.ps);
  if (v.width <= 0 || v.rows <= 0) return fail(PVAE_ERR_INVALID, "empty operand (%d x %lld)", v.width, (long long)v.rows);
  cuuint64_t dims[3] = {(cuuint64_t)v.width, (cuuint64_t)v.rows, (cuuint64_t)v.planes};
  cuuint64_t strides[2] = {(cuuint64_t)v.ld * 2, (cuuint64_t)(v.planes > 1 ? v.ps : v.ld * v.rows) * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(v.base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PVAE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for %d x %lld x %d, ld %lld, box %d x %d", (int)r, v.width,
                (long long)v.rows, v.planes, (long long)v.ld, box_cols, box_rows);
  return PVAE_OK;
}

// One GEMM launch, described on the host.
struct GemmDesc {
  View A[2];
  int nseg = 1;
  int a_major = MAJOR_K;
  View B;
  int b_major = MAJOR_K;
  int b_k0[2] = {0, 0};   // where each A segment's k range starts inside B (B k coordinate)
  int b_n0 = 0;           // first n coordinate inside B
  int M = 0, N = 0;       // valid extents of D
  int K[2] = {0, 0};      // valid k extent per segment
  int passes = 1;
  bool split = false;     // split K over CTAs (wgrad)
  bool mseg = false;      // A[0] / A[1] are M segments (MN-major wgrad of a two-segment input) instead of K segments
  int m_gap0 = 0, m_gap = 0;   // rows [m_gap0, m_gap0 + m_gap) of D do not exist in the output (alignment gap of a fused input), later rows move up
  EpiParams epi;
  GemmDesc() { memset(&epi, 0, sizeof(epi)); }
};

// L2-aware traversal: the big batch-indexed tensors of a step (134 MB each at batch 65536) do not fit the 126 MB L2, but
// most of one does.  A GEMM that walks the batch in the direction OPPOSITE to the last kernel that touched its main
// batch-indexed operand starts on the rows that kernel touched last, i.e. the ones still resident.  The table remembers, per
// buffer, the direction of the last walk; the launch sequence of a step is static, so a captured graph freezes a
// consistent zig-zag.  One table per engine (Device): pointers of a destroyed engine cannot alias a later one's.  (Order never changes results: tiles are independent, weight gradients accumulate with atomics.)
struct WalkTable {
  static constexpr int N = 64;
  const void* ptr[N];
  int dir[N];
  int n = 0;
  int last(const void* p) const { for (int i = 0; i < n; ++i) if (ptr[i] == p) return dir[i]; return -1; }
  void set(const void* p, int d) {
    if (!p) return;
    for (int i = 0; i < n; ++i) if (ptr[i] == p) { dir[i] = d; return; }
    if (n < N) { ptr[n] = p; dir[n] = d; ++n; }
  }
};

struct Device {
  int id = 0;
  int sms = 148;
  int mn_bn_align = 64;   // UMMA N granularity used when B is MN-major (PVAE_MN_BN_ALIGN)
  int tma_epilogue = 1;   // bf16 outputs leave through shared memory + TMA stores (PVAE_TMA_EPILOGUE=0 disables)
  int cluster = 2;        // CTA pairs run tcgen05.mma.cta_group::2 on two adjacent M tiles (PVAE_CLUSTER=1 disables)
  int dbg = 0;            // PVAE_DBG: epilogue timing experiments (see GemmParams::dbg)
  int bn_cap = MAX_BN;    // PVAE_BN_CAP: widest N tile (experiments)
  int pdl = 1;            // PVAE_PDL=0: plain stream-ordered GEMM launches
  int snake = 1;          // PVAE_SNAKE=0: every GEMM walks the batch front to back (see batch_direction)
  int cs_mma = 0;         // PVAE_CS_MMA=1: bias-gradient column sums on mma.sync instead of lane adds (slower, kept for experiments)
  int prefetch = 0;       // PVAE_PREFETCH (debug-hooks build only): L2 prefetch distance of the streamed operands in units (-1 = by K depth); measured slower (profiles/r02_bench.md)
  int fast_epi = 1;       // PVAE_FAST_EPI=0: never use the lean ReLU store / dgrad epilogue (A/B experiments)
  int small_fwd = 1;      // PVAE_SMALL_FWD=0: batches <= 16 of the inference API also take the tensor-core path
  int reserved_sms = 0;   // SMs the GEMM grids leave free while the gradient exchange of a data-parallel step runs beside them (pvae_set_exchange)
  int deterministic = 0;  // pvae_set_deterministic: no split-K, bias gradients by ordered column sums (run-to-run bit-identical gradients)
  float* det_scratch = nullptr;   // [DET_CHUNKS][cols] partial column sums (the engine's small scratch buffer)
  int det_scratch_elems = 0;
  int32_t* cursor = nullptr;   // device int: first row of the current mini-batch
  bool attr_set = false;
  mutable WalkTable walk;
};

static int batch_direction(const Device& dev, const GemmDesc& d) {
  if (!dev.snake) return 0;
  // main batch-indexed input: A (row-streaming GEMMs: A is [batch x K]); weight gradients stream A and B along K = batch,
  // the wider one decides
  const void* main_in = d.A[0].base;
  if (d.split && d.B.width >= d.A[0].width + (d.nseg > 1 ? d.A[1].width : 0)) main_in = d.B.base;   // (tie: the gradient is the fresher one)
  const int prev = dev.walk.last(main_in);
  const int dir = prev < 0 ? 0 : 1 - prev;
  dev.walk.set(d.A[0].base, dir);
  if (d.nseg > 1) dev.walk.set(d.A[1].base, dir);
  if (d.split) dev.walk.set(d.B.base, dir);
  dev.walk.set(d.epi.out, dir);
  dev.walk.set(d.epi.out2, dir);
  if (d.epi.type != EPI_WGRAD) dev.walk.set(d.epi.out_f32, dir);
  dev.walk.set(d.epi.aux, dir);
  dev.walk.set(d.epi.add, dir);
  return dir;
}

static int launch_gemm(const Device& dev, const GemmDesc& d, cudaStream_t st) {
  CKR(resolve_driver());
  if (d.M <= 0 || d.N <= 0) return fail(PVAE_ERR_INVALID, "empty GEMM %d x %d", d.M, d.N);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.a_major = d.a_major;
  p.b_major = d.b_major;
  p.passes = d.passes;
  p.m_tiles = cdiv(d.M, BM);
  if (d.mseg) {
    if (d.a_major != MAJOR_MN || d.nseg != 2) return fail(PVAE_ERR_INVALID, "M segments need an MN-major two-segment A operand");
    p.m_seg_tiles = cdiv(d.A[0].width, BM);
    p.m_seg_rows[0] = d.A[0].width; p.m_seg_rows[1] = d.A[1].width;
    p.m_seg_out0 = d.A[0].width;
    p.m_tiles = p.m_seg_tiles + cdiv(d.A[1].width, BM);
  }
  // CTA pairs (tcgen05.mma.cta_group::2, 256 x bn per instruction) whenever there are two M tiles to pair
  const int cluster = (dev.cluster == 2 && p.m_tiles >= 2) ? 2 : 1;
  p.cg = cluster;
  p.dbg = dev.dbg;
  {  // PVAE_TRACE_IDX=n: record the role timeline (GemmParams::dbg bit 5) of the n-th GEMM launch of this process only
    static const char* sel = getenv("PVAE_TRACE_IDX");
    static long gemm_idx = 0;
    if (sel && atol(sel) == gemm_idx) p.dbg |= 32;
    ++gemm_idx;
  }
  int n_tiles = cdiv(d.N, dev.bn_cap);
  // bn: K-major B -- one N tile: any multiple of 16 (the TMA store clips at the tensor edge), several: whole 64-column
  // sub-tiles; MN-major B -- whole 64-column TMA boxes per CTA (each CTA of a pair stages bn / 2 columns)
  int bn;
  if (d.b_major == MAJOR_MN) bn = rup(cdiv(d.N, n_tiles), (n_tiles > 1 ? 64 : dev.mn_bn_align) * cluster);
  else bn = rup(cdiv(d.N, n_tiles), n_tiles > 1 ? 64 : 16);
  if (bn > MAX_BN) bn = MAX_BN;
  p.cs_mma = dev.cs_mma;
  p.pf_dist = dev.prefetch;
  p.pf_b = d.epi.type == EPI_WGRAD;
  p.reverse = batch_direction(dev, d);
  n_tiles = cdiv(d.N, bn);
  p.n_tiles = n_tiles;
  p.bn = bn;
  for (int s = 0; s < 2; ++s) {
    if (s < d.nseg) {
      p.kb[s] = (d.mseg && s == 1) ? 0 : cdiv(d.K[s], BK);
      p.klen[s] = d.K[s];
      p.a_c0[s] = 0;
      p.a_r0[s] = 0;
      p.a_dyn[s] = d.A[s].dyn;
      p.b_k0[s] = d.b_k0[s];
      CKR(encode_map(&p.tmA[s], d.A[s], 64, d.a_major == MAJOR_K ? BM : BK));
    }
  }
  if (d.nseg == 1) p.tmA[1] = p.tmA[0];
  p.b_n0 = d.b_n0;
  p.b_dyn = d.B.dyn;
  p.m_gap0 = d.m_gap0; p.m_gap = d.m_gap;
  CKR(encode_map(&p.tmB, d.B, 64, d.b_major == MAJOR_K ? bn / cluster : BK));
  const int kb_total = p.kb[0] + p.kb[1];
  const int iters = kb_total * d.passes;
  int splits = 1;
  if (d.split && !dev.deterministic) {      // (deterministic mode: one CTA pair walks the whole K range of its tile, in order)
    // Split K so that the work units fill whole waves of the persistent grid: the kernel lasts as long as its busiest
    // CTA (pair), i.e. waves(s) * (k-blocks per unit + the unit's non-overlapped epilogue share).  1024x1024 wgrad at
    // batch 65536: 16 pair tiles on 74 pairs -- 10 splits need 3 waves (72 % busy), 9 splits need 2 (97 %).
    const int tiles = cdiv(p.m_tiles, cluster) * n_tiles;
    const int slots = (dev.sms - dev.reserved_sms) / cluster;
    const int epi_cost = 2;                       // k-block equivalents of the fp32 red.add epilogue that is not hidden
    const int max_splits = iters < 96 ? iters : 96;
    long best = -1;
    for (int s = 1; s <= max_splits; ++s) {
      const long cost = (long)cdiv(tiles * s, slots) * (cdiv(iters, s) + epi_cost);
      if (best < 0 || cost < best) { best = cost; splits = s; }
    }
  }
  p.splits = splits;
  // L2 prefetch distance: short units (thin K) are over before a DRAM round trip completes -- look two units ahead
  if (p.pf_dist < 0) p.pf_dist = cdiv(iters, splits) <= 6 ? 2 : 1;
  p.row_cursor = dev.cursor;
  p.epi = d.epi;
  p.epi.m_valid = d.M;
  p.epi.n_valid = d.N;
  // deterministic mode: bias gradients are not summed by the epilogue (atomics across CTAs) but by ordered column sums of the
  // primary output after the launch
  float* det_colsum = nullptr;
  if (dev.deterministic && p.epi.colsum && p.epi.out && dev.det_scratch && (int64_t)DET_CHUNKS * d.N <= dev.det_scratch_elems) {
    det_colsum = p.epi.colsum;
    p.epi.colsum = nullptr;
  }
  const EpiParams& e = p.epi;
  // TMA epilogue: one precision plane, bf16 primary output; ReLU dgrad needs the sign-bit mask its forward wrote
  const bool has_aux = e.type == EPI_MSE || (e.type == EPI_DGRAD && e.act != ACT_LINEAR && e.act != ACT_RELU);
  const bool tma = dev.tma_epilogue && e.type != EPI_WGRAD && e.out != nullptr && e.out_planes == 1 &&
                   !(e.type == EPI_STORE && e.out2 != nullptr) &&       // (swish forward keeping its pre-activation: direct epilogue)
                   !(e.type == EPI_DGRAD && e.act == ACT_RELU && e.mask == nullptr) &&
                   (!has_aux || (e.aux != nullptr && e.aux_planes == 1 && (reinterpret_cast<uintptr_t>(e.aux) & 15) == 0));
  if (tma) {
    View o; o.base = e.out; o.ld = e.out_ld; o.ps = e.out_ps; o.planes = 1; o.width = d.N; o.rows = d.M;
    CKR(encode_map(&p.tmOut, o, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));   // one epilogue warp's 32 x 32 chunk
    if (has_aux) {
      View a; a.base = e.aux; a.ld = e.aux_ld; a.ps = e.aux_ps; a.planes = 1; a.width = d.N; a.rows = e.aux_rows ? e.aux_rows : d.M;
      CKR(encode_map(&p.tmAux, a, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
    }
  }
  const int units = cdiv(p.m_tiles, cluster) * n_tiles * splits;       // units per CTA (pair)
  const int slots = (dev.sms - dev.reserved_sms) / cluster;
  const int grid = (units < slots ? units : slots) * cluster;
  // lean epilogue: ReLU store / dgrad whose tiles have no ragged edge and no optional operand (the kernel's FAST block)
  const bool fast = dev.fast_epi && tma && cluster == 2 && e.act == ACT_RELU && (e.type == EPI_STORE || e.type == EPI_DGRAD) &&
                    d.M % (BM * cluster) == 0 && d.N % 32 == 0 && bn % 32 == 0 && !d.mseg && d.m_gap == 0 && e.add == nullptr &&
                    e.out_f32 == nullptr && e.mask != nullptr && (e.type == EPI_DGRAD || e.bias != nullptr) && dev.cs_mma == 0 &&
                    splits == 1 && (!DEBUG_HOOKS || (p.dbg & ~32) == 0);     // (role-timeline stamps exist in the lean block too)
  GemmKernelFn fn = select_kernel(p.epi.type, p.epi.act, tma, cluster, fast);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (dev.pdl) {      // programmatic dependent launch: this grid's prologue overlaps the previous kernel's tail (see the kernel)
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  static const int log_level = getenv("PVAE_LOG_GEMM") ? atoi(getenv("PVAE_LOG_GEMM")) : 0;   // 1: print every launch, 2: and synchronise after it
  if (log_level)
    fprintf(stderr, "[pvae_gemm] epi %d act %d tma %d cg %d fast %d | M %d N %d K %d+%d majors %d%d passes %d | m_tiles %d n_tiles %d bn %d splits %d grid %d | b_k0 %d,%d b_n0 %d mseg %d rev %d\n",
            p.epi.type, p.epi.act, (int)tma, cluster, (int)fast, d.M, d.N, d.K[0], d.nseg > 1 ? d.K[1] : 0, d.a_major, d.b_major, d.passes, p.m_tiles, n_tiles, bn,
            splits, grid, p.b_k0[0], p.b_k0[1], p.b_n0, p.m_seg_tiles, p.reverse);
  CK(cudaLaunchKernelEx(&cfg, fn, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  if (det_colsum) {
    colsum_det_partial_kernel<<<dim3(cdiv(d.N, 32), DET_CHUNKS), 256, 0, st>>>(p.epi.out, p.epi.out_ld, p.epi.out_ps, p.epi.out_planes, d.M, d.N, dev.det_scratch);
    colsum_det_final_kernel<<<cdiv(d.N, 128), 128, 0, st>>>(dev.det_scratch, d.N, det_colsum);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    CK(cudaGetLastError());
  }
  if (log_level >= 2) CK(cudaStreamSynchronize(st));
  return PVAE_OK;
}

// ------------------------------------------------------------------------------------------------------------------
struct Net {
  int n_layers = 0;
  int in_dim = 0;
  int k0 = 0, k1 = 0;          // layer-0 input segments (k1 == 0: single segment)
  int K0pad = 0;               // shadow column where segment 1 starts
  int in_dims[PVAE_MAX_LAYERS], out_dims[PVAE_MAX_LAYERS], acts[PVAE_MAX_LAYERS], kpad[PVAE_MAX_LAYERS];
  const float* W[PVAE_MAX_LAYERS];
  const float* b[PVAE_MAX_LAYERS];
  float* grad = nullptr;
  int64_t gW[PVAE_MAX_LAYERS], gb[PVAE_MAX_LAYERS];
  int64_t grad_elems = 0;
  __nv_bfloat16* Wsh[PVAE_MAX_LAYERS];
  int64_t wsh_ps[PVAE_MAX_LAYERS];
  __nv_bfloat16* act[PVAE_MAX_LAYERS];
  __nv_bfloat16* g[PVAE_MAX_LAYERS];
  uint32_t* mask[PVAE_MAX_LAYERS];   // ReLU sign bits of act[l], [mask_ld[l] words of 32 columns][max_batch]
  int act_ld[PVAE_MAX_LAYERS];
  int mask_ld[PVAE_MAX_LAYERS];
  const float* beta = nullptr;      // [n_layers] fp32: beta of layer l's swish activation (pvae_bind_act_params), null = 1 everywhere
  float* dbeta = nullptr;           // [n_layers] fp32 gradient accumulators, or null
  bool bound = false;
  bool generic = false;        // input widths given explicitly (stand-alone FC): usable through pvae_fc_forward only
};

struct NetIO {     // what feeds layer 0: one or two column segments (the reference's torch.cat inputs)
  View seg[2];
  int nseg = 1;
};

}  // namespace pvae

using namespace pvae;

struct pvae_engine {
  pvae_model_desc desc;
  Device dev;
  Net nets[PVAE_NUM_NETS];
  int planes = 1, passes = 1;
  int max_batch = 0;
  int dsb = 0, dsb8 = 0, dsbp = 0, da = 0, z = 0, te_out = 0;   // dsbp: 128-byte aligned column where s_{t+1} starts inside a transition row
  // workspace
  void* ws = nullptr;
  size_t ws_bytes = 0;
  float* ml = nullptr;      // [B][te_out] fp32 encoder output (mu | logvar)
  float* eps = nullptr;     // [B][z]
  float* dz = nullptr;      // [B][z]
  __nv_bfloat16* zb = nullptr;    int zb_ld = 0;
  __nv_bfloat16* ahat = nullptr;  int a_ld = 0;
  __nv_bfloat16* ga = nullptr;
  __nv_bfloat16* xin = nullptr;   int x_ld = 0;
  __nv_bfloat16* ain = nullptr;
  __nv_bfloat16* fut = nullptr;   int f_ld = 0;      // predicted next state of a rollout step (the next step's body state), bf16 planes
  // rollout (lookahead > 1): `roll_slots` copies of the workspace, one per (step, world-model pass)
  uint8_t* roll_ws = nullptr;
  int roll_slots = 0;
  // transitions
  const __nv_bfloat16* tbuf = nullptr;
  int64_t tbuf_rows = 0;
  int tx_ld = 0, ty_ld = 0;
  // scalars
  double* acc = nullptr;    // [4] loss accumulators
  bool acc_dirty = false;   // a step was entered but its finalize kernel (which clears acc) was not enqueued
  unsigned int* adam_counter = nullptr;   // finished-blocks counter of adam_net_kernel
  float* small_scratch = nullptr;            // fp32 [2][16][SMALL_SCRATCH_COLS]: z and action rows between the chains of the small-batch forward
  // overlapped gradient exchange (pvae_set_exchange): the early range is exchanged on a side stream while the last backward GEMMs run
  struct {
    bool on = false;
    SymmArgs args;                 // peers, rank, world, flags; off / count = the early range
    int ctas = 8;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool forked = false;
  } xchg;
  unsigned long long* noise_ctr = nullptr;   // device-side Philox offset counter (pvae_noise_counter)
  bool noise_auto = false;
  unsigned long long noise_stride = 1;
};

namespace pvae {

static int64_t plane_elems(const pvae_engine* h, int ld) { return (int64_t)h->max_batch * ld; }

// carve the workspace; with base == nullptr only the size is computed
static size_t carve(pvae_engine* h, uint8_t* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> uint8_t* {
    uint8_t* p = base ? base + off : nullptr;
    off += (bytes + 1023) & ~size_t(1023);
    return p;
  };
  const int64_t B = h->max_batch;
  for (int n = 0; n < PVAE_NUM_NETS; ++n) {
    Net& net = h->nets[n];
    for (int l = 0; l < net.n_layers; ++l) {
      net.act_ld[l] = rup(net.out_dims[l], 64);
      const size_t bytes = (size_t)h->planes * B * net.act_ld[l] * 2;
      net.act[l] = (l < net.n_layers - 1) ? reinterpret_cast<__nv_bfloat16*>(take(bytes)) : nullptr;
      net.mask_ld[l] = rup(cdiv(net.out_dims[l], 32), 2);
      net.mask[l] = (l < net.n_layers - 1 && net.acts[l] == ACT_RELU)
                        ? reinterpret_cast<uint32_t*>(take((size_t)B * net.mask_ld[l] * 4)) : nullptr;
      net.g[l] = (n != PVAE_NET_VALUE_BRANCH) ? reinterpret_cast<__nv_bfloat16*>(take(bytes)) : nullptr;
    }
  }
  h->ml = reinterpret_cast<float*>(take((size_t)B * h->te_out * 4));
  h->eps = reinterpret_cast<float*>(take((size_t)B * h->z * 4));
  h->dz = reinterpret_cast<float*>(take((size_t)B * h->z * 4));
  h->zb_ld = rup(h->z, 64);
  h->a_ld = rup(h->da, 64);
  h->x_ld = rup(h->dsbp + h->dsb, 64);
  for (int n = 0; n < PVAE_NUM_NETS; ++n)
    if (h->nets[n].n_layers) { const int need = rup(rup(h->nets[n].k0, 64) + h->nets[n].k1, 64); if (need > h->x_ld) h->x_ld = need; }
  h->zb = reinterpret_cast<__nv_bfloat16*>(take((size_t)h->planes * B * h->zb_ld * 2));
  h->ahat = reinterpret_cast<__nv_bfloat16*>(take((size_t)h->planes * B * h->a_ld * 2));
  h->ga = reinterpret_cast<__nv_bfloat16*>(take((size_t)h->planes * B * h->a_ld * 2));
  h->xin = reinterpret_cast<__nv_bfloat16*>(take((size_t)h->planes * B * h->x_ld * 2));
  h->ain = reinterpret_cast<__nv_bfloat16*>(take((size_t)h->planes * B * h->a_ld * 2));
  h->f_ld = rup(h->dsb, 64);
  h->fut = reinterpret_cast<__nv_bfloat16*>(take((size_t)h->planes * B * h->f_ld * 2));
  return off;
}

static View ws_view(const pvae_engine* h, const __nv_bfloat16* p, int ld, int width, int batch) {
  View v;
  v.base = p; v.ld = ld; v.ps = plane_elems(h, ld); v.planes = h->planes; v.width = width; v.rows = batch; v.dyn = 0;
  return v;
}
static View shadow_view(const pvae_engine* h, const Net& net, int l) {
  View v;
  v.base = net.Wsh[l]; v.ld = net.kpad[l]; v.ps = net.wsh_ps[l]; v.planes = h->planes; v.width = net.kpad[l];
  v.rows = net.out_dims[l]; v.dyn = 0;
  return v;
}
static void set_out(EpiParams& e, const pvae_engine* h, __nv_bfloat16* p, int ld) {
  e.out = p; e.out_ld = ld; e.out_ps = plane_elems(h, ld); e.out_planes = h->planes;
}
static void set_out2(EpiParams& e, const pvae_engine* h, __nv_bfloat16* p, int ld) {
  e.out2 = p; e.out2_ld = ld; e.out2_ps = plane_elems(h, ld); e.out2_planes = h->planes;
}
static void set_aux(EpiParams& e, const View& v, int col0) {
  e.aux = v.base + col0; e.aux_ld = v.ld; e.aux_ps = v.ps; e.aux_planes = v.planes; e.aux_dyn = v.dyn; e.aux_rows = v.rows;
}

// views of the resident transition buffer: x planes then y planes
static View tx_view(const pvae_engine* h, int col0, int width) {
  View v;
  v.base = h->tbuf + col0; v.ld = h->tx_ld; v.ps = h->tbuf_rows * h->tx_ld; v.planes = h->planes; v.width = width;
  v.rows = h->tbuf_rows; v.dyn = 1;
  return v;
}
static View ty_view(const pvae_engine* h, int width) {
  View v;
  v.base = h->tbuf + (int64_t)h->planes * h->tbuf_rows * h->tx_ld; v.ld = h->ty_ld; v.ps = h->tbuf_rows * h->ty_ld;
  v.planes = h->planes; v.width = width; v.rows = h->tbuf_rows; v.dyn = 1;
  return v;
}

// ---- forward through one FC stack (rllib_model_torch.FC.forward, rllib_model_torch.py:274-275) -----------------------
// `last` is the epilogue of the output layer (type/outputs/loss wiring chosen by the caller).
// keep_preact: a backward pass follows (swish layers then store their pre-activation in the layer's gradient buffer).
static int net_forward(pvae_engine* h, Net& net, const NetIO& in, int batch, const EpiParams& last, cudaStream_t st, bool keep_preact = false) {
  if (!net.bound) return fail(PVAE_ERR_STATE, "net not bound (pvae_bind_net)");
  for (int l = 0; l < net.n_layers; ++l) {
    GemmDesc d;
    d.a_major = MAJOR_K;
    if (l == 0) {
      d.nseg = in.nseg;
      for (int s = 0; s < in.nseg; ++s) { d.A[s] = in.seg[s]; d.K[s] = in.seg[s].width; }
      d.b_k0[0] = 0; d.b_k0[1] = net.K0pad;
    } else {
      d.nseg = 1;
      d.A[0] = ws_view(h, net.act[l - 1], net.act_ld[l - 1], net.out_dims[l - 1], batch);
      d.K[0] = net.out_dims[l - 1];
    }
    d.B = shadow_view(h, net, l);
    d.b_major = MAJOR_K;
    d.M = batch; d.N = net.out_dims[l];
    d.passes = h->passes;
    if (l < net.n_layers - 1) {
      d.epi.type = EPI_STORE; d.epi.act = net.acts[l]; d.epi.bias = net.b[l];
      set_out(d.epi, h, net.act[l], net.act_ld[l]);
      d.epi.mask = net.mask[l]; d.epi.mask_ld = h->max_batch;
      if (net.acts[l] == ACT_SWISH && net.g[l] && keep_preact) set_out2(d.epi, h, net.g[l], net.act_ld[l]);
    } else {
      d.epi = last;
      d.epi.act = net.acts[l]; d.epi.bias = net.b[l];
    }
    if (net.acts[l] == ACT_SWISH && net.beta) d.epi.act_param = net.beta + l;
    CKR(launch_gemm(h->dev, d, st));
  }
  return PVAE_OK;
}

// The exchange kernel of a launch: the bulk-copy variant (default; measured 54 -> 38 us per 6 MB call on 8 GPUs, 118 -> 86 us per 24 MB,
// profiles/r02_scaling.md) or per-thread peer loads (PVAE_SYMM_BULK=0, and whenever the switch reduces: multimem).
static bool symm_bulk() {
  static const int bulk = [] { const char* e = getenv("PVAE_SYMM_BULK"); return e ? atoi(e) : 1; }();
  return bulk != 0;
}
static int launch_symm(const SymmArgs& a, int ctas, cudaStream_t st) {
  const bool bulk = symm_bulk();
  if (bulk && !a.mc) {
    static const cudaError_t attr = cudaFuncSetAttribute(symm_allreduce_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ARB_SMEM_BYTES);
    CK(attr);
    symm_allreduce_bulk_kernel<<<ctas, ARB_THREADS, ARB_SMEM_BYTES, st>>>(a);
  } else {
    symm_allreduce_kernel<<<ctas, AR_THREADS, 0, st>>>(a);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return PVAE_OK;
}

// ---- overlapped gradient exchange (data parallel): fork = the gradients of the early range are complete once everything launched so
// far on `st` has run -> exchange them on the side stream with a few CTAs while the remaining GEMMs run on the other SMs; join before
// the step ends.  Both are stream-ordered events: capturable (fork / join inside one graph).
static int exchange_fork(pvae_engine* h, cudaStream_t st) {
  if (!h->xchg.on || h->xchg.forked) return PVAE_OK;
  CK(cudaEventRecord(h->xchg.ev_fork, st));
  CK(cudaStreamWaitEvent(h->xchg.side, h->xchg.ev_fork, 0));
  CKR(launch_symm(h->xchg.args, h->xchg.ctas, h->xchg.side));
  CK(cudaEventRecord(h->xchg.ev_join, h->xchg.side));
  h->xchg.forked = true;
  h->dev.reserved_sms = h->xchg.ctas;              // the GEMMs launched from here on leave that many SMs to the exchange kernel
  return PVAE_OK;
}
static int exchange_join(pvae_engine* h, cudaStream_t st) {
  h->dev.reserved_sms = 0;
  if (!h->xchg.forked) return PVAE_OK;
  h->xchg.forked = false;
  CK(cudaStreamWaitEvent(st, h->xchg.ev_join, 0));
  return PVAE_OK;
}

// ---- backward through one FC stack: g[L-1] (gradient w.r.t. the output layer's pre-activation) is already in place ---
// train: accumulate dW / db into the bound gradient buffer.  in_epi: if non-null, also produce the gradient w.r.t. the
// SECOND input segment of layer 0 (z for the decoder, the action for the world model) with this epilogue.
// in0_epi: likewise for the FIRST input segment (the body state: autoregressive rollouts differentiate through it).
// fork_after: launch the overlapped gradient exchange (exchange_fork) right after the weight gradient of that layer -- by then the
// gradients of layers fork_after .. L-1 (weights and biases) are all in the stream.
static int net_backward(pvae_engine* h, Net& net, const NetIO& in, int batch, bool train, const EpiParams* in_epi, cudaStream_t st,
                        const EpiParams* in0_epi = nullptr, int fork_after = -1) {
  const int L = net.n_layers;
  if (train && !net.grad) return fail(PVAE_ERR_STATE, "net has no gradient buffer bound");
  for (int l = L - 1; l >= 0; --l) {
    const View gl = ws_view(h, net.g[l], net.act_ld[l], net.out_dims[l], batch);
    if (train) {
      // dW[l]^T [in][out] = input^T . g[l]; the two input segments of layer 0 are M segments of one launch (g[0] is read once)
      GemmDesc d;
      d.a_major = MAJOR_MN; d.b_major = MAJOR_MN;
      d.K[0] = batch; d.K[1] = batch;
      d.passes = h->passes;
      d.split = true;
      d.epi.type = EPI_WGRAD;
      d.epi.out_f32 = net.grad + net.gW[l];
      d.epi.f32_atomic = 1;
      d.epi.f32_sm = 1; d.epi.f32_sn = net.in_dims[l];
      d.B = gl;
      d.N = net.out_dims[l];
      if (l == 0 && in.nseg == 2) {
        d.nseg = 2; d.mseg = true;
        d.A[0] = in.seg[0]; d.A[1] = in.seg[1];
        d.M = in.seg[0].width + in.seg[1].width;
      } else {
        d.nseg = 1;
        // the operand that comes from the resident buffer walks the batch along k: the dynamic cursor applies to its rows
        d.A[0] = (l == 0) ? in.seg[0] : ws_view(h, net.act[l - 1], net.act_ld[l - 1], net.out_dims[l - 1], batch);
        d.M = d.A[0].width;
        if (l == 0 && net.k1 && d.M > net.k0) { d.m_gap0 = net.k0; d.m_gap = net.K0pad - net.k0; }   // fused (first | gap | second) input row
      }
      CKR(launch_gemm(h->dev, d, st));
      if (l == fork_after) CKR(exchange_fork(h, st));
    }
    if (l > 0) {
      GemmDesc d;
      d.a_major = MAJOR_K; d.b_major = MAJOR_MN;
      d.A[0] = gl; d.K[0] = net.out_dims[l];
      d.B = shadow_view(h, net, l);
      d.M = batch; d.N = net.out_dims[l - 1];
      d.passes = h->passes;
      d.epi.type = EPI_DGRAD; d.epi.act = net.acts[l - 1];
      // act' needs the forward output -- for swish the pre-activation, which the forward pass left in g[l - 1] (overwritten in place here)
      const View al = ws_view(h, net.acts[l - 1] == ACT_SWISH ? net.g[l - 1] : net.act[l - 1], net.act_ld[l - 1], net.out_dims[l - 1], batch);
      if (net.acts[l - 1] != ACT_LINEAR) set_aux(d.epi, al, 0);
      d.epi.mask = net.mask[l - 1]; d.epi.mask_ld = h->max_batch;
      set_out(d.epi, h, net.g[l - 1], net.act_ld[l - 1]);
      d.epi.colsum = train ? net.grad + net.gb[l - 1] : nullptr;
      if (net.acts[l - 1] == ACT_SWISH && net.beta) {
        d.epi.act_param = net.beta + (l - 1);
        d.epi.act_grad = (train && net.dbeta) ? net.dbeta + (l - 1) : nullptr;
      }
      CKR(launch_gemm(h->dev, d, st));
    } else if (in_epi) {
      GemmDesc d;
      d.a_major = MAJOR_K; d.b_major = MAJOR_MN;
      d.A[0] = gl; d.K[0] = net.out_dims[0];
      d.B = shadow_view(h, net, 0);
      d.b_n0 = net.K0pad;
      d.M = batch; d.N = net.k1;
      d.passes = h->passes;
      d.epi = *in_epi;
      CKR(launch_gemm(h->dev, d, st));
    }
    if (l == 0 && in0_epi) {
      GemmDesc d;
      d.a_major = MAJOR_K; d.b_major = MAJOR_MN;
      d.A[0] = gl; d.K[0] = net.out_dims[0];
      d.B = shadow_view(h, net, 0);
      d.b_n0 = 0;
      d.M = batch; d.N = net.k0;
      d.passes = h->passes;
      d.epi = *in0_epi;
      CKR(launch_gemm(h->dev, d, st));
    }
  }
  return PVAE_OK;
}

static int check_trainable_acts(const Net& net) {
  if (net.n_layers > 0 && net.acts[net.n_layers - 1] != ACT_LINEAR)
    return fail(PVAE_ERR_INVALID, "the training steps need a linear output layer (train_physics_vae.py:188-190)");
  return PVAE_OK;
}

constexpr int SMALL_SCRATCH_COLS = 4096;

// ---- latency path: one cluster kernel per FC stack for <= 16 rows (pvae_small.cuh) ------------------------------------------------
// in0 / in1: fp32 rows of the two layer-0 input segments (in1 may be null when the net has one segment)
static bool small_ok(const pvae_engine* h, const Net& net, int batch) {
  // (measured: past 4 rows the CUDA-core dot products lose to the 26 us tensor-core path -- every weight piece needs one shared-memory read per row)
  if (!h->dev.small_fwd || batch > SF_SMALL_ROWS || net.n_layers > SF_MAX_LAYERS) return false;
  int width = 0;
  for (int l = 0; l < net.n_layers; ++l) { width = net.kpad[l] > width ? net.kpad[l] : width; width = net.out_dims[l] > width ? net.out_dims[l] : width; }
  int bt = 1; while (bt < batch) bt <<= 1;
  return (size_t)2 * bt * rup(width, 32) * sizeof(float) <= 200 * 1024;
}
static int small_chain(pvae_engine* h, Net& net, int batch, const float* in0, int64_t in0_ld, const float* in1, int64_t in1_ld, float* out,
                       int64_t out_ld, cudaStream_t st) {
  if (!net.bound) return fail(PVAE_ERR_STATE, "net not bound (pvae_bind_net)");
  SmallNet sn;
  memset(&sn, 0, sizeof(sn));
  sn.n_layers = net.n_layers; sn.planes = h->planes;
  sn.k0 = net.k0; sn.k1 = net.k1; sn.K0pad = net.K0pad;
  int width = 0;
  for (int l = 0; l < net.n_layers; ++l) {
    sn.L[l].W = net.Wsh[l]; sn.L[l].ps = net.wsh_ps[l]; sn.L[l].bias = net.b[l];
    sn.L[l].act_param = (net.acts[l] == ACT_SWISH && net.beta) ? net.beta + l : nullptr;
    sn.L[l].out = net.out_dims[l]; sn.L[l].kpad = net.kpad[l]; sn.L[l].act = net.acts[l];
    width = net.kpad[l] > width ? net.kpad[l] : width;
    width = net.out_dims[l] > width ? net.out_dims[l] : width;
  }
  sn.width = rup(width, 32);
  static const int small_trace = getenv("PVAE_SMALL_TRACE") ? atoi(getenv("PVAE_SMALL_TRACE")) : 0;
  sn.trace = small_trace;
  int bt = 1; while (bt < batch) bt <<= 1;
  const size_t smem = (size_t)2 * bt * sn.width * sizeof(float);
  void (*fn)(const SmallNet, const float*, int64_t, const float*, int64_t, int, float*, int64_t) =
      bt == 1 ? small_fc_kernel<1> : bt == 2 ? small_fc_kernel<2> : small_fc_kernel<4>;
  static bool attr_done[3] = {false, false, false};
  const int ai = bt == 1 ? 0 : bt == 2 ? 1 : 2;
  if (!attr_done[ai]) { CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_done[ai] = true; }
  fn<<<SF_CLUSTER, SF_THREADS, smem, st>>>(sn, in0, in0_ld, in1, in1_ld, batch, out, out_ld);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return PVAE_OK;
}

static int grid_for(int64_t total, int threads, int sms) {
  int64_t g = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)sms * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pvae

// ====================================================================================================================
// C ABI
// ====================================================================================================================
extern "C" {

const char* pvae_last_error(void) { return g_err; }
int pvae_abi_version(void) { return PVAE_ABI_VERSION; }
uint64_t pvae_launch_count(void) { return g_launches.load(); }

static int ensure_kernel_attr(Device& dev) {
  if (dev.attr_set) return PVAE_OK;
  for (int epi = 0; epi < 4; ++epi)
    for (int act = 0; act <= ACT_SWISH; ++act)
      for (int tma = 0; tma < 2; ++tma)
        for (int cg = 1; cg <= 2; ++cg)
          CK(cudaFuncSetAttribute(select_kernel(epi, act, tma != 0, cg), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CK(cudaFuncSetAttribute(select_kernel(EPI_STORE, ACT_RELU, true, 2, true), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CK(cudaFuncSetAttribute(select_kernel(EPI_DGRAD, ACT_RELU, true, 2, true), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  dev.attr_set = true;
  return PVAE_OK;
}

static int init_device(Device& dev, int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(PVAE_ERR_CUDA, "no CUDA device: libpvae_sm100 has no CPU fallback (%s)", cudaGetErrorString(e));
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(PVAE_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  dev.id = device;
  dev.sms = prop.multiProcessorCount;
  const char* env = getenv("PVAE_MN_BN_ALIGN");
  if (env) { int v = atoi(env); if (v == 16 || v == 32 || v == 64) dev.mn_bn_align = v; }
  env = getenv("PVAE_TMA_EPILOGUE");
  if (env) dev.tma_epilogue = atoi(env) != 0;
  env = getenv("PVAE_DBG");
  if (env) dev.dbg = atoi(env);
  env = getenv("PVAE_BN_CAP");
  if (env) { int v = atoi(env); if (v >= 64 && v <= MAX_BN && v % 64 == 0) dev.bn_cap = v; }
  env = getenv("PVAE_PDL");
  if (env) dev.pdl = atoi(env) != 0;
  env = getenv("PVAE_SNAKE");
  if (env) dev.snake = atoi(env) != 0;
  env = getenv("PVAE_CS_MMA");
  if (env) dev.cs_mma = atoi(env) != 0;
  env = getenv("PVAE_FAST_EPI");
  if (env) dev.fast_epi = atoi(env) != 0;
  env = getenv("PVAE_PREFETCH");
  if (env) { int v = atoi(env); if (v >= -1 && v <= 8) dev.prefetch = v; }
  env = getenv("PVAE_SMALL_FWD");
  if (env) dev.small_fwd = atoi(env) != 0;
  env = getenv("PVAE_CLUSTER");
  if (env) dev.cluster = atoi(env) == 1 ? 1 : 2;
  CKR(resolve_driver());
  CKR(ensure_kernel_attr(dev));
  return PVAE_OK;
}

int pvae_create(pvae_handle* out, const pvae_model_desc* desc, int device) {
  if (!out || !desc) return fail(PVAE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (desc->dim_state_body <= 0 || desc->dim_action <= 0 || desc->latent_dim <= 0 || desc->max_batch <= 0)
    return fail(PVAE_ERR_INVALID, "dims and max_batch must be positive");
  if (desc->precision != PVAE_PREC_BF16 && desc->precision != PVAE_PREC_BF16X3)
    return fail(PVAE_ERR_INVALID, "unknown precision %d", desc->precision);
  pvae_engine* h = new pvae_engine();
  h->desc = *desc;
  int r = init_device(h->dev, device);
  if (r != PVAE_OK) { delete h; return r; }
  h->planes = desc->precision == PVAE_PREC_BF16X3 ? 2 : 1;
  h->passes = desc->precision == PVAE_PREC_BF16X3 ? 3 : 1;
  h->max_batch = desc->max_batch;
  h->dsb = desc->dim_state_body; h->da = desc->dim_action; h->z = desc->latent_dim;
  h->te_out = desc->latent_prior ? 2 * h->z : h->z;
  // A resident transition row is (s_t | 0.. | a_t | 0.. | s_{t+1} | 0..): s_t and s_{t+1} each start on a 128-byte boundary and rows
  // are whole cache lines, so that every 64-column TMA box row is exactly one L2 line (a 16-byte-aligned 800 B row stride made every
  // box row straddle two).  a_t sits behind s_t at the next 16-byte boundary (TMA box coordinates must be 16-byte aligned, and the
  // layer-0 shadow weights use the same column numbering), which makes the world model's input cat[s_t, a_t] ONE K segment of
  // the row (200 + 45 = 245 columns = 4 k-blocks instead of 4 + 1, and two instead of three M tiles in the layer-0 weight
  // gradient); GEMMs that want s_t alone describe the row with a 197-column tensor map and the hardware zero-fills the rest.
  h->dsb8 = rup(h->dsb, 8);
  h->dsbp = rup(h->dsb8 + h->da, 64);
  h->tx_ld = h->dsbp + rup(h->dsb, 64);
  h->ty_ld = rup(h->da, 64);
  for (int n = 0; n < PVAE_NUM_NETS; ++n) {
    Net& net = h->nets[n];
    const pvae_net_desc& nd = desc->nets[n];
    if (nd.n_layers < 0 || nd.n_layers > PVAE_MAX_LAYERS) { delete h; return fail(PVAE_ERR_INVALID, "net %d: bad layer count %d", n, nd.n_layers); }
    net.n_layers = nd.n_layers;
    if (nd.n_layers == 0) continue;
    switch (n) {
      case PVAE_NET_TASK_ENCODER: net.k0 = h->dsb; net.k1 = h->dsb; break;       // (s_t | s_{t+1}) as two segments
      case PVAE_NET_MOTOR_DECODER: net.k0 = h->dsb; net.k1 = h->z; break;
      case PVAE_NET_WORLD_MODEL: net.k0 = h->dsb; net.k1 = h->da; break;
      default: net.k0 = h->dsb; net.k1 = h->dsb; break;
    }
    if (nd.in_dims[0] > 0) {       // stand-alone FC stack (pvae_fc_forward only): explicit input segments instead of the role's
      if (nd.in_dims[1] < 0) { delete h; return fail(PVAE_ERR_INVALID, "net %d: bad input override", n); }
      net.k0 = nd.in_dims[0]; net.k1 = nd.in_dims[1];
      net.generic = true;
    }
    net.in_dim = net.k0 + net.k1;
    // layer-0 shadow columns: the second input segment's weights start at the next 16-byte boundary behind the first one's (a TMA
    // box coordinate must be a multiple of 16 bytes: an odd start raises an illegal-instruction fault) -- a two-segment A operand
    // addresses them at k offset K0pad, a one-segment A operand that keeps the same gap (resident rows) sees one K range
    net.K0pad = rup(net.k0, 8);
    int64_t goff = 0;
    for (int l = 0; l < nd.n_layers; ++l) {
      if (nd.out_dims[l] <= 0 || nd.acts[l] < 0 || nd.acts[l] > PVAE_ACT_SWISH) { delete h; return fail(PVAE_ERR_INVALID, "net %d layer %d: bad size/activation", n, l); }
      net.in_dims[l] = l == 0 ? net.in_dim : nd.out_dims[l - 1];
      net.out_dims[l] = nd.out_dims[l];
      net.acts[l] = nd.acts[l];
      net.kpad[l] = l == 0 ? rup(net.K0pad + net.k1, 64) : rup(net.in_dims[l], 64);
      net.gW[l] = goff; goff += (int64_t)net.in_dims[l] * net.out_dims[l];
      net.gb[l] = goff; goff += net.out_dims[l];
      net.wsh_ps[l] = (int64_t)net.out_dims[l] * net.kpad[l];
      net.Wsh[l] = nullptr;
    }
    net.grad_elems = goff;
  }
  const int want[PVAE_NUM_NETS] = {h->te_out, h->da, h->dsb, 1};
  for (int n = 0; n < PVAE_NUM_NETS; ++n) {
    Net& net = h->nets[n];
    if (net.n_layers && !net.generic && net.out_dims[net.n_layers - 1] != want[n]) {
      int got = net.out_dims[net.n_layers - 1];
      delete h;
      return fail(PVAE_ERR_INVALID, "net %d: output width %d, expected %d", n, got, want[n]);
    }
  }
  // shadow weights + scalars
  for (int n = 0; n < PVAE_NUM_NETS; ++n) {
    Net& net = h->nets[n];
    for (int l = 0; l < net.n_layers; ++l) {
      cudaError_t e = cudaMalloc(&net.Wsh[l], (size_t)h->planes * net.wsh_ps[l] * 2);
      if (e != cudaSuccess) { pvae_destroy(h); return fail(PVAE_ERR_CUDA, "cudaMalloc(shadow weights) failed: %s", cudaGetErrorString(e)); }
    }
  }
  if (cudaMalloc(&h->acc, 4 * sizeof(double)) != cudaSuccess || cudaMalloc(&h->dev.cursor, sizeof(int32_t)) != cudaSuccess) {
    pvae_destroy(h);
    return fail(PVAE_ERR_CUDA, "cudaMalloc(scalars) failed");
  }
  if (cudaMalloc(&h->adam_counter, sizeof(unsigned int)) == cudaSuccess) cudaMemset(h->adam_counter, 0, sizeof(unsigned int));
  else h->adam_counter = nullptr;
  if (cudaMalloc(&h->small_scratch, 2 * SF_MAX_ROWS * SMALL_SCRATCH_COLS * sizeof(float)) != cudaSuccess) h->small_scratch = nullptr;
  if (cudaMalloc(&h->noise_ctr, sizeof(unsigned long long)) == cudaSuccess) cudaMemset(h->noise_ctr, 0, sizeof(unsigned long long));
  else h->noise_ctr = nullptr;
  cudaMemset(h->dev.cursor, 0, sizeof(int32_t));
  cudaMemset(h->acc, 0, 4 * sizeof(double));
  h->ws_bytes = carve(h, nullptr);
  *out = h;
  return PVAE_OK;
}

int pvae_destroy(pvae_handle h) {
  if (!h) return PVAE_OK;
  for (int n = 0; n < PVAE_NUM_NETS; ++n)
    for (int l = 0; l < h->nets[n].n_layers; ++l)
      if (h->nets[n].Wsh[l]) cudaFree(h->nets[n].Wsh[l]);
  if (h->acc) cudaFree(h->acc);
  if (h->adam_counter) cudaFree(h->adam_counter);
  if (h->noise_ctr) cudaFree(h->noise_ctr);
  if (h->small_scratch) cudaFree(h->small_scratch);
  if (h->xchg.side) cudaStreamDestroy(h->xchg.side);
  if (h->xchg.ev_fork) cudaEventDestroy(h->xchg.ev_fork);
  if (h->xchg.ev_join) cudaEventDestroy(h->xchg.ev_join);
  if (h->dev.cursor) cudaFree(h->dev.cursor);
  delete h;
  return PVAE_OK;
}

int pvae_bind_net(pvae_handle h, int net_id, const float* const* W_dev, const float* const* b_dev, float* grad_flat_dev) {
  if (!h || net_id < 0 || net_id >= PVAE_NUM_NETS) return fail(PVAE_ERR_INVALID, "bad handle / net id");
  Net& net = h->nets[net_id];
  if (net.n_layers == 0) return fail(PVAE_ERR_INVALID, "net %d is absent from the model description", net_id);
  for (int l = 0; l < net.n_layers; ++l) {
    if (!W_dev[l] || !b_dev[l]) return fail(PVAE_ERR_INVALID, "net %d layer %d: null parameter pointer", net_id, l);
    net.W[l] = W_dev[l]; net.b[l] = b_dev[l];
  }
  net.grad = grad_flat_dev;
  net.bound = true;
  return PVAE_OK;
}

int pvae_bind_act_params(pvae_handle h, int net_id, const float* beta_dev, float* dbeta_dev) {
  if (!h || net_id < 0 || net_id >= PVAE_NUM_NETS) return fail(PVAE_ERR_INVALID, "bad handle / net id");
  h->nets[net_id].beta = beta_dev;
  h->nets[net_id].dbeta = beta_dev ? dbeta_dev : nullptr;
  return PVAE_OK;
}

int64_t pvae_net_grad_elems(pvae_handle h, int net_id) {
  if (!h || net_id < 0 || net_id >= PVAE_NUM_NETS) return -1;
  return h->nets[net_id].grad_elems;
}

int pvae_sync_weights(pvae_handle h, uint32_t net_mask, pvae_stream s) {
  if (!h) return fail(PVAE_ERR_INVALID, "null handle");
  cudaStream_t st = (cudaStream_t)s;
  for (int n = 0; n < PVAE_NUM_NETS; ++n) {
    if (!(net_mask & (1u << n))) continue;
    Net& net = h->nets[n];
    if (net.n_layers == 0) continue;
    if (!net.bound) return fail(PVAE_ERR_STATE, "net %d not bound", n);
    for (int l = 0; l < net.n_layers; ++l) {
      const int k0 = l == 0 ? net.k0 : net.in_dims[l];
      const int k1 = l == 0 ? net.k1 : 0;
      const int K0pad = l == 0 && net.k1 ? net.K0pad : net.kpad[l];
      const int64_t total = (int64_t)net.out_dims[l] * net.kpad[l];
      sync_weights_kernel<<<grid_for(total, 256, h->dev.sms), 256, 0, st>>>(net.W[l], net.out_dims[l], net.in_dims[l], k0, k1, K0pad,
                                                                            net.kpad[l], net.Wsh[l], net.wsh_ps[l], h->planes);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
  }
  CK(cudaGetLastError());
  return PVAE_OK;
}

int pvae_adam_step(pvae_handle h, int net_id, uint32_t layer_mask, float* exp_avg_dev, float* exp_avg_sq_dev, float* step_dev,
                   float lr, float beta1, float beta2, float eps, float weight_decay, pvae_stream s) {
  if (!h || net_id < 0 || net_id >= PVAE_NUM_NETS) return fail(PVAE_ERR_INVALID, "bad handle / net id");
  if (!exp_avg_dev || !exp_avg_sq_dev || !step_dev) return fail(PVAE_ERR_INVALID, "null optimizer state");
  Net& net = h->nets[net_id];
  if (!net.bound || !net.grad) return fail(PVAE_ERR_STATE, "net %d has no parameters / gradient buffer bound", net_id);
  cudaStream_t st = (cudaStream_t)s;
  {
    // one launch over the whole flat buffer when the layers are laid out back to back ([W0 | b0 | W1 | b1 | ...], which is what
    // physicsvae_b200 hands out) and 16-byte aligned; otherwise one launch per layer below
    bool flat = net.n_layers <= 8 && h->adam_counter != nullptr;
    for (int l = 0; l < net.n_layers && flat; ++l) {
      if (net.b[l] != net.W[l] + (int64_t)net.in_dims[l] * net.out_dims[l]) flat = false;
      if (l + 1 < net.n_layers && net.W[l + 1] != net.b[l] + net.out_dims[l]) flat = false;
    }
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (flat && al16(net.W[0]) && al16(net.grad) && al16(exp_avg_dev) && al16(exp_avg_sq_dev)) {
      AdamNet an;
      memset(&an, 0, sizeof(an));
      an.n_layers = net.n_layers;
      an.total = net.grad_elems;
      for (int l = 0; l < net.n_layers; ++l) {
        AdamLayer& L = an.L[l];
        L.off = net.gW[l]; L.out = net.out_dims[l]; L.in = net.in_dims[l];
        L.k0 = l == 0 ? net.k0 : net.in_dims[l];
        L.K0pad = l == 0 && net.k1 ? net.K0pad : net.kpad[l];
        L.Kpad = net.kpad[l]; L.active = (layer_mask >> l) & 1u;
        L.Wsh = net.Wsh[l]; L.ps = net.wsh_ps[l];
      }
      adam_net_kernel<<<grid_for((an.total + 3) / 4, 256, h->dev.sms), 256, 0, st>>>(
          const_cast<float*>(net.W[0]), net.grad, exp_avg_dev, exp_avg_sq_dev, step_dev, h->adam_counter, lr, beta1, beta2, eps,
          weight_decay, an, h->planes);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      CK(cudaGetLastError());
      return PVAE_OK;
    }
  }
  add_scalar_kernel<<<1, 32, 0, st>>>(step_dev, 1.f);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  for (int l = 0; l < net.n_layers; ++l) {
    if (!(layer_mask & (1u << l))) continue;
    // the flat layout requires b[l] to follow W[l] directly (true for the views physicsvae_b200 hands out)
    if (net.b[l] != net.W[l] + (int64_t)net.in_dims[l] * net.out_dims[l])
      return fail(PVAE_ERR_INVALID, "net %d layer %d: bias does not follow the weight in memory (flat [W|b] layout required)", net_id, l);
    const int k0 = l == 0 ? net.k0 : net.in_dims[l];
    const int K0pad = l == 0 && net.k1 ? net.K0pad : net.kpad[l];
    const int64_t total = (int64_t)net.out_dims[l] * net.in_dims[l] + net.out_dims[l];
    adam_layer_kernel<<<grid_for(total, 256, h->dev.sms), 256, 0, st>>>(
        const_cast<float*>(net.W[l]), net.grad + net.gW[l], exp_avg_dev + net.gW[l], exp_avg_sq_dev + net.gW[l], step_dev, lr, beta1,
        beta2, eps, weight_decay, net.out_dims[l], net.in_dims[l], k0, K0pad, net.kpad[l], net.Wsh[l], net.wsh_ps[l], h->planes);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  CK(cudaGetLastError());
  return PVAE_OK;
}

int pvae_workspace_bytes(pvae_handle h, size_t* bytes) {
  if (!h || !bytes) return fail(PVAE_ERR_INVALID, "null argument");
  *bytes = h->ws_bytes;
  return PVAE_OK;
}

int pvae_bind_workspace(pvae_handle h, void* ws_dev, size_t bytes) {
  if (!h || !ws_dev) return fail(PVAE_ERR_INVALID, "null argument");
  if (bytes < h->ws_bytes) return fail(PVAE_ERR_INVALID, "workspace too small: %zu < %zu", bytes, h->ws_bytes);
  if ((reinterpret_cast<uintptr_t>(ws_dev) & 1023) != 0) return fail(PVAE_ERR_INVALID, "workspace must be 1024-byte aligned");
  h->ws = ws_dev;
  carve(h, reinterpret_cast<uint8_t*>(ws_dev));
  return PVAE_OK;
}

int pvae_transitions_bytes(pvae_handle h, int64_t n_rows, size_t* bytes) {
  if (!h || !bytes || n_rows <= 0) return fail(PVAE_ERR_INVALID, "bad argument");
  *bytes = (size_t)h->planes * n_rows * (h->tx_ld + h->ty_ld) * 2;
  return PVAE_OK;
}

int pvae_ingest(pvae_handle h, void* buf_dev, int64_t buf_rows, int64_t dst_row, const void* x_raw_dev, int x_is_f64,
                const float* y_raw_dev, int64_t n_rows, pvae_stream s) {
  if (!h || !buf_dev || !x_raw_dev || !y_raw_dev) return fail(PVAE_ERR_INVALID, "null argument");
  if (n_rows <= 0 || dst_row < 0 || dst_row + n_rows > buf_rows) return fail(PVAE_ERR_INVALID, "row range [%lld, %lld) outside the buffer of %lld rows", (long long)dst_row, (long long)(dst_row + n_rows), (long long)buf_rows);
  if ((reinterpret_cast<uintptr_t>(buf_dev) & 15) != 0) return fail(PVAE_ERR_INVALID, "transition buffer must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)s;
  __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(buf_dev);
  __nv_bfloat16* yb = xb + (int64_t)h->planes * buf_rows * h->tx_ld;
  const int64_t xt = n_rows * (h->tx_ld / 2), yt = n_rows * (h->ty_ld / 2);
  if (x_is_f64)
    ingest_kernel<double><<<grid_for(xt, 256, h->dev.sms), 256, 0, st>>>(reinterpret_cast<const double*>(x_raw_dev), 2 * h->dsb, 2 * h->dsb, h->dsb, h->dsbp,
                                                                         y_raw_dev, h->da, h->da, h->dsb8,
                                                                         xb + dst_row * h->tx_ld, h->tx_ld, buf_rows * h->tx_ld, h->planes, n_rows);
  else
    ingest_kernel<float><<<grid_for(xt, 256, h->dev.sms), 256, 0, st>>>(reinterpret_cast<const float*>(x_raw_dev), 2 * h->dsb, 2 * h->dsb, h->dsb, h->dsbp,
                                                                        y_raw_dev, h->da, h->da, h->dsb8,
                                                                        xb + dst_row * h->tx_ld, h->tx_ld, buf_rows * h->tx_ld, h->planes, n_rows);
  ingest_kernel<float><<<grid_for(yt, 256, h->dev.sms), 256, 0, st>>>(y_raw_dev, h->da, h->da, h->da, 1 << 30, nullptr, 0, 0, 0, yb + dst_row * h->ty_ld, h->ty_ld,
                                                                      buf_rows * h->ty_ld, h->planes, n_rows);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return PVAE_OK;
}

int pvae_ingest_episodes(pvae_handle h, void* buf_dev, int64_t buf_rows, int64_t dst_row, const void* states_dev, int states_is_f64,
                         int64_t n_states, const void* actions_void, const int64_t* first_state_dev, int64_t n_rows, pvae_stream s) {
  const float* actions_dev = reinterpret_cast<const float*>(actions_void);
  if (!h || !buf_dev || !states_dev || !actions_dev || !first_state_dev) return fail(PVAE_ERR_INVALID, "null argument");
  if (n_rows <= 0 || dst_row < 0 || dst_row + n_rows > buf_rows) return fail(PVAE_ERR_INVALID, "row range [%lld, %lld) outside the buffer of %lld rows", (long long)dst_row, (long long)(dst_row + n_rows), (long long)buf_rows);
  if (n_states < 2) return fail(PVAE_ERR_INVALID, "an episode needs at least two states");
  if (states_is_f64 == 2 && h->planes > 1) return fail(PVAE_ERR_INVALID, "bf16 episode arrays carry one precision plane: use them with PVAE_PREC_BF16 only");
  if ((reinterpret_cast<uintptr_t>(buf_dev) & 15) != 0) return fail(PVAE_ERR_INVALID, "transition buffer must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)s;
  __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(buf_dev);
  __nv_bfloat16* yb = xb + (int64_t)h->planes * buf_rows * h->tx_ld;
  const int64_t xt = n_rows * (h->tx_ld / 2), yt = n_rows * (h->ty_ld / 2);
  if (states_is_f64 == 2) {
    // states AND actions already in bf16 (a loader that keeps the dataset in the engine's operand precision: half the upload of fp32)
    const __nv_bfloat16* sb = reinterpret_cast<const __nv_bfloat16*>(states_dev);
    const __nv_bfloat16* ab = reinterpret_cast<const __nv_bfloat16*>(actions_void);
    ingest_episodes_kernel<__nv_bfloat16, __nv_bfloat16><<<grid_for(xt, 256, h->dev.sms), 256, 0, st>>>(sb, ab, first_state_dev,
        h->dsb, h->da, h->dsb8, h->dsbp, xb + dst_row * h->tx_ld, h->tx_ld, buf_rows * h->tx_ld, h->planes, n_rows);
    ingest_episodes_kernel<__nv_bfloat16, __nv_bfloat16><<<grid_for(yt, 256, h->dev.sms), 256, 0, st>>>(sb, ab, first_state_dev,
        h->dsb, h->da, 0, -1, yb + dst_row * h->ty_ld, h->ty_ld, buf_rows * h->ty_ld, h->planes, n_rows);
  } else {
  if (states_is_f64)
    ingest_episodes_kernel<double><<<grid_for(xt, 256, h->dev.sms), 256, 0, st>>>(reinterpret_cast<const double*>(states_dev), actions_dev, first_state_dev,
        h->dsb, h->da, h->dsb8, h->dsbp, xb + dst_row * h->tx_ld, h->tx_ld, buf_rows * h->tx_ld, h->planes, n_rows);
  else
    ingest_episodes_kernel<float><<<grid_for(xt, 256, h->dev.sms), 256, 0, st>>>(reinterpret_cast<const float*>(states_dev), actions_dev, first_state_dev,
        h->dsb, h->da, h->dsb8, h->dsbp, xb + dst_row * h->tx_ld, h->tx_ld, buf_rows * h->tx_ld, h->planes, n_rows);
  ingest_episodes_kernel<float><<<grid_for(yt, 256, h->dev.sms), 256, 0, st>>>(reinterpret_cast<const float*>(states_dev), actions_dev, first_state_dev,
      h->dsb, h->da, 0, -1, yb + dst_row * h->ty_ld, h->ty_ld, buf_rows * h->ty_ld, h->planes, n_rows);
  }
  g_launches.fetch_add(2, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return PVAE_OK;
}

int pvae_bind_transitions(pvae_handle h, const void* buf_dev, int64_t buf_rows) {
  if (!h || !buf_dev || buf_rows <= 0) return fail(PVAE_ERR_INVALID, "bad argument");
  if (buf_rows > 0x7fffffffLL) return fail(PVAE_ERR_INVALID, "at most 2^31-1 rows per transition buffer");
  h->tbuf = reinterpret_cast<const __nv_bfloat16*>(buf_dev);
  h->tbuf_rows = buf_rows;
  return PVAE_OK;
}

int pvae_set_cursor(pvae_handle h, int64_t row, pvae_stream s) {
  if (!h || row < 0 || row >= (h->tbuf_rows ? h->tbuf_rows : 1)) return fail(PVAE_ERR_INVALID, "cursor row out of range");
  set_cursor_kernel<<<1, 32, 0, (cudaStream_t)s>>>(h->dev.cursor, (int32_t)row);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return PVAE_OK;
}

int pvae_advance_cursor(pvae_handle h, int64_t delta, int64_t batch, int64_t limit, pvae_stream s) {
  if (!h) return fail(PVAE_ERR_INVALID, "null handle");
  advance_cursor_kernel<<<1, 32, 0, (cudaStream_t)s>>>(h->dev.cursor, (int32_t)delta, (int32_t)batch, (int32_t)limit);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return PVAE_OK;
}

// ordered column sums of a bf16 gradient tensor [batch][cols] accumulated into colsum (deterministic mode)
static int det_colsum(pvae_engine* h, const __nv_bfloat16* g, int ld, int batch, int cols, float* colsum, cudaStream_t st) {
  if (!h->dev.det_scratch || (int64_t)DET_CHUNKS * cols > h->dev.det_scratch_elems) return fail(PVAE_ERR_INVALID, "layer too wide for the deterministic column sums");
  colsum_det_partial_kernel<<<dim3(cdiv(cols, 32), DET_CHUNKS), 256, 0, st>>>(g, ld, plane_elems(h, ld), h->planes, batch, cols, h->dev.det_scratch);
  colsum_det_final_kernel<<<cdiv(cols, 128), 128, 0, st>>>(h->dev.det_scratch, cols, colsum);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return PVAE_OK;
}

static int step_prologue(pvae_handle h, int batch) {
  if (!h) return fail(PVAE_ERR_INVALID, "null handle");
  if (!h->ws) return fail(PVAE_ERR_STATE, "no workspace bound (pvae_bind_workspace)");
  if (batch <= 0 || batch > h->max_batch) return fail(PVAE_ERR_INVALID, "batch %d outside [1, %d]", batch, h->max_batch);
  return PVAE_OK;
}

static int world_impl(pvae_handle h, int batch, float s_coeff, float* loss_dev, pvae_stream s, bool backward) {
  CKR(step_prologue(h, batch));
  if (!h->tbuf) return fail(PVAE_ERR_STATE, "no transition buffer bound (pvae_bind_transitions)");
  if (!loss_dev) return fail(PVAE_ERR_INVALID, "null loss pointer");
  cudaStream_t st = (cudaStream_t)s;
  Net& wm = h->nets[PVAE_NET_WORLD_MODEL];
  if (wm.n_layers == 0) return fail(PVAE_ERR_INVALID, "model has no world model");
  if (wm.generic) return fail(PVAE_ERR_INVALID, "a net with explicit input widths runs through pvae_fc_forward only");
  CKR(check_trainable_acts(wm));
  if (backward) {
    if (!wm.grad) return fail(PVAE_ERR_STATE, "world model has no gradient buffer bound");
    CK(cudaMemsetAsync(wm.grad, 0, wm.grad_elems * sizeof(float), st));
    if (wm.dbeta) CK(cudaMemsetAsync(wm.dbeta, 0, PVAE_MAX_LAYERS * sizeof(float), st));
  }
  if (h->acc_dirty) CK(cudaMemsetAsync(h->acc, 0, 4 * sizeof(double), st));   // only after a step that failed half-way: finalize clears them
  h->acc_dirty = true;
  NetIO in;                       // cat[s_t, a_t] (train_physics_vae.py:412-413) = the first dsb8 + da columns of a resident row
  in.nseg = 1;
  in.seg[0] = tx_view(h, 0, h->dsb8 + h->da);
  const int L = wm.n_layers;
  EpiParams last;
  memset(&last, 0, sizeof(last));
  last.type = EPI_MSE;
  set_aux(last, tx_view(h, h->dsbp, h->dsb), 0);
  last.scale = 2.f * s_coeff / ((float)batch * (float)h->dsb);
  set_out(last, h, wm.g[L - 1], wm.act_ld[L - 1]);
  last.colsum = backward ? wm.grad + wm.gb[L - 1] : nullptr;
  last.loss = h->acc + 2;
  CKR(net_forward(h, wm, in, batch, last, st, backward));
  if (backward) {
    // (data parallel: the gradients of layers 1 .. L-1 are exchanged while dgrad L0 / wgrad L0 run, see pvae_set_exchange)
    const int r = net_backward(h, wm, in, batch, true, nullptr, st, nullptr, wm.n_layers >= 2 ? 1 : -1);
    const int rj = exchange_join(h, st);
    if (r != PVAE_OK) return r;
    CKR(rj);
  }
  finalize_loss_kernel<<<1, 32, 0, st>>>(h->acc, loss_dev, batch, h->da, h->dsb, 0.f, 0.f, s_coeff, 0.f, nullptr, 0ull);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  h->acc_dirty = false;
  return PVAE_OK;
}

int pvae_world_step(pvae_handle h, int batch, float s_coeff, float* loss_dev, pvae_stream s) {
  return world_impl(h, batch, s_coeff, loss_dev, s, true);
}

static int vae_impl(pvae_handle h, int batch, const float* eps_dev, uint64_t seed, uint64_t offset, int noise, float a_coeff,
                    float kl_coeff, float cyc_coeff, float* loss_dev, pvae_stream s, bool backward) {
  CKR(step_prologue(h, batch));
  if (!h->tbuf) return fail(PVAE_ERR_STATE, "no transition buffer bound (pvae_bind_transitions)");
  if (!loss_dev) return fail(PVAE_ERR_INVALID, "null loss pointer");
  cudaStream_t st = (cudaStream_t)s;
  Net& te = h->nets[PVAE_NET_TASK_ENCODER];
  Net& md = h->nets[PVAE_NET_MOTOR_DECODER];
  Net& wm = h->nets[PVAE_NET_WORLD_MODEL];
  if (te.n_layers == 0 || md.n_layers == 0) return fail(PVAE_ERR_INVALID, "model has no encoder / decoder");
  if (te.generic || md.generic || wm.generic) return fail(PVAE_ERR_INVALID, "a net with explicit input widths runs through pvae_fc_forward only");
  const bool cyc = cyc_coeff != 0.f && wm.n_layers > 0;
  CKR(check_trainable_acts(te));
  CKR(check_trainable_acts(md));
  if (cyc) CKR(check_trainable_acts(wm));
  const int prior = h->desc.latent_prior;
  if (backward) {
    if (!te.grad || !md.grad) return fail(PVAE_ERR_STATE, "encoder / decoder have no gradient buffers bound");
    CK(cudaMemsetAsync(te.grad, 0, te.grad_elems * sizeof(float), st));
    CK(cudaMemsetAsync(md.grad, 0, md.grad_elems * sizeof(float), st));
    if (te.dbeta) CK(cudaMemsetAsync(te.dbeta, 0, PVAE_MAX_LAYERS * sizeof(float), st));
    if (md.dbeta) CK(cudaMemsetAsync(md.dbeta, 0, PVAE_MAX_LAYERS * sizeof(float), st));
  }
  const bool draws = prior && noise && !eps_dev;     // the step consumes the Philox stream
  if (h->acc_dirty) CK(cudaMemsetAsync(h->acc, 0, 4 * sizeof(double), st));   // only after a step that failed half-way: finalize clears them
  h->acc_dirty = true;
  const int z = h->z, Lte = te.n_layers, Lmd = md.n_layers, Lwm = wm.n_layers;

  // encoder: h = TE(cat[s1, s2]) -> (mu | logvar)                      rllib_model_torch.py:773-800
  NetIO te_in; te_in.nseg = 2; te_in.seg[0] = tx_view(h, 0, h->dsb); te_in.seg[1] = tx_view(h, h->dsbp, h->dsb);
  EpiParams e;
  memset(&e, 0, sizeof(e));
  e.type = EPI_STORE; e.out_f32 = h->ml; e.f32_sm = h->te_out; e.f32_sn = 1;
  CKR(net_forward(h, te, te_in, batch, e, st, backward));
  // z = mu + eps * exp(0.5 logvar), KL partial sums                     rllib_model_torch.py:734-740, train_physics_vae.py:384-389
  {
    const int64_t total = (int64_t)batch * z;
    reparam_fwd_kernel<<<grid_for(total, 256, h->dev.sms), 256, 0, st>>>(h->ml, eps_dev, h->eps, prior, prior && noise, seed, offset, batch, z,
                                                                         h->zb, h->zb_ld, plane_elems(h, h->zb_ld), h->planes, nullptr,
                                                                         nullptr, nullptr, (prior && kl_coeff != 0.f) ? h->acc + 1 : nullptr,
                                                                         (draws && h->noise_auto) ? h->noise_ctr : nullptr);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  // decoder: a_hat = MD(cat[s1, z]); action reconstruction loss          rllib_model_torch.py:822-837, train_physics_vae.py:381-382
  NetIO md_in; md_in.nseg = 2; md_in.seg[0] = tx_view(h, 0, h->dsb); md_in.seg[1] = ws_view(h, h->zb, h->zb_ld, z, batch);
  memset(&e, 0, sizeof(e));
  e.type = EPI_MSE;
  set_aux(e, ty_view(h, h->da), 0);
  e.scale = 2.f * a_coeff / ((float)batch * (float)h->da);
  e.loss = h->acc + 0;
  set_out2(e, h, h->ahat, h->a_ld);
  if (cyc) {
    set_out(e, h, h->ga, h->a_ld);
  } else {
    set_out(e, h, md.g[Lmd - 1], md.act_ld[Lmd - 1]);
    e.colsum = backward ? md.grad + md.gb[Lmd - 1] : nullptr;
  }
  CKR(net_forward(h, md, md_in, batch, e, st, backward));
  NetIO wm_in;
  if (cyc) {
    // frozen world model on the decoded action: cycle loss                rllib_model_torch.py:839-844, train_physics_vae.py:417-419
    if (!wm.bound) return fail(PVAE_ERR_STATE, "world model not bound");
    wm_in.nseg = 2; wm_in.seg[0] = tx_view(h, 0, h->dsb); wm_in.seg[1] = ws_view(h, h->ahat, h->a_ld, h->da, batch);
    memset(&e, 0, sizeof(e));
    e.type = EPI_MSE;
    set_aux(e, tx_view(h, h->dsbp, h->dsb), 0);
    e.scale = 2.f * cyc_coeff / ((float)batch * (float)h->dsb);
    e.loss = h->acc + 3;
    set_out(e, h, wm.g[Lwm - 1], wm.act_ld[Lwm - 1]);
    CKR(net_forward(h, wm, wm_in, batch, e, st, backward));
    // d/d a_hat through the frozen world model, plus the action-loss gradient -> decoder output gradient
    if (backward) {
    memset(&e, 0, sizeof(e));
    e.type = EPI_DGRAD; e.act = ACT_LINEAR;
    e.add = h->ga; e.add_ld = h->a_ld; e.add_ps = plane_elems(h, h->a_ld); e.add_planes = h->planes;
    set_out(e, h, md.g[Lmd - 1], md.act_ld[Lmd - 1]);
    e.colsum = md.grad + md.gb[Lmd - 1];
    CKR(net_backward(h, wm, wm_in, batch, false, &e, st));
    }
  }
  if (backward) {
  // decoder backward; gradient w.r.t. z lands in dz (fp32)
  memset(&e, 0, sizeof(e));
  e.type = EPI_DGRAD; e.act = ACT_LINEAR;
  e.out_f32 = h->dz; e.f32_sm = z; e.f32_sn = 1;
  CKR(net_backward(h, md, md_in, batch, true, &e, st));
  CKR(exchange_fork(h, st));      // (data parallel: the decoder's gradients are exchanged while the encoder's backward pass runs)
  // reparameterisation + KL backward -> gradient of the encoder's output layer
  {
    const int w = h->te_out;
    int rows_per_block = 256 / w; if (rows_per_block < 1) rows_per_block = 1;
    const int threads = w * rows_per_block;
    if (threads > 1024) return fail(PVAE_ERR_INVALID, "latent_dim %d too large for the reparameterisation kernel", z);
    int grid = cdiv(batch, rows_per_block); if (grid > h->dev.sms * 8) grid = h->dev.sms * 8;
    reparam_bwd_kernel<<<grid, threads, 0, st>>>(h->dz, h->ml, h->eps, prior, prior && noise, prior ? kl_coeff / (float)batch : 0.f, batch, z,
                                                 te.g[Lte - 1], te.act_ld[Lte - 1], plane_elems(h, te.act_ld[Lte - 1]), h->planes,
                                                 h->dev.deterministic ? nullptr : te.grad + te.gb[Lte - 1]);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (h->dev.deterministic) CKR(det_colsum(h, te.g[Lte - 1], te.act_ld[Lte - 1], batch, w, te.grad + te.gb[Lte - 1], st));
  }
  {
    const int r = net_backward(h, te, te_in, batch, true, nullptr, st);
    const int rj = exchange_join(h, st);
    if (r != PVAE_OK) return r;
    CKR(rj);
  }
  }
  finalize_loss_kernel<<<1, 32, 0, st>>>(h->acc, loss_dev, batch, h->da, h->dsb, a_coeff, prior ? kl_coeff : 0.f, 0.f, cyc ? cyc_coeff : 0.f,
                                         (draws && h->noise_auto) ? h->noise_ctr : nullptr, h->noise_stride);
  h->acc_dirty = false;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return PVAE_OK;
}

int pvae_vae_step(pvae_handle h, int batch, const float* eps_dev, uint64_t seed, uint64_t offset, int noise, float a_coeff,
                  float kl_coeff, float cyc_coeff, float* loss_dev, pvae_stream s) {
  return vae_impl(h, batch, eps_dev, seed, offset, noise, a_coeff, kl_coeff, cyc_coeff, loss_dev, s, true);
}

int pvae_eval_loss(pvae_handle h, int phase, int batch, const float* eps_dev, uint64_t seed, uint64_t offset, int noise, float a_coeff,
                   float kl_coeff, float s_coeff, float cyc_coeff, float* loss_dev, pvae_stream s) {
  if (phase == 0) return world_impl(h, batch, s_coeff, loss_dev, s, false);
  if (phase == 1) return vae_impl(h, batch, eps_dev, seed, offset, noise, a_coeff, kl_coeff, cyc_coeff, loss_dev, s, false);
  return fail(PVAE_ERR_INVALID, "phase must be 0 (world model) or 1 (VAE)");
}

// ---- autoregressive rollout: compute_loss with lookahead L > 1 (train_physics_vae.py:361-435) ------------------------------------------
// Step t runs the FULL model on (s1_t, s2_gt_t) where s1_0 is data and s1_{t+1} is the world model's prediction from the DECODED action
// of step t (`s1 = self.model._cur_future_state`, :421); the four loss terms are averaged over the steps (:423-428).  Gradients flow
// through time: every net is differentiated w.r.t. its body-state input segment as well, and those gradients are accumulated -- in
// place, by the dgrad epilogue's addend -- into the output gradient of the previous step's world-model pass.  Every step keeps its own
// copy of the workspace (activations, masks, mu | logvar, eps, z, decoded action, predicted state); the world phase needs a second
// world-model pass per step (on the ground-truth action, :412-414) which lives in slots L .. 2L-1.
int pvae_rollout_workspace_bytes(pvae_handle h, int lookahead, size_t* bytes) {
  if (!h || !bytes || lookahead < 1) return fail(PVAE_ERR_INVALID, "bad argument");
  *bytes = (size_t)2 * lookahead * h->ws_bytes;
  return PVAE_OK;
}

int pvae_bind_rollout_workspace(pvae_handle h, void* ws_dev, size_t bytes, int lookahead) {
  if (!h || !ws_dev || lookahead < 1) return fail(PVAE_ERR_INVALID, "bad argument");
  if (bytes < (size_t)2 * lookahead * h->ws_bytes) return fail(PVAE_ERR_INVALID, "rollout workspace too small: %zu < %zu", bytes, (size_t)2 * lookahead * h->ws_bytes);
  if ((reinterpret_cast<uintptr_t>(ws_dev) & 1023) != 0) return fail(PVAE_ERR_INVALID, "workspace must be 1024-byte aligned");
  h->roll_ws = reinterpret_cast<uint8_t*>(ws_dev);
  h->roll_slots = 2 * lookahead;
  return PVAE_OK;
}

static int rollout_impl(pvae_handle h, int phase, int batch, int L, const void* const* tbufs, int64_t buf_rows, const float* eps_dev,
                        uint64_t seed, uint64_t offset, int noise, float a_coeff, float kl_coeff, float s_coeff, float cyc_coeff,
                        float* loss_dev, cudaStream_t st) {
  Net& te = h->nets[PVAE_NET_TASK_ENCODER];
  Net& md = h->nets[PVAE_NET_MOTOR_DECODER];
  Net& wm = h->nets[PVAE_NET_WORLD_MODEL];
  const bool world = phase == 0;
  const bool train_wm = world, train_vae = !world;
  const int prior = h->desc.latent_prior;
  const int z = h->z, Lte = te.n_layers, Lmd = md.n_layers, Lwm = wm.n_layers;
  const float invL = 1.f / (float)L;
  if (world) { a_coeff = 0.f; kl_coeff = 0.f; cyc_coeff = 0.f; } else { s_coeff = 0.f; }
  if (train_wm) {
    CK(cudaMemsetAsync(wm.grad, 0, wm.grad_elems * sizeof(float), st));
    if (wm.dbeta) CK(cudaMemsetAsync(wm.dbeta, 0, PVAE_MAX_LAYERS * sizeof(float), st));
  }
  if (train_vae) {
    CK(cudaMemsetAsync(te.grad, 0, te.grad_elems * sizeof(float), st));
    CK(cudaMemsetAsync(md.grad, 0, md.grad_elems * sizeof(float), st));
    if (te.dbeta) CK(cudaMemsetAsync(te.dbeta, 0, PVAE_MAX_LAYERS * sizeof(float), st));
    if (md.dbeta) CK(cudaMemsetAsync(md.dbeta, 0, PVAE_MAX_LAYERS * sizeof(float), st));
  }
  if (h->acc_dirty) CK(cudaMemsetAsync(h->acc, 0, 4 * sizeof(double), st));
  h->acc_dirty = true;
  auto use_slot = [&](int s_) { carve(h, h->roll_ws + (size_t)s_ * h->ws_bytes); };
  auto bind_t = [&](int t) { h->tbuf = reinterpret_cast<const __nv_bfloat16*>(tbufs[t]); h->tbuf_rows = buf_rows; };
  // per-step pointers that another step needs: the predicted state (next step's body-state input) and the output gradient of the
  // cycle pass's world model (where the next step's body-state gradients are accumulated)
  std::vector<__nv_bfloat16*> fut(L), wm_glast(L);
  for (int t = 0; t < L; ++t) { use_slot(t); fut[t] = h->fut; wm_glast[t] = wm.g[Lwm - 1]; }
  auto s1_view = [&](int t) {        // body state of step t: data for t = 0, the previous step's prediction afterwards
    return t == 0 ? tx_view(h, 0, h->dsb) : ws_view(h, fut[t - 1], h->f_ld, h->dsb, batch);
  };
  EpiParams e;
  // ------------------------------------------------ forward, t = 0 .. L-1 ------------------------------------------------
  for (int t = 0; t < L; ++t) {
    use_slot(t);
    bind_t(t);
    NetIO te_in; te_in.nseg = 2; te_in.seg[0] = s1_view(t); te_in.seg[1] = tx_view(h, h->dsbp, h->dsb);
    memset(&e, 0, sizeof(e));
    e.type = EPI_STORE; e.out_f32 = h->ml; e.f32_sm = h->te_out; e.f32_sn = 1;
    CKR(net_forward(h, te, te_in, batch, e, st, true));
    {
      const int64_t total = (int64_t)batch * z;
      reparam_fwd_kernel<<<grid_for(total, 256, h->dev.sms), 256, 0, st>>>(h->ml, eps_dev ? eps_dev + (int64_t)t * batch * z : nullptr, h->eps, prior,
                                                                           prior && noise, seed, offset + (uint64_t)t, batch, z, h->zb, h->zb_ld,
                                                                           plane_elems(h, h->zb_ld), h->planes, nullptr, nullptr, nullptr,
                                                                           (prior && kl_coeff != 0.f && a_coeff > 0.f) ? h->acc + 1 : nullptr, nullptr);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    NetIO md_in; md_in.nseg = 2; md_in.seg[0] = s1_view(t); md_in.seg[1] = ws_view(h, h->zb, h->zb_ld, z, batch);
    memset(&e, 0, sizeof(e));
    e.type = EPI_MSE;
    set_aux(e, ty_view(h, h->da), 0);
    e.scale = 2.f * a_coeff * invL / ((float)batch * (float)h->da);
    e.loss = h->acc + 0;
    set_out2(e, h, h->ahat, h->a_ld);
    set_out(e, h, h->ga, h->a_ld);
    CKR(net_forward(h, md, md_in, batch, e, st, true));
    NetIO wm_in; wm_in.nseg = 2; wm_in.seg[0] = s1_view(t); wm_in.seg[1] = ws_view(h, h->ahat, h->a_ld, h->da, batch);
    memset(&e, 0, sizeof(e));
    e.type = EPI_MSE;
    set_aux(e, tx_view(h, h->dsbp, h->dsb), 0);
    e.scale = 2.f * cyc_coeff * invL / ((float)batch * (float)h->dsb);
    e.loss = h->acc + 3;
    set_out(e, h, wm.g[Lwm - 1], wm.act_ld[Lwm - 1]);
    set_out2(e, h, h->fut, h->f_ld);
    // (the bias gradient of this layer needs the gradient that later steps add to g: it is summed by the last accumulating launch
    //  of step t + 1, or right here for the last step)
    e.colsum = (train_wm && t == L - 1) ? wm.grad + wm.gb[Lwm - 1] : nullptr;
    CKR(net_forward(h, wm, wm_in, batch, e, st, true));
    if (world) {       // second world-model pass, on the ground-truth action: the world phase's loss term
      use_slot(L + t);
      NetIO gt_in; gt_in.nseg = 2; gt_in.seg[0] = s1_view(t); gt_in.seg[1] = tx_view(h, h->dsb8, h->da);
      memset(&e, 0, sizeof(e));
      e.type = EPI_MSE;
      set_aux(e, tx_view(h, h->dsbp, h->dsb), 0);
      e.scale = 2.f * s_coeff * invL / ((float)batch * (float)h->dsb);
      e.loss = h->acc + 2;
      set_out(e, h, wm.g[Lwm - 1], wm.act_ld[Lwm - 1]);
      e.colsum = wm.grad + wm.gb[Lwm - 1];
      CKR(net_forward(h, wm, gt_in, batch, e, st, true));
    }
  }
  // ------------------------------------------------ backward, t = L-1 .. 0 ------------------------------------------------
  for (int t = L - 1; t >= 0; --t) {
    bind_t(t);
    // gradient w.r.t. this step's body state accumulates into the previous step's cycle-pass output gradient
    EpiParams acc0;
    memset(&acc0, 0, sizeof(acc0));
    if (t > 0) {
      acc0.type = EPI_DGRAD; acc0.act = ACT_LINEAR;
      acc0.add = wm_glast[t - 1]; acc0.add_ld = h->f_ld; acc0.add_ps = plane_elems(h, h->f_ld); acc0.add_planes = h->planes;
      acc0.out = wm_glast[t - 1]; acc0.out_ld = h->f_ld; acc0.out_ps = plane_elems(h, h->f_ld); acc0.out_planes = h->planes;
    }
    const EpiParams* p0 = t > 0 ? &acc0 : nullptr;
    if (world) {
      use_slot(L + t);
      NetIO gt_in; gt_in.nseg = 2; gt_in.seg[0] = s1_view(t); gt_in.seg[1] = tx_view(h, h->dsb8, h->da);
      CKR(net_backward(h, wm, gt_in, batch, true, nullptr, st, p0));
    }
    use_slot(t);
    NetIO te_in; te_in.nseg = 2; te_in.seg[0] = s1_view(t); te_in.seg[1] = tx_view(h, h->dsbp, h->dsb);
    NetIO md_in; md_in.nseg = 2; md_in.seg[0] = s1_view(t); md_in.seg[1] = ws_view(h, h->zb, h->zb_ld, z, batch);
    NetIO wm_in; wm_in.nseg = 2; wm_in.seg[0] = s1_view(t); wm_in.seg[1] = ws_view(h, h->ahat, h->a_ld, h->da, batch);
    // cycle pass of the world model: d a_hat (+ the action-loss gradient) -> decoder output gradient
    memset(&e, 0, sizeof(e));
    e.type = EPI_DGRAD; e.act = ACT_LINEAR;
    e.add = h->ga; e.add_ld = h->a_ld; e.add_ps = plane_elems(h, h->a_ld); e.add_planes = h->planes;
    set_out(e, h, md.g[Lmd - 1], md.act_ld[Lmd - 1]);
    e.colsum = train_vae ? md.grad + md.gb[Lmd - 1] : nullptr;
    CKR(net_backward(h, wm, wm_in, batch, train_wm, &e, st, p0));
    // decoder
    memset(&e, 0, sizeof(e));
    e.type = EPI_DGRAD; e.act = ACT_LINEAR;
    e.out_f32 = h->dz; e.f32_sm = z; e.f32_sn = 1;
    CKR(net_backward(h, md, md_in, batch, train_vae, &e, st, p0));
    {
      const int w = h->te_out;
      int rows_per_block = 256 / w; if (rows_per_block < 1) rows_per_block = 1;
      const int threads = w * rows_per_block;
      if (threads > 1024) return fail(PVAE_ERR_INVALID, "latent_dim %d too large for the reparameterisation kernel", z);
      int grid = cdiv(batch, rows_per_block); if (grid > h->dev.sms * 8) grid = h->dev.sms * 8;
      const float kls = (prior && a_coeff > 0.f) ? kl_coeff * invL / (float)batch : 0.f;
      reparam_bwd_kernel<<<grid, threads, 0, st>>>(h->dz, h->ml, h->eps, prior, prior && noise, kls, batch, z, te.g[Lte - 1], te.act_ld[Lte - 1],
                                                   plane_elems(h, te.act_ld[Lte - 1]), h->planes,
                                                   (train_vae && !h->dev.deterministic) ? te.grad + te.gb[Lte - 1] : nullptr);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (train_vae && h->dev.deterministic) CKR(det_colsum(h, te.g[Lte - 1], te.act_ld[Lte - 1], batch, w, te.grad + te.gb[Lte - 1], st));
    }
    // encoder; its body-state gradient is the last contribution to the previous step's output gradient: that launch also sums the
    // world model's output-layer bias gradient of step t - 1
    EpiParams last0 = acc0;
    if (t > 0 && train_wm) last0.colsum = wm.grad + wm.gb[Lwm - 1];
    CKR(net_backward(h, te, te_in, batch, train_vae, nullptr, st, t > 0 ? &last0 : nullptr));
  }
  finalize_loss_kernel<<<1, 32, 0, st>>>(h->acc, loss_dev, batch, h->da, h->dsb, a_coeff, (prior && a_coeff > 0.f) ? kl_coeff : 0.f, s_coeff, cyc_coeff,
                                         nullptr, 0ull, L);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  h->acc_dirty = false;
  return PVAE_OK;
}

int pvae_rollout_step(pvae_handle h, int phase, int batch, int lookahead, const void* const* tbufs_host, int64_t buf_rows, const float* eps_dev,
                      uint64_t seed, uint64_t offset, int noise, float a_coeff, float kl_coeff, float s_coeff, float cyc_coeff, float* loss_dev,
                      pvae_stream s) {
  CKR(step_prologue(h, batch));
  if (phase != 0 && phase != 1) return fail(PVAE_ERR_INVALID, "phase must be 0 (world model) or 1 (VAE)");
  if (lookahead < 1 || !tbufs_host || !loss_dev) return fail(PVAE_ERR_INVALID, "bad argument");
  if (!h->roll_ws || h->roll_slots < 2 * lookahead) return fail(PVAE_ERR_STATE, "no rollout workspace bound for lookahead %d (pvae_bind_rollout_workspace)", lookahead);
  if (buf_rows <= 0 || buf_rows > 0x7fffffffLL) return fail(PVAE_ERR_INVALID, "bad buffer row count");
  for (int t = 0; t < lookahead; ++t) if (!tbufs_host[t]) return fail(PVAE_ERR_INVALID, "null transition buffer for step %d", t);
  Net& te = h->nets[PVAE_NET_TASK_ENCODER];
  Net& md = h->nets[PVAE_NET_MOTOR_DECODER];
  Net& wm = h->nets[PVAE_NET_WORLD_MODEL];
  if (te.n_layers == 0 || md.n_layers == 0 || wm.n_layers == 0) return fail(PVAE_ERR_INVALID, "a rollout needs encoder, decoder and world model");
  if (te.generic || md.generic || wm.generic) return fail(PVAE_ERR_INVALID, "a net with explicit input widths runs through pvae_fc_forward only");
  CKR(check_trainable_acts(te));
  CKR(check_trainable_acts(md));
  CKR(check_trainable_acts(wm));
  if (!te.bound || !md.bound || !wm.bound) return fail(PVAE_ERR_STATE, "nets not bound");
  if ((phase == 0 && !wm.grad) || (phase == 1 && (!te.grad || !md.grad))) return fail(PVAE_ERR_STATE, "no gradient buffers bound for the trained nets");
  const __nv_bfloat16* saved_tbuf = h->tbuf;
  const int64_t saved_rows = h->tbuf_rows;
  const int r = rollout_impl(h, phase, batch, lookahead, tbufs_host, buf_rows, eps_dev, seed, offset, noise, a_coeff, kl_coeff, s_coeff, cyc_coeff,
                             loss_dev, (cudaStream_t)s);
  if (h->ws) carve(h, reinterpret_cast<uint8_t*>(h->ws));      // back to the single-step workspace
  h->tbuf = saved_tbuf; h->tbuf_rows = saved_rows;
  return r;
}

int pvae_set_exchange(pvae_handle h, const uint64_t* peer_ptrs_host, uint64_t multicast_ptr, int rank, int world, int64_t offset_elems,
                      int64_t count_elems, int64_t flags_offset_elems, int ctas) {
  if (!h) return fail(PVAE_ERR_INVALID, "null handle");
  h->xchg.on = false;
  h->xchg.forked = false;
  h->dev.reserved_sms = 0;
  if (!peer_ptrs_host || world < 2) return PVAE_OK;                 // switched off
  if (world > AR_MAX_RANKS || rank < 0 || rank >= world) return fail(PVAE_ERR_INVALID, "rank %d / world %d outside [2, %d]", rank, world, AR_MAX_RANKS);
  if (offset_elems < 0 || count_elems <= 0 || (offset_elems & 3) || (count_elems & 3) || (flags_offset_elems & 3))
    return fail(PVAE_ERR_INVALID, "range / flag block must be multiples of 4 elements");
  if (ctas < 1 || ctas > AR_CTAS || ctas >= h->dev.sms / 2) return fail(PVAE_ERR_INVALID, "exchange CTAs %d outside [1, %d]", ctas, AR_CTAS);
  if (!h->xchg.side) {
    CK(cudaStreamCreateWithFlags(&h->xchg.side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->xchg.ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->xchg.ev_join, cudaEventDisableTiming));
  }
  SymmArgs& a = h->xchg.args;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < world; ++p) {
    if (!peer_ptrs_host[p] || (peer_ptrs_host[p] & 15)) return fail(PVAE_ERR_INVALID, "peer pointer %d is null or not 16-byte aligned", p);
    a.peer[p] = reinterpret_cast<float*>(peer_ptrs_host[p]);
  }
  a.mc = reinterpret_cast<float*>(multicast_ptr);
  a.rank = rank; a.world = world; a.off = offset_elems; a.count = count_elems; a.flags_off = flags_offset_elems;
  a.scale = 1.f / (float)world;
  h->xchg.ctas = ctas;
  h->xchg.on = true;
  return PVAE_OK;
}

int pvae_run_exchange(pvae_handle h, pvae_stream s) {
  if (!h) return fail(PVAE_ERR_INVALID, "null handle");
  if (!h->xchg.on) return PVAE_OK;
  return launch_symm(h->xchg.args, h->xchg.ctas, (cudaStream_t)s);
}

int pvae_set_deterministic(pvae_handle h, int enable) {
  if (!h) return fail(PVAE_ERR_INVALID, "null handle");
  if (enable && !h->small_scratch) return fail(PVAE_ERR_CUDA, "scratch buffer was not allocated");
  h->dev.deterministic = enable != 0;
  h->dev.det_scratch = h->small_scratch;
  h->dev.det_scratch_elems = 2 * SF_MAX_ROWS * SMALL_SCRATCH_COLS;
  return PVAE_OK;
}

int pvae_noise_counter(pvae_handle h, int enable, uint64_t value, uint64_t stride, pvae_stream s) {
  if (!h) return fail(PVAE_ERR_INVALID, "null handle");
  if (!h->noise_ctr) return fail(PVAE_ERR_CUDA, "noise counter was not allocated");
  h->noise_auto = enable != 0;
  h->noise_stride = stride ? stride : 1;
  set_u64_kernel<<<1, 32, 0, (cudaStream_t)s>>>(h->noise_ctr, (unsigned long long)value);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return PVAE_OK;
}

// FC.forward (rllib_model_torch.py:274-275) of one net on fp32 rows handed in by the caller.
int pvae_fc_forward(pvae_handle h, int net_id, int batch, const float* in_dev, int64_t in_ld, float* out_dev, int64_t out_ld, pvae_stream s) {
  CKR(step_prologue(h, batch));
  if (net_id < 0 || net_id >= PVAE_NUM_NETS) return fail(PVAE_ERR_INVALID, "bad net id");
  Net& net = h->nets[net_id];
  if (net.n_layers == 0) return fail(PVAE_ERR_INVALID, "net %d is absent from the model description", net_id);
  if (!in_dev || !out_dev) return fail(PVAE_ERR_INVALID, "null argument");
  const int out_w = net.out_dims[net.n_layers - 1];
  if (in_ld < net.in_dim || out_ld < out_w) return fail(PVAE_ERR_INVALID, "row strides %lld / %lld too small for %d -> %d", (long long)in_ld, (long long)out_ld, net.in_dim, out_w);
  cudaStream_t st = (cudaStream_t)s;
  if (small_ok(h, net, batch))                // a handful of rows: one cluster kernel for the whole stack
    return small_chain(h, net, batch, in_dev, in_ld, net.k1 ? in_dev + net.k0 : nullptr, in_ld, out_dev, out_ld, st);
  const int p0 = rup(net.k0, 64);            // second input segment starts on a 128-byte boundary of the staging row
  {
    const int64_t total = (int64_t)batch * h->x_ld;
    f32_to_planes_kernel<<<grid_for(total, 256, h->dev.sms), 256, 0, st>>>(in_dev, in_ld, net.in_dim, net.k0, p0, h->xin, h->x_ld, plane_elems(h, h->x_ld), h->planes, batch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  NetIO in;
  in.nseg = net.k1 ? 2 : 1;
  in.seg[0] = ws_view(h, h->xin, h->x_ld, net.k0, batch);
  if (net.k1) in.seg[1] = ws_view(h, h->xin + p0, h->x_ld, net.k1, batch);
  EpiParams e;
  memset(&e, 0, sizeof(e));
  e.type = EPI_STORE; e.out_f32 = out_dev; e.f32_sm = out_ld; e.f32_sn = 1;
  CKR(net_forward(h, net, in, batch, e, st));
  CK(cudaGetLastError());
  return PVAE_OK;
}

int pvae_forward(pvae_handle h, uint32_t parts, int batch, const float* obs_dev, int64_t obs_ld, const float* z_in_dev,
                 const float* act_in_dev, int64_t act_in_ld, const float* eps_dev, int noise, uint64_t seed,
                 uint64_t offset, float* act_out_dev, int64_t act_out_ld, float* mu_dev, float* logvar_dev, float* z_dev,
                 float* future_dev, float* value_dev, pvae_stream s) {
  CKR(step_prologue(h, batch));
  if (!obs_dev) return fail(PVAE_ERR_INVALID, "obs_dev is required (the body state feeds the decoder and the world model)");
  for (int n = 0; n < PVAE_NUM_NETS; ++n)
    if (h->nets[n].generic) return fail(PVAE_ERR_INVALID, "a net with explicit input widths runs through pvae_fc_forward only");
  cudaStream_t st = (cudaStream_t)s;
  const int z = h->z;
  const bool enc = parts & PVAE_PART_ENCODER, dec = parts & PVAE_PART_DECODER, wld = parts & PVAE_PART_WORLD, val = parts & PVAE_PART_VALUE;
  const int obs_w = (enc || val) ? 2 * h->dsb : h->dsb;
  if (obs_ld < obs_w) return fail(PVAE_ERR_INVALID, "obs row stride %lld < %d", (long long)obs_ld, obs_w);
  {
    // ---- latency path (batch <= 16): one cluster kernel per FC stack straight from / to the caller's fp32 rows -- at most four
    //      chain launches + the reparameterisation kernel for a full forward, ONE launch for the runtime's decoder pass-through
    bool small = h->small_scratch != nullptr && h->z <= SMALL_SCRATCH_COLS && h->da <= SMALL_SCRATCH_COLS;
    for (int n = 0; n < PVAE_NUM_NETS && small; ++n) {
      const bool used = (n == PVAE_NET_TASK_ENCODER && enc) || (n == PVAE_NET_MOTOR_DECODER && dec) || (n == PVAE_NET_WORLD_MODEL && wld) ||
                        (n == PVAE_NET_VALUE_BRANCH && val);
      if (used && (h->nets[n].n_layers == 0 || !small_ok(h, h->nets[n], batch))) small = false;
    }
    if (small) {
      float* z_rows = h->small_scratch;                                          // [16][z]
      float* a_rows = h->small_scratch + SF_MAX_ROWS * SMALL_SCRATCH_COLS;      // [16][da]
      const float* z_src = z_in_dev;
      int64_t z_ld = z;
      if (enc) {
        Net& te = h->nets[PVAE_NET_TASK_ENCODER];
        CKR(small_chain(h, te, batch, obs_dev, obs_ld, obs_dev + h->dsb, obs_ld, h->ml, h->te_out, st));
        const int prior = h->desc.latent_prior;
        float* zf = z_dev ? z_dev : z_rows;
        reparam_fwd_kernel<<<grid_for((int64_t)batch * z, 256, h->dev.sms), 256, 0, st>>>(h->ml, eps_dev, h->eps, prior, prior && noise, seed, offset, batch, z,
                                                                                          h->zb, h->zb_ld, plane_elems(h, h->zb_ld), h->planes, zf, mu_dev,
                                                                                          prior ? logvar_dev : nullptr, nullptr, nullptr);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        z_src = zf;
      } else if (dec && !z_in_dev) {
        return fail(PVAE_ERR_INVALID, "decoder without encoder needs z_in_dev");
      }
      const float* a_src = act_in_dev;
      int64_t a_ld = act_in_ld;
      if (dec) {
        float* af = act_out_dev ? act_out_dev : a_rows;
        const int64_t af_ld = act_out_dev ? act_out_ld : h->da;
        CKR(small_chain(h, h->nets[PVAE_NET_MOTOR_DECODER], batch, obs_dev, obs_ld, z_src, z_ld, af, af_ld, st));
        a_src = af; a_ld = af_ld;
      } else if (wld && !act_in_dev) {
        return fail(PVAE_ERR_INVALID, "world model without decoder needs act_in_dev");
      }
      if (wld) {
        if (!future_dev) return fail(PVAE_ERR_INVALID, "world part needs future_dev");
        CKR(small_chain(h, h->nets[PVAE_NET_WORLD_MODEL], batch, obs_dev, obs_ld, a_src, a_ld, future_dev, h->dsb, st));
      }
      if (val) {
        if (!value_dev) return fail(PVAE_ERR_INVALID, "value part needs value_dev");
        CKR(small_chain(h, h->nets[PVAE_NET_VALUE_BRANCH], batch, obs_dev, obs_ld, obs_dev + h->dsb, obs_ld, value_dev, 1, st));
      }
      CK(cudaGetLastError());
      return PVAE_OK;
    }
  }
  {
    const int64_t total = (int64_t)batch * h->x_ld;
    f32_to_planes_kernel<<<grid_for(total, 256, h->dev.sms), 256, 0, st>>>(obs_dev, obs_ld, obs_w, h->dsb, h->dsbp, h->xin, h->x_ld, plane_elems(h, h->x_ld), h->planes, batch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  EpiParams e;
  if (enc) {
    Net& te = h->nets[PVAE_NET_TASK_ENCODER];
    if (te.n_layers == 0) return fail(PVAE_ERR_INVALID, "model has no task encoder");
    NetIO in; in.nseg = 2; in.seg[0] = ws_view(h, h->xin, h->x_ld, h->dsb, batch); in.seg[1] = ws_view(h, h->xin + h->dsbp, h->x_ld, h->dsb, batch);
    memset(&e, 0, sizeof(e));
    e.type = EPI_STORE; e.out_f32 = h->ml; e.f32_sm = h->te_out; e.f32_sn = 1;
    CKR(net_forward(h, te, in, batch, e, st));
    const int prior = h->desc.latent_prior;
    const int64_t total = (int64_t)batch * z;
    reparam_fwd_kernel<<<grid_for(total, 256, h->dev.sms), 256, 0, st>>>(h->ml, eps_dev, h->eps, prior, prior && noise, seed, offset, batch, z,
                                                                         h->zb, h->zb_ld, plane_elems(h, h->zb_ld), h->planes, z_dev, mu_dev,
                                                                         prior ? logvar_dev : nullptr, nullptr, nullptr);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  } else if (dec) {
    if (!z_in_dev) return fail(PVAE_ERR_INVALID, "decoder without encoder needs z_in_dev");
    const int64_t total = (int64_t)batch * h->zb_ld;
    f32_to_planes_kernel<<<grid_for(total, 256, h->dev.sms), 256, 0, st>>>(z_in_dev, z, z, z, 1 << 30, h->zb, h->zb_ld, plane_elems(h, h->zb_ld), h->planes, batch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  const __nv_bfloat16* act_planes = h->ahat;
  if (dec) {
    Net& md = h->nets[PVAE_NET_MOTOR_DECODER];
    if (md.n_layers == 0) return fail(PVAE_ERR_INVALID, "model has no motor decoder");
    NetIO in; in.nseg = 2; in.seg[0] = ws_view(h, h->xin, h->x_ld, h->dsb, batch); in.seg[1] = ws_view(h, h->zb, h->zb_ld, z, batch);
    memset(&e, 0, sizeof(e));
    e.type = EPI_STORE;
    set_out(e, h, h->ahat, h->a_ld);
    if (act_out_dev) { e.out_f32 = act_out_dev; e.f32_sm = act_out_ld; e.f32_sn = 1; }
    CKR(net_forward(h, md, in, batch, e, st));
  } else if (wld) {
    if (!act_in_dev) return fail(PVAE_ERR_INVALID, "world model without decoder needs act_in_dev");
    const int64_t total = (int64_t)batch * h->a_ld;
    f32_to_planes_kernel<<<grid_for(total, 256, h->dev.sms), 256, 0, st>>>(act_in_dev, act_in_ld, h->da, h->da, 1 << 30, h->ain, h->a_ld, plane_elems(h, h->a_ld), h->planes, batch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    act_planes = h->ain;
  }
  if (wld) {
    Net& wm = h->nets[PVAE_NET_WORLD_MODEL];
    if (wm.n_layers == 0) return fail(PVAE_ERR_INVALID, "model has no world model");
    if (!future_dev) return fail(PVAE_ERR_INVALID, "world part needs future_dev");
    NetIO in; in.nseg = 2; in.seg[0] = ws_view(h, h->xin, h->x_ld, h->dsb, batch); in.seg[1] = ws_view(h, act_planes, h->a_ld, h->da, batch);
    memset(&e, 0, sizeof(e));
    e.type = EPI_STORE; e.out_f32 = future_dev; e.f32_sm = h->dsb; e.f32_sn = 1;
    CKR(net_forward(h, wm, in, batch, e, st));
  }
  if (val) {
    Net& vb = h->nets[PVAE_NET_VALUE_BRANCH];
    if (vb.n_layers == 0) return fail(PVAE_ERR_INVALID, "model has no value branch");
    if (!value_dev) return fail(PVAE_ERR_INVALID, "value part needs value_dev");
    NetIO in; in.nseg = 2; in.seg[0] = ws_view(h, h->xin, h->x_ld, h->dsb, batch); in.seg[1] = ws_view(h, h->xin + h->dsbp, h->x_ld, h->dsb, batch);
    memset(&e, 0, sizeof(e));
    e.type = EPI_STORE; e.out_f32 = value_dev; e.f32_sm = 1; e.f32_sn = 1;
    CKR(net_forward(h, vb, in, batch, e, st));
  }
  CK(cudaGetLastError());
  return PVAE_OK;
}

// Averaging all-reduce of a contiguous fp32 range of a symmetric (peer-mapped) allocation, one kernel (pvae_aux.cuh).
int pvae_symm_allreduce(const uint64_t* peer_ptrs_host, uint64_t multicast_ptr, int rank, int world, int64_t offset_elems, int64_t count_elems,
                        int64_t flags_offset_elems, pvae_stream s) {
  if (!peer_ptrs_host) return fail(PVAE_ERR_INVALID, "null peer pointer table");
  if (world < 2 || world > AR_MAX_RANKS || rank < 0 || rank >= world) return fail(PVAE_ERR_INVALID, "rank %d / world %d outside [2, %d]", rank, world, AR_MAX_RANKS);
  if (offset_elems < 0 || count_elems <= 0 || (offset_elems & 3) || (count_elems & 3) || (flags_offset_elems & 3))
    return fail(PVAE_ERR_INVALID, "range [%lld, +%lld) / flag block %lld must be multiples of 4 elements", (long long)offset_elems, (long long)count_elems, (long long)flags_offset_elems);
  if (flags_offset_elems < offset_elems + count_elems && flags_offset_elems + pvae_symm_flag_elems() > offset_elems)
    return fail(PVAE_ERR_INVALID, "the flag block overlaps the reduced range");
  SymmArgs a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < world; ++p) {
    if (!peer_ptrs_host[p] || (peer_ptrs_host[p] & 15)) return fail(PVAE_ERR_INVALID, "peer pointer %d is null or not 16-byte aligned", p);
    a.peer[p] = reinterpret_cast<float*>(peer_ptrs_host[p]);
  }
  a.mc = reinterpret_cast<float*>(multicast_ptr);
  a.rank = rank; a.world = world; a.off = offset_elems; a.count = count_elems; a.flags_off = flags_offset_elems;
  a.scale = 1.f / (float)world;
  static const int env_ctas = [] { const char* e = getenv("PVAE_SYMM_CTAS"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= AR_CTAS ? v : 0; }();
  const int ctas = env_ctas ? env_ctas : (symm_bulk() && !a.mc ? AR_CTAS / 2 : AR_CTAS);      // (bulk copies: 32 CTAs move as much as 64)
  return launch_symm(a, ctas, (cudaStream_t)s);
}

int64_t pvae_symm_flag_elems(void) { return (int64_t)AR_CTAS * AR_MAX_RANKS + AR_CTAS; }

// Debug: copy the per-unit clock stamps of the most recent GEMM launches (PVAE_DBG bit 5) to the host; returns the number
// of 64-bit words written (see g_trace in pvae_gemm.cuh).  Not part of the reference-facing interface.
int pvae_debug_trace(unsigned long long* out_host, int max_words, int clear) {
  const int n = TRACE_CTAS * TRACE_UNITS * 8;
  if (out_host) {
    if (max_words < n) return fail(PVAE_ERR_INVALID, "trace buffer needs %d words", n);
    CK(cudaMemcpyFromSymbol(out_host, g_trace, sizeof(unsigned long long) * n));
  }
  if (clear) {
    void* sym = nullptr;
    CK(cudaGetSymbolAddress(&sym, g_trace));
    CK(cudaMemset(sym, 0, sizeof(unsigned long long) * n));
  }
  return n;
}

// D = A . B^T on the tensor-core path, operands given as bf16 planes (unit tests / bench roofline).
int pvae_gemm_bf16(const void* A_dev, int a_major, const void* B_dev, int b_major, int M, int N, int K, int planes,
                   int splits, float* D_dev, pvae_stream s) {
  static Device dev;
  static bool dev_ok = false;
  if (!A_dev || !B_dev || !D_dev) return fail(PVAE_ERR_INVALID, "null argument");
  if (M <= 0 || N <= 0 || K <= 0 || (planes != 1 && planes != 2)) return fail(PVAE_ERR_INVALID, "bad shape / planes");
  if (!dev_ok) {
    int cur = 0;
    CK(cudaGetDevice(&cur));
    CKR(init_device(dev, cur));
    dev_ok = true;
  }
  GemmDesc d;
  d.a_major = a_major ? MAJOR_MN : MAJOR_K;
  d.b_major = b_major ? MAJOR_MN : MAJOR_K;
  View& a = d.A[0];
  a.base = reinterpret_cast<const __nv_bfloat16*>(A_dev); a.planes = planes;
  if (a_major) { a.ld = M; a.width = M; a.rows = K; } else { a.ld = K; a.width = K; a.rows = M; }
  a.ps = a.ld * a.rows;
  View& b = d.B;
  b.base = reinterpret_cast<const __nv_bfloat16*>(B_dev); b.planes = planes;
  if (b_major) { b.ld = N; b.width = N; b.rows = K; } else { b.ld = K; b.width = K; b.rows = N; }
  b.ps = b.ld * b.rows;
  d.M = M; d.N = N; d.K[0] = K;
  d.passes = planes == 2 ? 3 : 1;
  d.split = splits > 1;
  d.epi.type = EPI_WGRAD;
  d.epi.out_f32 = D_dev; d.epi.f32_sm = N; d.epi.f32_sn = 1; d.epi.f32_atomic = splits > 1;
  return launch_gemm(dev, d, (cudaStream_t)s);
}

}  // extern "C"
