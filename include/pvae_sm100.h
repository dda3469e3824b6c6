/* pvae_sm100.h -- C ABI of libpvae_sm100.so, the B200 (sm_100a) implementation of the PhysicsVAE training hot path.
 *
 * The reference (facebookresearch/PhysicsVAE) has no FFI: the path sits behind two Python plugin APIs
 * (RLlib custom model `PhysicsVAE`, rllib_model_torch.py:461-950; Ray Tune `Trainable`, torch_models.py:109-216 and
 * train_physics_vae.py:313-467).  Each entry point below names the reference code it replaces; the Python package
 * `physicsvae_b200` binds them with ctypes (physicsvae_b200/_abi.py) and mirrors the reference's classes on top.
 *
 * Conventions
 *   - every function returns 0 on success, a negative pvae_status otherwise; the message is in pvae_last_error();
 *     nothing throws across the ABI.
 *   - all pointers named *_dev are device pointers owned by the caller (PyTorch's caching allocator); the library
 *     allocates only its handle, bf16 shadow weights, TMA descriptors and a few scalars, and frees them in pvae_destroy.
 *   - all work is enqueued on the given stream, never synchronises, never allocates after pvae_create(): every step
 *     entry point is CUDA-graph capturable.
 *   - a handle is not thread-safe; distinct handles are independent (one per process / GPU).
 *   - fp32 parameters / gradients use the reference's nn.Linear layout: W[out][in] row-major, b[out].
 */
#ifndef PVAE_SM100_H
#define PVAE_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVAE_ABI_VERSION 2
#define PVAE_MAX_LAYERS 8

typedef struct pvae_engine* pvae_handle;
typedef void* pvae_stream;                 /* cudaStream_t */

enum pvae_status {
  PVAE_OK = 0,
  PVAE_ERR_INVALID = -1,                   /* bad argument / unsupported configuration */
  PVAE_ERR_CUDA = -2,                      /* a CUDA runtime / driver call failed */
  PVAE_ERR_STATE = -3                      /* call order violated (nothing bound yet) */
};

/* activation registry of the reference: get_activation_fn, rllib_model_torch.py:30-46 */
enum pvae_act { PVAE_ACT_LINEAR = 0, PVAE_ACT_RELU = 1, PVAE_ACT_TANH = 2, PVAE_ACT_SIGMOID = 3, PVAE_ACT_ELU = 4, PVAE_ACT_SWISH = 5 };

/* arithmetic of the tensor-core contractions */
enum pvae_precision {
  PVAE_PREC_BF16 = 1,                      /* bf16 operands, fp32 accumulate (performance mode) */
  PVAE_PREC_BF16X3 = 3                     /* hi/lo split bf16, 3 products, fp32 accumulate: matches fp32 to ~1e-6 (parity mode) */
};

/* the four MLPs of PhysicsVAE (rllib_model_torch.py:638-699) */
enum pvae_net { PVAE_NET_TASK_ENCODER = 0, PVAE_NET_MOTOR_DECODER = 1, PVAE_NET_WORLD_MODEL = 2, PVAE_NET_VALUE_BRANCH = 3, PVAE_NUM_NETS = 4 };

/* one FC stack: rllib_model_torch.FC (rllib_model_torch.py:234-282); the input width is implied by the net's role */
typedef struct {
  int32_t n_layers;                        /* number of Linear layers (hidden + output); 0 = net absent */
  int32_t out_dims[PVAE_MAX_LAYERS];       /* nn.Linear out_features per layer */
  int32_t acts[PVAE_MAX_LAYERS];           /* pvae_act per layer */
  int32_t in_dims[2];                      /* {0, 0}: input widths follow from the net's role.  {k0 > 0, k1 >= 0}: a stand-alone FC
                                              stack whose input is k0 (+ k1) columns wide; such a net runs through pvae_fc_forward
                                              only and its output width is free (FC(size_in, size_out, layers), :234-246) */
} pvae_net_desc;

/* PhysicsVAE.__init__ (rllib_model_torch.py:511-727), default wiring: encoder sees (body, task), decoder sees (body, z) */
typedef struct {
  int32_t dim_state_body;                  /* dsb; dim_state_task == dsb on this path (train_physics_vae.py:198-214) */
  int32_t dim_action;                      /* da */
  int32_t latent_dim;                      /* z  (task_encoder_output_dim) */
  int32_t latent_prior;                    /* 1 = "normal_zero_mean_one_std" (encoder emits 2z), 0 = False (encoder emits z) */
  int32_t precision;                       /* pvae_precision */
  int32_t max_batch;                       /* capacity of the workspace in rows */
  pvae_net_desc nets[PVAE_NUM_NETS];
} pvae_model_desc;

/* loss bookkeeping written by the step functions (device, fp32):
 *   [0] total  [1] MSE(a, a_hat)  [2] KL  [3] MSE(s2, world(s1, a_gt))  [4] MSE(s2, world(s1, a_hat))
 * weights as in train_physics_vae.py:430-434 */
#define PVAE_LOSS_SLOTS 8                   /* slots 5-7 are reserved (callers allocate 8 floats) */

const char* pvae_last_error(void);
int pvae_abi_version(void);

/* --- lifetime ------------------------------------------------------------------------------------------------ */
/* replaces: PhysicsVAE.__init__ network construction (rllib_model_torch.py:638-699) */
int pvae_create(pvae_handle* out, const pvae_model_desc* desc, int device);
int pvae_destroy(pvae_handle h);

/* fp32 master parameters of one net and its flat gradient buffer.
 * W_dev[l] / b_dev[l]: nn.Linear weight/bias of layer l.  grad_flat_dev: [dW0 | db0 | dW1 | db1 | ...] (may be NULL for
 * nets that are never trained).  replaces: model.parameters() / .grad consumed by torch.optim.Adam (torch_models.py:119-122) */
int pvae_bind_net(pvae_handle h, int net, const float* const* W_dev, const float* const* b_dev, float* grad_flat_dev);
int64_t pvae_net_grad_elems(pvae_handle h, int net);
/* Parameters of the activations: rllib's Swish (ray 1.11, what get_activation_fn("swish") returns, rllib_model_torch.py:33-35) computes
 * x * sigmoid(beta x) with beta a trainable scalar per layer.  beta_dev[l] (fp32, PVAE_MAX_LAYERS entries) = beta of layer l's activation
 * (entries of non-swish layers are ignored); dbeta_dev[l] receives d loss / d beta from the training steps (zeroed by them; may be
 * NULL).  Not bound: beta = 1. */
int pvae_bind_act_params(pvae_handle h, int net, const float* beta_dev, float* dbeta_dev);

/* refresh the bf16 (hi/lo) shadow copies of the fp32 masters of every net whose bit is set in net_mask
 * (after load_state_dict or optimizer.step()).  */
int pvae_sync_weights(pvae_handle h, uint32_t net_mask, pvae_stream s);

/* --- optimizer -------------------------------------------------------------------------------------------------------- */
/* One Adam step for the layers of `net` selected by layer_mask (bit l), on the bound fp32 masters and the bound gradient
 * buffer, followed -- in the same kernels -- by the refresh of the bf16 shadow operands.  exp_avg_dev / exp_avg_sq_dev: fp32
 * state in the flat layout of the gradient buffer ([W0|b0|W1|b1|...]), owned by the caller.  step_dev: device fp32 scalar
 * holding the number of steps taken so far by these parameters; the call increments it on the device (the update uses
 * step + 1 and the last thread block to finish writes it back), so the call is CUDA-graph replayable.  One launch covers
 * the whole net when its layers are laid out back to back and 16-byte aligned, otherwise one launch per layer.  Arithmetic follows torch.optim.Adam(amsgrad=False, maximize=False).
 * replaces: optimizer.step() (torch_models.py:143) + pvae_sync_weights for the stepped layers. */
int pvae_adam_step(pvae_handle h, int net, uint32_t layer_mask, float* exp_avg_dev, float* exp_avg_sq_dev, float* step_dev,
                   float lr, float beta1, float beta2, float eps, float weight_decay, pvae_stream s);

/* --- workspace ----------------------------------------------------------------------------------------------- */
int pvae_workspace_bytes(pvae_handle h, size_t* bytes);
int pvae_bind_workspace(pvae_handle h, void* ws_dev, size_t bytes);

/* --- transition buffers -------------------------------------------------------------------------------------- */
/* Resident transition buffer = the reference's DatasetBase.X / .Y (torch_models.py:39-58; built by
 * load_dataset_for_PhysicsVAE, train_physics_vae.py:117-164) converted once to bf16 hi(/lo) planes:
 *   x: one row per transition (s_t | 0.. | a_t | 0.. | s_{t+1} | 0..), a_t at column roundup(dsb, 8), s_{t+1} at column
 *      roundup(roundup(dsb, 8) + da, 64), row length a multiple of 64 columns (cat[s_t, a_t] is then one
 *      K segment of the row);  y: [n_rows][roundup(da, 64)], a_t once more as the 16-byte aligned MSE target.
 * The layout is private to the library: callers size the buffer with pvae_transitions_bytes and fill it with pvae_ingest.
 * pvae_ingest converts rows [0, n_rows) of the raw arrays (x_raw_dev: float64 if x_is_f64 else float32, row stride
 * 2*dsb; y_raw_dev: float32, row stride da) into rows [dst_row, dst_row + n_rows) of buf_dev.
 * replaces: DatasetBase.__getitem__ + default collate (torch.Tensor(x) per item, torch_models.py:52-68). */
int pvae_transitions_bytes(pvae_handle h, int64_t n_rows, size_t* bytes);
int pvae_ingest(pvae_handle h, void* buf_dev, int64_t buf_rows, int64_t dst_row, const void* x_raw_dev, int x_is_f64,
                const float* y_raw_dev, int64_t n_rows, pvae_stream s);
/* Dataset build on the device: the same resident rows straight from the episode arrays, every state uploaded once.
 * states_dev: [n_states][dsb] float64 / float32 -- the state_body rows of all episodes back to back; actions_dev:
 * [n_states][da] float32; first_state_dev: [n_rows] int64, the state row of s_t of each transition (s_{t+1} is the next row:
 * transitions never cross an episode boundary because the builder never emits the last state of an episode as s_t).
 * replaces: the per-transition np.hstack loop of load_dataset_for_PhysicsVAE (train_physics_vae.py:133-156) + DatasetBase
 * collation, and halves the upload (1.76 KB instead of 3.3 KB per transition at 197 / 45).
 * states_is_f64: 1 = float64 states, 0 = float32 states (actions float32 in both); 2 = states AND actions are bf16 (a loader that keeps
 * the dataset in the engine's operand precision -- PVAE_PREC_BF16 only; the same values the device-side conversion of the fp32 data
 * produces, at half the upload again). */
int pvae_ingest_episodes(pvae_handle h, void* buf_dev, int64_t buf_rows, int64_t dst_row, const void* states_dev,
                         int states_is_f64, int64_t n_states, const void* actions_dev, const int64_t* first_state_dev,
                         int64_t n_rows, pvae_stream s);
/* select the buffer the step functions read from; mini-batch b = rows [cursor, cursor + batch) */
int pvae_bind_transitions(pvae_handle h, const void* buf_dev, int64_t buf_rows);
int pvae_set_cursor(pvae_handle h, int64_t row, pvae_stream s);
/* cursor += delta; if cursor + batch > limit then cursor = cursor mod delta -- a rank that started on row s < delta (its slice of
 * the first global mini-batch of `delta` rows) is back on row s, a single rank on row 0  (device-side, graph-replayable) */
int pvae_advance_cursor(pvae_handle h, int64_t delta, int64_t batch, int64_t limit, pvae_stream s);

/* --- the training steps (forward + loss + backward; gradients land in the bound grad buffers) ------------------ */
/* World-model phase: loss = s_coeff * MSE(s2, WM(cat[s1, a_gt])); gradients for the world model only.
 * replaces: TrainModel.compute_loss world branch + loss.backward() (train_physics_vae.py:412-414; torch_models.py:141-142).
 * The reference's discarded full-model forward (train_physics_vae.py:377-378) is not executed. */
int pvae_world_step(pvae_handle h, int batch, float s_coeff, float* loss_dev, pvae_stream s);

/* VAE phase: loss = a_coeff*MSE(a_gt, a_hat) + kl_coeff*KL + cyc_coeff*MSE(s2, WM(cat[s1, a_hat])), gradients for task
 * encoder + motor decoder, world model differentiated through but frozen.
 * eps_dev: [batch][z] fp32 standard-normal noise (the reference draws it with torch.randn_like, rllib_model_torch.py:737);
 * if NULL, a Philox4x32-10 stream keyed by (seed, offset) is generated in-kernel.  noise == 0: z = mu
 * (PhysicsVAE.latent_prior_noise False, rllib_model_torch.py:734-740).
 * replaces: TrainModel.compute_loss VAE branch + backward (train_physics_vae.py:377-434). */
int pvae_vae_step(pvae_handle h, int batch, const float* eps_dev, uint64_t seed, uint64_t offset, int noise, float a_coeff,
                  float kl_coeff, float cyc_coeff, float* loss_dev, pvae_stream s);

/* compute_loss with lookahead L > 1 (train_physics_vae.py:361-435): an autoregressive rollout.  Step t runs the whole model on
 * (s1_t, s2_gt_t) with s1_0 from the data and s1_{t+1} = the world model's prediction from the decoded action of step t (:421); the loss
 * terms are means over the L steps (:423-428); gradients flow through time (every net is differentiated w.r.t. its body-state
 * input).  phase 0: coefficients (a, kl, s, cyc) = (0, 0, s_coeff, 0), gradients for the world model; phase 1: (a, kl, 0, cyc),
 * gradients for encoder + decoder.  tbufs_host: host array of L resident transition buffers (pvae_ingest of X[:, t, :], Y[:, t, :]),
 * all of buf_rows rows; the mini-batch is rows [cursor, cursor + batch) of each.  eps_dev: [L][batch][z] or NULL (Philox, offset + t).
 * Needs 2 * L copies of the workspace (pvae_rollout_workspace_bytes / pvae_bind_rollout_workspace): one per step, and in phase 0 one
 * more per step for the second world-model pass (on the ground-truth action, :412-414). */
int pvae_rollout_workspace_bytes(pvae_handle h, int lookahead, size_t* bytes);
int pvae_bind_rollout_workspace(pvae_handle h, void* ws_dev, size_t bytes, int lookahead);
int pvae_rollout_step(pvae_handle h, int phase, int batch, int lookahead, const void* const* tbufs_host, int64_t buf_rows, const float* eps_dev,
                      uint64_t seed, uint64_t offset, int noise, float a_coeff, float kl_coeff, float s_coeff, float cyc_coeff, float* loss_dev,
                      pvae_stream s);

/* Forward + loss only (no gradients are touched): the test pass of the reference, `with torch.no_grad(): compute_test_loss`
 * (torch_models.py:147-155).  phase 0: world model, loss = s_coeff * MSE(s2, WM(cat[s1, a_gt])); phase 1: the VAE loss with the
 * arguments of pvae_vae_step.  loss_dev as for the step functions. */
int pvae_eval_loss(pvae_handle h, int phase, int batch, const float* eps_dev, uint64_t seed, uint64_t offset, int noise, float a_coeff,
                   float kl_coeff, float s_coeff, float cyc_coeff, float* loss_dev, pvae_stream s);

/* Deterministic gradients (a decided reduction order).  enable != 0: weight gradients are computed without split-K (one
 * CTA pair walks the whole batch of its tile in order) and bias gradients by two-pass ordered column sums instead of the epilogues'
 * atomics -- every gradient is then bit-identical from run to run (slower: the 1024 x 1024 weight gradient uses 16 of 74 CTA pairs).
 * Default off: fp32 red.global.add in arrival order (differences in the last bits). */
int pvae_set_deterministic(pvae_handle h, int enable);

/* Device-side counter of the Philox noise stream.  enable != 0: every pvae_vae_step / pvae_eval_loss call that draws noise itself
 * (eps_dev == NULL, noise != 0) uses offset + counter as its stream offset and adds `stride` to the counter when it finishes, so a
 * captured CUDA graph draws fresh noise on every replay (the reference draws torch.randn_like per call, rllib_model_torch.py:737).
 * The call (re)sets the counter to `value`.  enable == 0 (default): the offset argument is used as given. */
int pvae_noise_counter(pvae_handle h, int enable, uint64_t value, uint64_t stride, pvae_stream s);

/* --- inference API ------------------------------------------------------------------------------------------- */
/* FC.forward of one net (rllib_model_torch.py:274-275): out[batch][out_width] = FC(in[batch][in_width]), fp32 rows in and out
 * (row strides in_ld / out_ld, elements); for a net with two input segments `in` is their concatenation. */
int pvae_fc_forward(pvae_handle h, int net, int batch, const float* in_dev, int64_t in_ld, float* out_dev, int64_t out_ld, pvae_stream s);

/* PhysicsVAE.forward and its parts (rllib_model_torch.py:742-853).  obs_dev: fp32 [batch][2*dsb] (row stride obs_ld).
 * parts: bit 0 encoder (+reparameterise), bit 1 decoder, bit 2 world, bit 3 value branch.
 * Inputs consumed only when the producing part is not run: z_in_dev ([batch][z], decoder without encoder),
 * act_in_dev ([batch][da], world without decoder).  eps_dev NULL + noise=0 -> z = mu (latent_prior_noise False).
 * Outputs (each may be NULL): act_out [batch][da] (the first half of `logits`; AppendLogStd's constant half is
 * bookkeeping done by the caller), mu/logvar/z [batch][z], future [batch][dsb], value [batch]. */
#define PVAE_PART_ENCODER 1u
#define PVAE_PART_DECODER 2u
#define PVAE_PART_WORLD 4u
#define PVAE_PART_VALUE 8u
int pvae_forward(pvae_handle h, uint32_t parts, int batch, const float* obs_dev, int64_t obs_ld, const float* z_in_dev,
                 const float* act_in_dev, int64_t act_in_ld, const float* eps_dev, int noise, uint64_t seed,
                 uint64_t offset, float* act_out_dev, int64_t act_out_ld, float* mu_dev, float* logvar_dev, float* z_dev,
                 float* future_dev, float* value_dev, pvae_stream s);

/* --- data-parallel gradient exchange (the reference is single-process, there is nothing it replaces) ------------------
 * Averaging all-reduce, in place, of the fp32 range [offset, offset + count) of a SYMMETRIC allocation: the same buffer layout
 * allocated on every rank of one node and mapped into every peer (torch.distributed._symmetric_memory / CUDA IPC / VMM), its base
 * address on rank p being peer_ptrs_host[p] (a host array of `world` device addresses as seen from THIS rank).  One kernel: device-side
 * rank barrier through flags inside the allocation, rank r reduces slice r over NVLink (cp.async.bulk copies of the peers' slices
 * into shared memory, five 32 KiB stages in flight per CTA; PVAE_SYMM_BULK=0: per-thread peer loads; multicast_ptr != 0:
 * multimem.ld_reduce, the switch adds) and writes the averaged slice into every rank's copy (bulk / peer stores / multimem.st), barrier.  All ranks must call
 * it the same number of times; every replica ends up with bit-identical values.  The flag block -- pvae_symm_flag_elems() fp32-sized
 * words at flags_offset inside the same allocation, zeroed on every rank before the first call -- carries the sequence numbers, so the
 * call is CUDA-graph replayable.  world <= 8. */
int pvae_symm_allreduce(const uint64_t* peer_ptrs_host, uint64_t multicast_ptr, int rank, int world, int64_t offset_elems, int64_t count_elems,
                        int64_t flags_offset_elems, pvae_stream s);
int64_t pvae_symm_flag_elems(void);
/* Overlap (experimental, the Python trainer arms it only with PVAE_OVERLAP=1: it pays at N = 2 and loses at N = 8):
 * arm the training steps of `h` to exchange the range [offset, offset + count) themselves -- on a side stream, with `ctas`
 * CTAs, as soon as the gradients inside it are complete (world-model step: after the weight gradient of layer 1, i.e. everything but
 * layer 0; VAE step: after the decoder's backward pass, i.e. the decoder's gradients) -- while the remaining backward GEMMs run on the
 * other SMs; the step joins the side stream before it finishes (fork / join are stream events: one CUDA graph).  The caller
 * exchanges the rest (layer 0 / the encoder, plus the loss slots) after the step with pvae_symm_allreduce.  Every rank must arm the
 * same range.  peer_ptrs_host == NULL switches it off.  pvae_eval_loss and pvae_rollout_step never fork. */
int pvae_set_exchange(pvae_handle h, const uint64_t* peer_ptrs_host, uint64_t multicast_ptr, int rank, int world, int64_t offset_elems,
                      int64_t count_elems, int64_t flags_offset_elems, int ctas);
/* The armed exchange on its own (same range, same CTAs), for a rank whose slice of a mini-batch is empty: it runs no training step but
 * must still take part in the exchange the other ranks' steps perform. */
int pvae_run_exchange(pvae_handle h, pvae_stream s);

/* --- kernel-level entry (unit tests, bench roofline) ----------------------------------------------------------- */
/* D[M][N] (fp32, row-major) = A . B^T on the tcgen05 path.
 *   a_major 0: A is [M][K] row-major (ld K);  1: A is [K][M] row-major (ld M)
 *   b_major 0: B is [N][K] row-major;         1: B is [K][N] row-major
 * planes 1: bf16 operands; 2: operands are hi plane followed by lo plane, bf16x3 products.
 * splits > 1 splits K over CTAs and accumulates with red.global.add (D must be zeroed by the caller). */
int pvae_gemm_bf16(const void* A_dev, int a_major, const void* B_dev, int b_major, int M, int N, int K, int planes,
                   int splits, float* D_dev, pvae_stream s);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t pvae_launch_count(void);

/* profiling aid (no reference counterpart): with PVAE_DBG bit 5 set every GEMM launch records per-unit clock64 stamps of its
 * producer / MMA / epilogue roles; this copies them to the host (out_host may be NULL) and optionally clears them.
 * Returns the number of 64-bit words of the trace, or a negative status. */
int pvae_debug_trace(unsigned long long* out_host, int max_words, int clear);

#ifdef __cplusplus
}
#endif
#endif /* PVAE_SM100_H */
