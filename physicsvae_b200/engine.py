"""Host-side handle on the sm_100a engine: owns the torch tensors (workspace, transition buffer, flat gradients) whose raw
device pointers are handed to libpvae_sm100.so, and nothing else.  PyTorch is plumbing here -- device memory, streams --
all arithmetic of the hot path happens inside the library.
"""
import ctypes as C

import torch

from . import _abi

NET_NAMES = ("task_encoder", "motor_decoder", "world_model", "value_branch")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One engine per process / GPU (the handle is not thread-safe; include/pvae_sm100.h).

    nets: dict net name -> list of (out_features, activation name) per Linear layer, in the reference's FC order
    (rllib_model_torch.py:243-264).  precision: "bf16" (performance) or "bf16x3" (fp32-accurate parity mode).
    """

    def __init__(self, dim_state_body, dim_action, latent_dim, nets, latent_prior=True, precision="bf16x3", max_batch=4096,
                 device=None, in_dims=None):
        if not torch.cuda.is_available():
            raise _abi.PvaeError("physicsvae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _abi.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.dsb, self.da, self.z = int(dim_state_body), int(dim_action), int(latent_dim)
        self.latent_prior = bool(latent_prior)
        self.precision = precision
        self.planes = 2 if precision == "bf16x3" else 1
        self.max_batch = int(max_batch)
        desc = _abi.ModelDesc()
        desc.dim_state_body, desc.dim_action, desc.latent_dim = self.dsb, self.da, self.z
        desc.latent_prior = 1 if latent_prior else 0
        desc.precision = {"bf16": _abi.PREC_BF16, "bf16x3": _abi.PREC_BF16X3}[precision]
        desc.max_batch = self.max_batch
        self.layers = {}
        for i, name in enumerate(NET_NAMES):
            spec = nets.get(name) or []
            if len(spec) > _abi.PVAE_MAX_LAYERS:
                raise ValueError("%s: at most %d layers" % (name, _abi.PVAE_MAX_LAYERS))
            desc.nets[i].n_layers = len(spec)
            if in_dims and name in in_dims:        # stand-alone FC stack: explicit input widths (fc_forward only)
                desc.nets[i].in_dims[0], desc.nets[i].in_dims[1] = int(in_dims[name][0]), int(in_dims[name][1])
            for l, (out, act) in enumerate(spec):
                if act not in _abi.ACT_IDS:
                    raise ValueError("Unknown activation ({})!".format(act))
                desc.nets[i].out_dims[l] = int(out)
                desc.nets[i].acts[l] = _abi.ACT_IDS[act]
            self.layers[name] = list(spec)
        self._h = C.c_void_p(0)
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_create(C.byref(self._h), C.byref(desc), self.device.index))
            nbytes = C.c_size_t(0)
            _abi.check(self.lib.pvae_workspace_bytes(self._h, C.byref(nbytes)))
            # zero-initialised: padding columns are never written and must stay finite
            self.workspace = torch.zeros(nbytes.value + 1024, dtype=torch.uint8, device=self.device)
            off = (-self.workspace.data_ptr()) % 1024
            _abi.check(self.lib.pvae_bind_workspace(self._h, C.c_void_p(self.workspace.data_ptr() + off), nbytes.value))
        self.loss = torch.zeros(_abi.PVAE_LOSS_SLOTS, dtype=torch.float32, device=self.device)
        self._keep = {}          # tensors whose pointers the library holds
        self.transitions = None
        self.n_rows = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.pvae_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters ---------------------------------------------------------------------------------------------
    def grad_elems(self, name):
        return int(self.lib.pvae_net_grad_elems(self._h, NET_NAMES.index(name)))

    def bind_net(self, name, weights, biases, grad_flat=None):
        """weights[l]: fp32 [out, in] contiguous CUDA tensors (nn.Linear layout); grad_flat: fp32 [grad_elems] or None."""
        n = len(weights)
        if n != len(self.layers[name]) or n != len(biases):
            raise ValueError("%s: expected %d layers" % (name, len(self.layers[name])))
        for t in list(weights) + list(biases) + ([grad_flat] if grad_flat is not None else []):
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("%s: parameters must be contiguous fp32 tensors on %s" % (name, self.device))
        if grad_flat is not None and grad_flat.numel() != self.grad_elems(name):
            raise ValueError("%s: gradient buffer has %d elements, expected %d" % (name, grad_flat.numel(), self.grad_elems(name)))
        W = (C.c_void_p * n)(*[t.data_ptr() for t in weights])
        b = (C.c_void_p * n)(*[t.data_ptr() for t in biases])
        _abi.check(self.lib.pvae_bind_net(self._h, NET_NAMES.index(name), W, b, _ptr(grad_flat)))
        self._keep[name] = (list(weights), list(biases), grad_flat)

    def bind_act_params(self, name, beta, dbeta=None):
        """beta: fp32 [PVAE_MAX_LAYERS] CUDA tensor, entry l = beta of layer l's swish activation; dbeta: its gradient accumulator
        (include/pvae_sm100.h, pvae_bind_act_params)."""
        for t in (beta, dbeta):
            if t is not None and (t.device != self.device or t.dtype != torch.float32 or t.numel() != _abi.PVAE_MAX_LAYERS or not t.is_contiguous()):
                raise ValueError("activation parameters must be contiguous fp32 [%d] tensors on %s" % (_abi.PVAE_MAX_LAYERS, self.device))
        _abi.check(self.lib.pvae_bind_act_params(self._h, NET_NAMES.index(name), _ptr(beta), _ptr(dbeta)))
        self._keep[name + "/act"] = (beta, dbeta)

    def sync_weights(self, names=None):
        mask = 0
        for name in (names or [n for n in NET_NAMES if n in self._keep]):
            if name.endswith("/act"):
                continue
            mask |= 1 << NET_NAMES.index(name)
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_sync_weights(self._h, mask, _stream()))

    def adam_step(self, name, layer_mask, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
        """Fused Adam + shadow refresh for the selected layers of one net (include/pvae_sm100.h, pvae_adam_step)."""
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_adam_step(self._h, NET_NAMES.index(name), int(layer_mask), _ptr(exp_avg), _ptr(exp_avg_sq),
                                               _ptr(step), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
                                               _stream()))

    # ---- resident transition buffer -------------------------------------------------------------------------------
    def transitions_bytes(self, n_rows):
        nbytes = C.c_size_t(0)
        _abi.check(self.lib.pvae_transitions_bytes(self._h, int(n_rows), C.byref(nbytes)))
        return nbytes.value

    def alloc_transitions(self, n_rows):
        return self.bind_transitions(torch.zeros(self.transitions_bytes(n_rows), dtype=torch.uint8, device=self.device), n_rows)

    def bind_transitions(self, buf, n_rows):
        """Select the resident transition buffer the step functions read (several may be kept: train / test sets)."""
        self.transitions = buf
        self.n_rows = int(n_rows)
        _abi.check(self.lib.pvae_bind_transitions(self._h, _ptr(buf), self.n_rows))
        return buf

    def ingest(self, x_raw, y_raw, dst_row=0):
        """x_raw: CUDA [n, 2*dsb] float64 or float32; y_raw: CUDA [n, da] float32 (DatasetBase.X / .Y, torch_models.py:39-58)."""
        if self.transitions is None:
            raise _abi.PvaeError("alloc_transitions() first")
        x_raw = x_raw.reshape(x_raw.shape[0], -1)
        y_raw = y_raw.reshape(y_raw.shape[0], -1)
        if x_raw.shape[1] != 2 * self.dsb or y_raw.shape[1] != self.da or x_raw.shape[0] != y_raw.shape[0]:
            raise ValueError("transition arrays must be [n, %d] and [n, %d]" % (2 * self.dsb, self.da))
        if x_raw.dtype not in (torch.float64, torch.float32):
            raise ValueError("x must be float64 or float32")
        x_raw = x_raw.to(self.device).contiguous()
        y_raw = y_raw.to(self.device, torch.float32).contiguous()
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_ingest(self._h, _ptr(self.transitions), self.n_rows, int(dst_row), _ptr(x_raw),
                                            1 if x_raw.dtype == torch.float64 else 0, _ptr(y_raw), x_raw.shape[0], _stream()))

    def ingest_episodes(self, states, actions, first_state, dst_row=0, check=True):
        """Dataset build on the device: states CUDA [S, dsb] float64 / float32 (all episodes back to back, every state once),
        actions CUDA [S, da] float32, first_state CUDA [n] int64 (state row of s_t per transition; s_{t+1} is the next row)."""
        if self.transitions is None:
            raise _abi.PvaeError("alloc_transitions() first")
        if states.dim() != 2 or states.shape[1] != self.dsb or actions.shape != (states.shape[0], self.da):
            raise ValueError("episode arrays must be [S, %d] and [S, %d]" % (self.dsb, self.da))
        if states.dtype == torch.bfloat16:
            # a loader that keeps the dataset in the engine's operand precision (bf16 mode only): states and actions both bf16
            if self.planes != 1:
                raise ValueError("bf16 episode arrays need precision='bf16' (the fp32-accurate mode keeps hi + lo planes)")
            code = 2
            states = states.to(self.device).contiguous()
            actions = actions.to(self.device, torch.bfloat16).contiguous()
        elif states.dtype in (torch.float64, torch.float32):
            code = 1 if states.dtype == torch.float64 else 0
            states = states.to(self.device).contiguous()
            actions = actions.to(self.device, torch.float32).contiguous()
        else:
            raise ValueError("states must be float64, float32 or bfloat16")
        first_state = first_state.to(self.device, torch.int64).contiguous()
        # (the range check reads the index back to the host: per-step callers that built the index themselves skip it)
        if check and first_state.numel() and (int(first_state.min()) < 0 or int(first_state.max()) + 1 >= states.shape[0]):
            raise ValueError("first_state index out of range")
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_ingest_episodes(self._h, _ptr(self.transitions), self.n_rows, int(dst_row), _ptr(states), code,
                                                     states.shape[0], _ptr(actions), _ptr(first_state), first_state.numel(), _stream()))

    def set_cursor(self, row):
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_set_cursor(self._h, int(row), _stream()))

    def advance_cursor(self, delta, batch, limit):
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_advance_cursor(self._h, int(delta), int(batch), int(limit), _stream()))

    # ---- training steps ---------------------------------------------------------------------------------------------
    def world_step(self, batch, s_coeff=1.0):
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_world_step(self._h, int(batch), float(s_coeff), _ptr(self.loss), _stream()))
        return self.loss

    def vae_step(self, batch, eps=None, seed=0, offset=0, noise=True, a_coeff=1.0, kl_coeff=1.0, cyc_coeff=1e-3):
        if eps is not None:
            if eps.shape != (batch, self.z) or eps.dtype != torch.float32 or eps.device != self.device or not eps.is_contiguous():
                raise ValueError("eps must be a contiguous fp32 [%d, %d] tensor on %s" % (batch, self.z, self.device))
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_vae_step(self._h, int(batch), _ptr(eps), int(seed), int(offset), 1 if noise else 0,
                                              float(a_coeff), float(kl_coeff), float(cyc_coeff), _ptr(self.loss), _stream()))
        return self.loss

    def rollout_step(self, batch, world, buffers, buf_rows, eps=None, seed=0, offset=0, noise=True, a_coeff=1.0, kl_coeff=1.0, s_coeff=1.0,
                     cyc_coeff=1e-3):
        """compute_loss with lookahead L = len(buffers) > 1 (train_physics_vae.py:361-435): autoregressive rollout, forward + loss +
        backward through time.  buffers[t]: resident transition buffer of step t (ingest of X[:, t, :], Y[:, t, :])."""
        L = len(buffers)
        if eps is not None:
            if eps.shape != (L, batch, self.z) or eps.dtype != torch.float32 or eps.device != self.device or not eps.is_contiguous():
                raise ValueError("eps must be a contiguous fp32 [%d, %d, %d] tensor on %s" % (L, batch, self.z, self.device))
        need = C.c_size_t(0)
        _abi.check(self.lib.pvae_rollout_workspace_bytes(self._h, L, C.byref(need)))
        ws = getattr(self, "_rollout_ws", None)
        if ws is None or ws.numel() < need.value + 1024 or self._rollout_L < L:
            ws = self._rollout_ws = torch.zeros(need.value + 1024, dtype=torch.uint8, device=self.device)
            self._rollout_L = L
            off = (-ws.data_ptr()) % 1024
            _abi.check(self.lib.pvae_bind_rollout_workspace(self._h, C.c_void_p(ws.data_ptr() + off), need.value, L))
        ptrs = (C.c_void_p * L)(*[b.data_ptr() for b in buffers])
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_rollout_step(self._h, 0 if world else 1, int(batch), L, ptrs, int(buf_rows), _ptr(eps), int(seed), int(offset),
                                                  1 if noise else 0, float(a_coeff), float(kl_coeff), float(s_coeff), float(cyc_coeff),
                                                  _ptr(self.loss), _stream()))
        return self.loss

    def eval_loss(self, batch, world, eps=None, seed=0, offset=0, noise=True, a_coeff=1.0, kl_coeff=1.0, s_coeff=1.0, cyc_coeff=1e-3):
        """Forward + loss only, no gradient is touched (the reference's test pass, torch_models.py:147-155)."""
        if eps is not None:
            if eps.shape != (batch, self.z) or eps.dtype != torch.float32 or eps.device != self.device or not eps.is_contiguous():
                raise ValueError("eps must be a contiguous fp32 [%d, %d] tensor on %s" % (batch, self.z, self.device))
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_eval_loss(self._h, 0 if world else 1, int(batch), _ptr(eps), int(seed), int(offset),
                                               1 if noise else 0, float(a_coeff), float(kl_coeff), float(s_coeff), float(cyc_coeff),
                                               _ptr(self.loss), _stream()))
        return self.loss

    def run_exchange(self):
        """The armed early-range exchange on its own (a rank with an empty slice of the mini-batch), pvae_run_exchange."""
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_run_exchange(self._h, _stream()))

    def set_deterministic(self, enable=True):
        """Run-to-run bit-identical gradients: no split-K, ordered bias-gradient sums (include/pvae_sm100.h, pvae_set_deterministic)."""
        _abi.check(self.lib.pvae_set_deterministic(self._h, 1 if enable else 0))

    def noise_counter(self, enable, value=0, stride=1):
        """Device-side Philox offset counter (include/pvae_sm100.h, pvae_noise_counter)."""
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_noise_counter(self._h, 1 if enable else 0, int(value), int(stride), _stream()))

    # ---- inference ------------------------------------------------------------------------------------------------
    def fc_forward(self, name, x):
        """FC.forward of one net on fp32 rows [B, in] -> [B, out] (rllib_model_torch.py:274-275)."""
        x = x.to(self.device, torch.float32).contiguous()
        if x.dim() != 2:
            raise ValueError("fc_forward wants [batch, features]")
        out = torch.empty(x.shape[0], self.layers[name][-1][0], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _abi.check(self.lib.pvae_fc_forward(self._h, NET_NAMES.index(name), x.shape[0], _ptr(x), x.shape[1], _ptr(out),
                                                out.shape[1], _stream()))
        return out

    def forward(self, obs, parts, z_in=None, act_in=None, eps=None, noise=False, seed=0, offset=0):
        """Runs the selected parts of PhysicsVAE.forward (rllib_model_torch.py:742-853); returns a dict of fp32 outputs."""
        B = obs.shape[0]
        dev, f32 = self.device, torch.float32
        obs = obs.to(dev, f32).contiguous()
        out = {}
        enc, dec = parts & _abi.PART_ENCODER, parts & _abi.PART_DECODER
        wld, val = parts & _abi.PART_WORLD, parts & _abi.PART_VALUE
        if enc:
            out["z"] = torch.empty(B, self.z, dtype=f32, device=dev)
            out["mu"] = torch.empty(B, self.z, dtype=f32, device=dev)
            if self.latent_prior:
                out["logvar"] = torch.empty(B, self.z, dtype=f32, device=dev)
        if dec:
            out["action"] = torch.empty(B, self.da, dtype=f32, device=dev)
        if wld:
            out["future"] = torch.empty(B, self.dsb, dtype=f32, device=dev)
        if val:
            out["value"] = torch.empty(B, dtype=f32, device=dev)
        if z_in is not None:
            z_in = z_in.to(dev, f32).contiguous()
        if act_in is not None:
            act_in = act_in.to(dev, f32).contiguous()
        if eps is not None:
            eps = eps.to(dev, f32).contiguous()
        with torch.cuda.device(dev):
            _abi.check(self.lib.pvae_forward(
                self._h, int(parts), int(B), _ptr(obs), obs.shape[1], _ptr(z_in), _ptr(act_in),
                act_in.shape[1] if act_in is not None else 0, _ptr(eps), 1 if noise else 0, int(seed), int(offset),
                _ptr(out.get("action")), self.da, _ptr(out.get("mu")), _ptr(out.get("logvar")), _ptr(out.get("z")),
                _ptr(out.get("future")), _ptr(out.get("value")), _stream()))
        return out


def gemm_bf16(A, B, M, N, K, a_major=0, b_major=0, planes=1, splits=1):
    """Kernel-level entry: D[M,N] = A . B^T on the tcgen05 path (include/pvae_sm100.h, pvae_gemm_bf16)."""
    lib = _abi.load()
    D = torch.zeros(M, N, dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        _abi.check(lib.pvae_gemm_bf16(_ptr(A), a_major, _ptr(B), b_major, M, N, K, planes, splits, _ptr(D), _stream()))
    return D
