"""Host-side mirror of the reference's model zoo for the hot path: `FC`, `AppendLogStd`, `PhysicsVAE`
(reference: rllib_model_torch.py:30-46, 160-282, 461-950).  Same constructor signatures, attribute names and state-dict
keys (`_task_encoder._model.{i}._model.0.{weight,bias}` ...), so checkpoints interchange with the reference class.

The modules only HOLD the fp32 master parameters.  All arithmetic -- Linear+activation chains, reparameterisation, the
log-std append is bookkeeping -- runs in libpvae_sm100.so through `physicsvae_b200.engine.Engine`; calling forward on CPU
tensors raises, there is no PyTorch fallback path.

ray / gym are optional: when `ray.rllib` is importable the classes subclass the real TorchModelV2 and register with
ModelCatalog exactly like the reference (rllib_model_torch.py:952-953); otherwise local stand-ins keep the call protocol
(`model(input_dict=..., state=..., seq_lens=...)` -> `(logits, state)`, the ModelV2.__call__ dispatch the reference
relies on, train_physics_vae.py:356-359).
"""
import logging
import os
import weakref

import numpy as np
import torch
import torch.nn as nn

from . import _abi
from .engine import Engine, NET_NAMES

logger = logging.getLogger(__name__)

try:  # pragma: no cover - ray is not installed in the build image
    from ray.rllib.models.torch.torch_modelv2 import TorchModelV2
    from ray.rllib.models import ModelCatalog
    HAVE_RLLIB = True
except Exception:  # noqa
    HAVE_RLLIB = False

    class TorchModelV2(object):
        """Minimal TorchModelV2/ModelV2: stores ctor args; __call__ routes to forward() like ray 1.11's ModelV2.__call__."""

        def __init__(self, obs_space, action_space, num_outputs, model_config, name):
            self.obs_space = obs_space
            self.action_space = action_space
            self.num_outputs = num_outputs
            self.model_config = model_config
            self.name = name or "default_model"
            self.framework = "torch"

        def __call__(self, input_dict, state=None, seq_lens=None):
            restored = dict(input_dict)
            restored["obs_flat"] = input_dict["obs"]
            outputs, state_out = self.forward(restored, state or [], seq_lens)
            return outputs, state_out if len(state_out) > 0 else (state or [])

    class ModelCatalog(object):
        _registry = {}

        @staticmethod
        def register_custom_model(name, cls):
            ModelCatalog._registry[name] = cls


class Swish(nn.Module):
    """ray.rllib.utils.torch_ops.Swish (ray 1.11, what get_activation_fn("swish") returns, rllib_model_torch.py:33-35):
    x * sigmoid(beta * x) with `_beta` a trainable parameter initialised to 1.0 (state-dict key `..._model.{i}._model.1._beta`).
    Inside a PhysicsVAE the engine reads beta from the device and deposits its gradient like any other parameter's
    (pvae_bind_act_params); a stand-alone FC evaluates swish with beta = 1."""

    def __init__(self):
        super().__init__()
        self._beta = nn.Parameter(torch.tensor(1.0))

    def forward(self, x):
        return x * torch.sigmoid(self._beta * x)


def get_activation_fn(name=None):
    """Activation registry (rllib_model_torch.py:30-46).  Returns the nn.Module class (None for linear); the class is
    only a marker here -- the engine applies the activation inside the GEMM epilogue."""
    if name in ["linear", None]:
        return None
    if name in ["swish", "silu"]:
        return Swish
    if name == "relu":
        return nn.ReLU
    if name == "tanh":
        return nn.Tanh
    if name == "sigmoid":
        return nn.Sigmoid
    if name == "elu":
        return nn.ELU
    raise ValueError("Unknown activation ({})!".format(name))


def normc_initializer(std=1.0):
    """ray.rllib.models.torch.misc.normc_initializer: N(0,1) rows rescaled to L2 norm `std`."""
    def initializer(tensor):
        tensor.data.normal_(0, 1)
        tensor.data *= std / torch.sqrt(tensor.data.pow(2).sum(1, keepdim=True))
    return initializer


def get_initializer(info):
    """rllib_model_torch.py:220-232."""
    if info["name"] == "normc":
        return normc_initializer(info["std"])
    elif info["name"] == "xavier_normal":
        def initializer(tensor):
            return nn.init.xavier_normal_(tensor, gain=info["gain"])
        return initializer
    elif info["name"] == "xavier_uniform":
        def initializer(tensor):
            return nn.init.xavier_uniform_(tensor, gain=info["gain"])
        return initializer
    else:
        raise NotImplementedError


class SlimFC(nn.Module):
    """ray.rllib.models.torch.misc.SlimFC: `_model = Sequential(Linear[, act])`; weight <- initializer, bias <- 0.
    Kept for the state-dict nesting (`._model.0.weight`); `activation` is the name handed to the engine."""

    def __init__(self, in_size, out_size, initializer=None, activation_fn=None, activation=None):
        super().__init__()
        layers = []
        linear = nn.Linear(in_size, out_size, bias=True)
        if initializer is None:
            initializer = nn.init.xavier_uniform_
        initializer(linear.weight)
        nn.init.constant_(linear.bias, 0.0)
        layers.append(linear)
        if activation_fn is not None:
            layers.append(activation_fn())
        self._model = nn.Sequential(*layers)
        self.activation = activation if activation is not None else "linear"

    @property
    def linear(self):
        return self._model[0]


class AppendLogStd(nn.Module):
    """rllib_model_torch.py:160-206: concatenates a constant (plain tensor, not in the state dict) or learnable log-std."""

    def __init__(self, type, init_val, dim):
        super().__init__()
        self.type = type
        if np.isscalar(init_val):
            init_val = init_val * np.ones(dim)
        elif isinstance(init_val, (np.ndarray, list)):
            assert len(init_val) == dim
        else:
            raise NotImplementedError
        self.init_val = init_val
        if self.type == "constant":
            self.log_std = torch.Tensor(init_val)
        elif self.type == "state_independent":
            self.log_std = torch.nn.Parameter(torch.Tensor(init_val))
            self.register_parameter("log_std", self.log_std)
        else:
            raise NotImplementedError

    def set_val(self, val):
        assert self.type == "constant", "Change value is only allowed in constant logstd"
        assert np.isscalar(val), "Only scalar is currently supported"
        self.log_std[:] = val

    def forward(self, x):
        assert x.shape[-1] == self.log_std.shape[-1]
        log_std = self.log_std.to(x.device, x.dtype).reshape([1] * (x.dim() - 1) + [-1]).expand(*x.shape[:-1], -1)
        return torch.cat([x, log_std], axis=-1)


class FC(nn.Module):
    """A network with fully connected layers (rllib_model_torch.py:234-282).  Layer-spec DSL as in the reference:
    [{"type": "fc", "hidden_size": int | "output", "activation": str, "init_weight": {"name": "normc", "std": float}}, ...]"""

    def __init__(self, size_in, size_out, layers, append_log_std=False, log_std_type="constant", sample_std=1.0):
        super().__init__()
        nn_layers = []
        prev_layer_size = size_in
        for l in layers:
            layer_type = l["type"]
            if layer_type == "fc":
                assert isinstance(l["hidden_size"], int) or l["hidden_size"] == "output"
                hidden_size = l["hidden_size"] if l["hidden_size"] != "output" else size_out
                layer = SlimFC(in_size=prev_layer_size, out_size=hidden_size, initializer=get_initializer(l["init_weight"]),
                               activation_fn=get_activation_fn(l["activation"]), activation=l["activation"])
                prev_layer_size = hidden_size
            else:
                # bn / softmax / hardmax layers are never instantiated by train_physics_vae.py
                raise NotImplementedError("Unknown Layer Type:", layer_type)
            nn_layers.append(layer)
        if append_log_std:
            nn_layers.append(AppendLogStd(type=log_std_type, init_val=np.log(sample_std), dim=size_out))
        self._model = nn.Sequential(*nn_layers)
        self.size_in, self.size_out = size_in, size_out

    def fc_layers(self):
        return [m for m in self._model if isinstance(m, SlimFC)]

    def layer_spec(self):
        """[(out_features, activation name)] for the engine."""
        return [(m.linear.out_features, m.activation) for m in self.fc_layers()]

    # ---- forward (rllib_model_torch.py:274-275: `return self._model(x)`) --------------------------------------------------
    def _bind_owner(self, owner, role):
        """A sub-net of a PhysicsVAE runs on its owner's engine (same parameters, same shadow operands)."""
        object.__setattr__(self, "_owner_ref", weakref.ref(owner))
        object.__setattr__(self, "_role", role)

    def _own_engine(self, batch):
        """Stand-alone FC (no PhysicsVAE around it): a private one-net engine whose only net takes `size_in` columns."""
        eng = getattr(self, "_engine", None)
        p0 = next(self.parameters())
        if p0.device.type != "cuda":
            raise _abi.PvaeError("FC runs on the sm_100a engine only: move the module to a CUDA device first (no CPU / eager fallback exists)")
        if eng is None or eng.max_batch < batch or eng.device != p0.device:
            if eng is not None:
                eng.close()
            eng = Engine(8, 8, 8, {"world_model": self.layer_spec()}, precision=getattr(self, "engine_precision", "bf16x3"),
                         max_batch=max(int(batch), 256), device=p0.device.index, in_dims={"world_model": (self.size_in, 0)})
            layers = self.fc_layers()
            flat = torch.empty(eng.grad_elems("world_model"), dtype=torch.float32, device=p0.device)
            off, Ws, bs = 0, [], []
            for m in layers:
                for p in (m.linear.weight, m.linear.bias):
                    k = p.numel()
                    view = flat[off:off + k].view_as(p)
                    view.copy_(p.data)
                    p.data = view
                    off += k
                Ws.append(m.linear.weight.data)
                bs.append(m.linear.bias.data)
            eng.bind_net("world_model", Ws, bs, None)
            object.__setattr__(self, "_engine", eng)
            object.__setattr__(self, "_synced_version", None)
        version = tuple(p._version for p in self.parameters())
        if self._synced_version != version:           # parameters were modified in place since the last call
            eng.sync_weights(["world_model"])
            object.__setattr__(self, "_synced_version", version)
        return eng

    def _apply(self, fn, *a, **kw):
        eng = getattr(self, "_engine", None)
        if eng is not None:                            # .to() / .cuda() re-create parameter storage
            eng.close()
            object.__setattr__(self, "_engine", None)
        return super()._apply(fn, *a, **kw)

    def forward(self, x):
        lead = x.shape[:-1]
        x2 = x.float().reshape(-1, x.shape[-1])
        if x2.shape[-1] != self.size_in:
            raise ValueError("FC expects %d input features, got %d" % (self.size_in, x2.shape[-1]))
        owner = getattr(self, "_owner_ref", None)
        owner = owner() if owner is not None else None
        if owner is not None:
            out = owner._ready(x2.shape[0]).fc_forward(self._role, x2)
        else:
            out = self._own_engine(x2.shape[0]).fc_forward("world_model", x2)
        out = out.reshape(*lead, -1)
        tail = self._model[-1]
        return tail(out) if isinstance(tail, AppendLogStd) else out

    def save_weights(self, file):
        torch.save(self.state_dict(), file)

    def load_weights(self, file):
        self.load_state_dict(torch.load(file))
        self.eval()


def _fc_spec(width, depth=2):
    return [{"type": "fc", "hidden_size": width, "activation": "relu", "init_weight": {"name": "normc", "std": 1.0}}
            for _ in range(depth)] + \
           [{"type": "fc", "hidden_size": "output", "activation": "linear", "init_weight": {"name": "normc", "std": 0.01}}]


DEFAULT_FC_64X2 = _fc_spec(64)
DEFAULT_FC_128X2 = _fc_spec(128)
DEFAULT_FC_256X2 = _fc_spec(256)
DEFAULT_FC_512X2 = _fc_spec(512)
DEFAULT_FC_512X3 = _fc_spec(512, 3)
DEFAULT_FC_1024X2 = _fc_spec(1024)

_NET_ATTR = {"task_encoder": "_task_encoder", "motor_decoder": "_motor_decoder", "world_model": "_world_model",
             "value_branch": "_value_branch"}


class PhysicsVAE(TorchModelV2, nn.Module):
    """Conditional VAE + world model (rllib_model_torch.py:461-950): task encoder -> reparameterised latent -> motor
    decoder, world model on (s_body, action), value branch.  Same ctor / forward protocol as the RLlib custom model."""

    DEFAULT_CONFIG = {
        "project_dir": None,
        "log_std_type": "constant",
        "sample_std": 0.1,
        "load_weights": None,
        "task_encoder_inputs": ["body", "task"],
        "task_encoder_layers": DEFAULT_FC_256X2,
        "task_encoder_load_weights": None,
        "task_encoder_learnable": True,
        "task_encoder_output_dim": 32,
        "latent_prior_type": "normal_zero_mean_one_std",
        "latent_prior_layers": None,
        "motor_decoder_inputs": ["body", "task"],
        "motor_decoder_layers": DEFAULT_FC_512X3,
        "motor_decoder_load_weights": None,
        "motor_decoder_learnable": True,
        "motor_decoder_helper_enable": False,
        "motor_decoder_helper_layers": None,
        "motor_decoder_helper_load_weights": None,
        "motor_decoder_helper_learnable": True,
        "motor_decoder_helper_range": 0.5,
        "value_fn_layers": DEFAULT_FC_256X2,
        "world_model_layers": DEFAULT_FC_1024X2,
        "world_model_load_weights": None,
        "world_model_learnable": True,
        "observation_space": None,
        "observation_space_body": None,
        "observation_space_task": None,
        "action_space": None,
        # engine knobs (not in the reference): arithmetic of the tensor-core contractions and workspace capacity
        "engine_precision": "bf16x3",
        "engine_max_batch": 4096,
    }

    def __init__(self, obs_space, action_space, num_outputs, model_config, name, **model_kwargs):
        TorchModelV2.__init__(self, obs_space, action_space, num_outputs, model_config, name)
        nn.Module.__init__(self)

        assert num_outputs % 2 == 0, ("num_outputs must be divisible by two", num_outputs)
        num_outputs = num_outputs // 2

        custom_model_config = PhysicsVAE.DEFAULT_CONFIG.copy()
        custom_model_config_by_user = model_config.get("custom_model_config")
        if custom_model_config_by_user:
            custom_model_config.update(custom_model_config_by_user)
        cfg = custom_model_config

        log_std_type = cfg.get("log_std_type")
        assert log_std_type in ["constant", "state_independent"]
        sample_std = cfg.get("sample_std")
        assert sample_std > 0.0, "The value shoulde be positive"

        project_dir = cfg.get("project_dir")
        paths = {}
        for key in ("load_weights", "task_encoder_load_weights", "motor_decoder_load_weights", "world_model_load_weights"):
            p = cfg.get(key)
            if p and project_dir:
                p = os.path.join(project_dir, p)
            paths[key] = p

        self._task_encoder_inputs = cfg.get("task_encoder_inputs")
        self._motor_decoder_inputs = cfg.get("motor_decoder_inputs")
        # the B200 engine implements the wiring the training CLI uses (train_physics_vae.py:236-286)
        if list(self._task_encoder_inputs) != ["body", "task"] or list(self._motor_decoder_inputs) != ["body", "task"]:
            raise NotImplementedError("only task_encoder_inputs = motor_decoder_inputs = ['body', 'task'] is supported")
        if cfg.get("motor_decoder_helper_enable"):
            raise NotImplementedError("motor_decoder_helper is disabled on the training path (rllib_model_torch.py:490)")
        latent_prior_type = cfg.get("latent_prior_type")
        self._latent_prior_type = latent_prior_type
        self._task_encoder_output_dim = task_encoder_output_dim = cfg.get("task_encoder_output_dim")

        self.dim_state_body = int(np.prod(cfg.get("observation_space_body").shape))
        self.dim_state_task = int(np.prod(cfg.get("observation_space_task").shape))
        self.dim_state = int(np.prod(obs_space.shape))
        self.dim_action = int(np.prod(action_space.shape))
        assert self.dim_state == self.dim_state_body + self.dim_state_task
        if self.dim_state_task != self.dim_state_body:
            raise NotImplementedError("the hot path uses s_task := next s_body (train_physics_vae.py:198-214)")

        if latent_prior_type in ["normal_zero_mean_one_std"]:
            size_out_task_encoder = 2 * task_encoder_output_dim
        elif latent_prior_type == False:  # noqa: E712  (the reference compares with == False)
            size_out_task_encoder = task_encoder_output_dim
        elif latent_prior_type in ["normal_state_mean_one_std", "hypersphere_uniform"]:
            # both are broken upstream (constructor TypeError / reads an attribute that is never set)
            raise NotImplementedError("latent_prior_type %r does not work in the reference either" % (latent_prior_type,))
        else:
            raise NotImplementedError("Unknown latent_prior_type:" + str(latent_prior_type))
        self._latent_prior = None

        # construction order = the reference's, so that a seeded init draws the same weights (rllib_model_torch.py:638-699)
        self._task_encoder = FC(size_in=self.dim_state, size_out=size_out_task_encoder, layers=cfg.get("task_encoder_layers"))
        self._motor_decoder = FC(size_in=self.dim_state_body + task_encoder_output_dim, size_out=num_outputs,
                                 layers=cfg.get("motor_decoder_layers"), append_log_std=True, log_std_type=log_std_type,
                                 sample_std=sample_std)
        self._motor_decoder_helper = None
        self._world_model = FC(size_in=self.dim_action + self.dim_state_body, size_out=self.dim_state_body,
                               layers=cfg.get("world_model_layers"))
        self._value_branch = FC(size_in=self.dim_state, size_out=1, layers=cfg.get("value_fn_layers"))
        for role, attr in _NET_ATTR.items():
            getattr(self, attr)._bind_owner(self, role)

        self._cur_value = None
        self._cur_task_encoder_variable = None
        self._cur_body_encoder_variable = None
        self._cur_task_encoder_mu = None
        self._cur_task_encoder_logvar = None
        self._cur_future_state = None
        self.latent_prior_noise = True

        self._engine = None
        self._engine_precision = cfg.get("engine_precision")
        self._engine_max_batch = int(cfg.get("engine_max_batch"))
        self._flat = {}         # net name -> (flat params, flat grads)
        self._weights_dirty = True
        self._noise_offset = 0

        if paths["load_weights"]:
            self.load_weights(paths["load_weights"])
            print("load_weights:", paths["load_weights"])
        if paths["task_encoder_load_weights"]:
            self.load_weights_task_encoder(paths["task_encoder_load_weights"])
            self.set_learnable_task_encoder(cfg.get("task_encoder_learnable"))
        if paths["motor_decoder_load_weights"]:
            self.load_weights_motor_decoder(paths["motor_decoder_load_weights"])
            self.set_learnable_motor_decoder(cfg.get("motor_decoder_learnable"))
        if paths["world_model_load_weights"]:
            self.load_weights_world_model(paths["world_model_load_weights"])
            self.set_learnable_world_model(cfg.get("world_model_learnable"))

    # nn.Module.__call__ would bypass the dict protocol; the reference's MRO puts ModelV2.__call__ first
    def __call__(self, input_dict, state=None, seq_lens=None):
        return TorchModelV2.__call__(self, input_dict, state, seq_lens)

    def get_initial_state(self):
        return []

    # ---- engine plumbing ------------------------------------------------------------------------------------------
    def net(self, name):
        return getattr(self, _NET_ATTR[name])

    def engine(self, max_batch=None, precision=None):
        """The sm_100a engine bound to this module's parameters (created on first use; the module must be on CUDA).
        Parameters are re-pointed at views of one flat fp32 buffer per net ([W0|b0|W1|b1|...], the layout of the
        library's gradient buffers) so that Adam and the data-parallel all-reduce see one tensor per net."""
        want_batch = int(max_batch or self._engine_max_batch)
        want_prec = precision or self._engine_precision
        if self._engine is not None and self._engine.max_batch >= want_batch and self._engine.precision == want_prec:
            return self._engine
        p0 = next(self.parameters())
        if p0.device.type != "cuda":
            raise _abi.PvaeError("PhysicsVAE runs on the sm_100a engine only: move the module to a CUDA device first "
                                 "(no CPU / eager fallback exists)")
        if self._engine is not None:
            self._engine.close()
        spec = {name: self.net(name).layer_spec() for name in NET_NAMES}
        eng = Engine(self.dim_state_body, self.dim_action, self._task_encoder_output_dim, spec,
                     latent_prior=bool(self._latent_prior_type), precision=want_prec, max_batch=want_batch, device=p0.device.index)
        # One gradient pool for the whole model, laid out [motor decoder | task encoder | loss slots | world model | value branch]
        # (every piece 16-byte aligned): what a training step has to exchange between data-parallel ranks -- the gradients of
        # the nets it trains plus the loss slots -- is ONE contiguous range in either phase (reduce_range()), and so are the part
        # that is complete early in the backward pass and the rest (reduce_ranges()).
        # The "scalar block" between encoder and world model holds the loss slots and, per net, the gradients of the activations'
        # parameters (swish beta, one slot per layer): it is part of the exchanged range in both phases.
        al = lambda k: (k + 3) // 4 * 4
        order = ("motor_decoder", "task_encoder", "world_model", "value_branch")
        sizes = {name: eng.grad_elems(name) for name in NET_NAMES}
        SCALARS = _abi.PVAE_LOSS_SLOTS + 3 * _abi.PVAE_MAX_LAYERS
        offs, cur = {}, 0
        for name in order:
            offs[name] = cur
            cur += al(sizes[name])
            if name == "task_encoder":
                loss_off = cur
                cur += al(SCALARS)
        pool = self._grad_pool = self._pool_alloc(cur, p0.device)
        eng.loss = pool[loss_off:loss_off + _abi.PVAE_LOSS_SLOTS]
        self._pool_off = dict(offs, loss=loss_off, end=cur, scalars=SCALARS)
        self._pool_sizes = sizes
        # swish layers: beta values of a net live in one small tensor (entry l = layer l), their gradients in the scalar block
        self._betas = {}
        for i, name in enumerate(("task_encoder", "motor_decoder", "world_model")):
            acts = [m._model[1] if len(m._model) > 1 else None for m in self.net(name).fc_layers()]
            if not any(isinstance(a, Swish) for a in acts):
                continue
            beta = torch.ones(_abi.PVAE_MAX_LAYERS, dtype=torch.float32, device=p0.device)
            g0 = loss_off + _abi.PVAE_LOSS_SLOTS + i * _abi.PVAE_MAX_LAYERS
            dbeta = pool[g0:g0 + _abi.PVAE_MAX_LAYERS]
            for l, a in enumerate(acts):
                if isinstance(a, Swish):
                    beta[l] = a._beta.data.to(p0.device)
                    a._beta.data = beta[l]
                    a._beta._pvae_grad_view = dbeta[l]
            eng.bind_act_params(name, beta, dbeta)
            self._betas[name] = (beta, dbeta)
        for name in NET_NAMES:
            layers = self.net(name).fc_layers()
            n = sizes[name]
            flat = torch.empty(n, dtype=torch.float32, device=p0.device)
            gflat = pool[offs[name]:offs[name] + n]
            off = 0
            Ws, bs = [], []
            for m in layers:
                for p in (m.linear.weight, m.linear.bias):
                    k = p.numel()
                    view = flat[off:off + k].view_as(p)
                    view.copy_(p.data)
                    p.data = view
                    p._pvae_grad_view = gflat[off:off + k].view_as(p)
                    off += k
                Ws.append(m.linear.weight.data)
                bs.append(m.linear.bias.data)
            eng.bind_net(name, Ws, bs, gflat if name != "value_branch" else None)
            self._flat[name] = (flat, gflat)
        self._engine = eng
        self._engine_max_batch, self._engine_precision = want_batch, want_prec
        self._weights_dirty = True
        self._attach_grads()
        return eng

    def _attach_grads(self):
        """Expose the library's gradient buffers as `.grad` of every learnable parameter (None for frozen ones, which
        is what makes torch.optim.Adam skip them exactly like the reference)."""
        if self._engine is None:
            return
        for name in NET_NAMES:
            for p in self.net(name).parameters():
                view = getattr(p, "_pvae_grad_view", None)
                p.grad = view if (p.requires_grad and view is not None and name != "value_branch") else None

    def _pool_alloc(self, n, device):
        """The gradient pool; `grad_pool_factory` (set by a trainer) may hand out memory registered for peer access
        (physicsvae_b200.parallel.SymmetricPool) instead of a plain tensor."""
        factory = getattr(self, "grad_pool_factory", None)
        if factory is not None:
            t = factory(n, device)
            t.zero_()
            return t
        return torch.zeros(n, dtype=torch.float32, device=device)

    def reduce_range(self, world_phase):
        """The contiguous slice of the gradient pool a data-parallel step all-reduces: [loss slots | world model] in the world
        phase, [motor decoder | task encoder | loss slots] in the VAE phase."""
        o, sz = self._pool_off, self._pool_sizes
        if world_phase:
            return self._grad_pool[o["loss"]:o["world_model"] + sz["world_model"]]
        return self._grad_pool[o["motor_decoder"]:o["loss"] + o["scalars"]]

    def reduce_ranges(self, world_phase):
        """reduce_range() split in two for the overlapped exchange (include/pvae_sm100.h, pvae_set_exchange): (early, late).
        early: complete while backward GEMMs are still running -- world phase: the world model's layers 1 .. L-1 (everything behind
        layer 0's [W0 | b0]); VAE phase: the motor decoder.  late: the rest plus the loss slots -- [loss | W0 | b0] / [task encoder |
        loss].  early is None when the split is not 16-byte aligned or the net has a single layer."""
        o, sz = self._pool_off, self._pool_sizes
        if world_phase:
            layers = self._world_model.fc_layers()
            if len(layers) < 2:
                return None, self.reduce_range(True)
            l0 = layers[0].linear
            cut = o["world_model"] + l0.weight.numel() + l0.bias.numel()
            end = o["world_model"] + sz["world_model"]
            if cut % 4 or cut >= end:
                return None, self.reduce_range(True)
            return self._grad_pool[cut:end], self._grad_pool[o["loss"]:cut]
        return (self._grad_pool[o["motor_decoder"]:o["motor_decoder"] + sz["motor_decoder"]],
                self._grad_pool[o["task_encoder"]:o["loss"] + o["scalars"]])

    def flat_params(self, name):
        return self._flat[name][0]

    def flat_grads(self, name):
        return self._flat[name][1]

    def mark_weights_dirty(self):
        """Call after changing parameters in place (optimizer.step, load_state_dict): the bf16 shadow operands are
        refreshed before the next engine call."""
        self._weights_dirty = True

    def sync_weights(self, names=None):
        eng = self.engine()
        eng.sync_weights(names)
        if names is None:
            self._weights_dirty = False

    def _ready(self, batch):
        eng = self.engine(max_batch=max(batch, self._engine_max_batch))
        if self._weights_dirty:
            self.sync_weights()
        return eng

    def _apply(self, fn, *a, **kw):
        # .to() / .cuda() / .float() re-create parameter storage: drop the engine binding, it is rebuilt lazily
        if self._engine is not None:
            self._engine.close()
            self._engine = None
            self._flat = {}
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, state_dict, strict=True):
        r = super().load_state_dict(state_dict, strict)
        self._weights_dirty = True
        return r

    # ---- forward API ----------------------------------------------------------------------------------------------
    def _reparameterize(self, mu, logvar):
        raise _abi.PvaeError("reparameterisation is fused into the engine (see forward_encoder)")

    def _next_noise(self, n):
        off = self._noise_offset
        self._noise_offset += 1
        return off

    def forward(self, input_dict, state, seq_lens):
        """PhysicsVAE.forward (rllib_model_torch.py:742-771): one fused engine call for encoder, decoder, world model
        and value branch."""
        obs = input_dict["obs_flat"].float()
        obs2 = obs.reshape(-1, obs.shape[-1])
        eng = self._ready(obs2.shape[0])
        eps = input_dict.get("eps") if isinstance(input_dict, dict) else None
        out = eng.forward(obs2, _abi.PART_ENCODER | _abi.PART_DECODER | _abi.PART_WORLD | _abi.PART_VALUE, eps=eps,
                          noise=bool(self.latent_prior_noise and self._latent_prior_type), seed=self._seed(),
                          offset=self._next_noise(obs2.shape[0]))
        lead = obs.shape[:-1]
        self._store_encoder(out, lead)
        logits = self._motor_decoder._model[-1](out["action"].reshape(*lead, -1))
        self._cur_body_encoder_variable = obs[..., :self.dim_state_body]
        self._cur_value = out["value"].reshape(*lead, 1).squeeze(1)
        self._cur_future_state = out["future"].reshape(*lead, -1)
        return logits, state

    def _seed(self):
        return int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF)

    def _store_encoder(self, out, lead):
        self._cur_task_encoder_variable = out["z"].reshape(*lead, -1)
        if self._latent_prior_type:
            self._cur_task_encoder_mu = out["mu"].reshape(*lead, -1)
            self._cur_task_encoder_logvar = out["logvar"].reshape(*lead, -1)

    def forward_encoder(self, obs, state, seq_lens, state_cnt, eps=None):
        """rllib_model_torch.py:773-820.  `eps` (optional, [B, z]) replaces the reference's torch.randn_like draw."""
        obs = obs.float()
        obs2 = obs.reshape(-1, obs.shape[-1])
        eng = self._ready(obs2.shape[0])
        out = eng.forward(obs2, _abi.PART_ENCODER, eps=eps, noise=bool(self.latent_prior_noise and self._latent_prior_type),
                          seed=self._seed(), offset=self._next_noise(obs2.shape[0]))
        lead = obs.shape[:-1]
        self._store_encoder(out, lead)
        return obs[..., :self.dim_state_body], self._cur_task_encoder_variable, state_cnt

    def forward_decoder(self, z_body, z_task, state, seq_lens, state_cnt):
        """rllib_model_torch.py:822-837: logits = AppendLogStd(MD(cat[z_body, z_task]))."""
        zb = z_body.float().reshape(-1, z_body.shape[-1])
        zt = z_task.float().reshape(-1, z_task.shape[-1])
        eng = self._ready(zb.shape[0])
        out = eng.forward(zb, _abi.PART_DECODER, z_in=zt)
        logits = self._motor_decoder._model[-1](out["action"].reshape(*z_body.shape[:-1], -1))
        return logits, state_cnt

    def forward_world(self, obs, logits):
        """rllib_model_torch.py:839-844: WM(cat[obs[:, :dsb], logits[:, :da]])."""
        o2 = obs.float().reshape(-1, obs.shape[-1])
        a2 = logits.float().reshape(-1, logits.shape[-1])
        eng = self._ready(o2.shape[0])
        out = eng.forward(o2, _abi.PART_WORLD, act_in=a2)
        return out["future"].reshape(*obs.shape[:-1], -1)

    def forward_value_branch(self, obs, state, seq_lens, state_cnt):
        """rllib_model_torch.py:846-853."""
        o2 = obs.float().reshape(-1, obs.shape[-1])
        eng = self._ready(o2.shape[0])
        out = eng.forward(o2, _abi.PART_VALUE)
        return out["value"].reshape(*obs.shape[:-1], 1), state_cnt

    def value_function(self):
        assert self._cur_value is not None, "must call forward() first"
        return self._cur_value

    def set_exploration_std(self, std):
        log_std = np.log(std)
        self._motor_decoder._model[-1].set_val(log_std)

    def task_encoder_variable(self):
        return self._cur_task_encoder_variable

    def body_encoder_variable(self):
        return self._cur_body_encoder_variable

    # ---- checkpoint interchange (rllib_model_torch.py:870-928) ----------------------------------------------------------
    def _cpu_sd(self, module):
        return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}

    def save_weights(self, file):
        torch.save(self._cpu_sd(self), file)

    def load_weights(self, file):
        self.load_state_dict(torch.load(file, map_location="cpu"))
        self.eval()

    def save_weights_task_encoder(self, file):
        torch.save({"task_encoder": self._cpu_sd(self._task_encoder)}, file)

    def load_weights_task_encoder(self, file):
        state_dict = torch.load(file, map_location="cpu")
        self._task_encoder.load_state_dict(state_dict["task_encoder"])
        self._task_encoder.eval()
        self._weights_dirty = True

    def save_weights_motor_decoder(self, file):
        if self._motor_decoder:
            torch.save(self._cpu_sd(self._motor_decoder), file)

    def load_weights_motor_decoder(self, file):
        if self._motor_decoder:
            # log_std entries of the file are ignored for valid exploration (rllib_model_torch.py:892-904)
            dict_weights_orig = self._motor_decoder.state_dict()
            dict_weights_loaded = torch.load(file, map_location="cpu")
            for key in dict_weights_loaded.keys():
                if "log_std" in key:
                    dict_weights_loaded[key] = dict_weights_orig[key]
            self._motor_decoder.load_state_dict(dict_weights_loaded)
            self._motor_decoder.eval()
            self._weights_dirty = True

    def save_weights_motor_decoder_helper(self, file):
        pass

    def save_weights_world_model(self, file):
        if self._world_model:
            torch.save(self._cpu_sd(self._world_model), file)

    def load_weights_world_model(self, file):
        if self._world_model:
            self._world_model.load_state_dict(torch.load(file, map_location="cpu"))
            self._world_model.eval()
            self._weights_dirty = True

    def save_weights_latent_prior(self, file):
        pass

    def load_weights_latent_prior(self, file):
        pass

    # ---- freezing (rllib_model_torch.py:930-950) --------------------------------------------------------------------
    def set_learnable_task_encoder(self, learnable):
        if self._task_encoder:
            for name, param in self._task_encoder.named_parameters():
                param.requires_grad = learnable
        self._attach_grads()

    def set_learnable_motor_decoder(self, learnable, free_log_std=True):
        if self._motor_decoder:
            for name, param in self._motor_decoder.named_parameters():
                param.requires_grad = learnable
                if "log_std" in name:
                    param.requires_grad = free_log_std
        self._attach_grads()

    def set_learnable_motor_decoder_helper(self, learnable):
        pass

    def set_learnable_world_model(self, learnable):
        if self._world_model:
            for name, param in self._world_model.named_parameters():
                param.requires_grad = learnable
        self._attach_grads()


ModelCatalog.register_custom_model("physics_vae", PhysicsVAE)
