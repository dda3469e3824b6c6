"""CPU oracle: a plain PyTorch-fp32 restatement of the reference's algorithm for the PhysicsVAE training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under physicsvae_b200/ imports this module; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do, and only as the checker / the CPU arm being timed.

Pinning: the reference ships no tests for this path, so the oracle is pinned against the
reference's OWN code executed unchanged under oracle/ref_stub (oracle/make_golden.py -> tests/golden/*.npz, and
tests/test_oracle_vs_reference.py live in this container) and against the shipped checkpoint
data/pretrained/loco_modelV1.pt (golden numbers of SURVEY.md section 8c).

Every function cites the reference lines (relative to /root/reference) it restates.
"""
import math
import pickle

import numpy as np
import torch
import torch.nn.functional as F

NETS = ("_task_encoder", "_motor_decoder", "_world_model", "_value_branch")


# ---------------------------------------------------------------------------------------------------------------------
# layer specs, initialisation
# ---------------------------------------------------------------------------------------------------------------------
def gen_layers(width, depth, out_size="output", act_hidden="relu", act_out="linear"):
    """train_physics_vae.py:180-192 (without the unused softmax tail)."""
    assert depth > 0 and width > 0
    layers = [{"type": "fc", "hidden_size": width, "activation": act_hidden, "init_weight": {"name": "normc", "std": 1.0}}
              for _ in range(depth)]
    layers.append({"type": "fc", "hidden_size": out_size, "activation": act_out, "init_weight": {"name": "normc", "std": 0.01}})
    return layers


def activation(name, x, beta=None):
    """get_activation_fn, rllib_model_torch.py:30-46.  swish = ray 1.11's rllib Swish: x * sigmoid(beta * x), beta a trainable scalar
    initialised to 1.0 (ray/rllib/utils/torch_ops.py; restated in oracle/ref_stub)."""
    if name in ("linear", None):
        return x
    if name == "relu":
        return torch.relu(x)
    if name == "tanh":
        return torch.tanh(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "elu":
        return F.elu(x)
    if name in ("swish", "silu"):
        return x * torch.sigmoid(x if beta is None else beta * x)
    raise ValueError("Unknown activation ({})!".format(name))


def _init_fc(prefix, size_in, size_out, layers, params, acts):
    """FC.__init__ + SlimFC + normc_initializer (rllib_model_torch.py:243-272; ray 1.11 misc.py).  Consumes the torch RNG
    exactly like the reference: nn.Linear's own reset_parameters first, then normal_(0,1) rescaled per output row."""
    prev = size_in
    names = []
    for i, l in enumerate(layers):
        assert l["type"] == "fc"
        out = l["hidden_size"] if l["hidden_size"] != "output" else size_out
        lin = torch.nn.Linear(prev, out, bias=True)
        w = lin.weight.data
        info = l["init_weight"]
        if info["name"] == "normc":
            w.normal_(0, 1)
            w *= info["std"] / torch.sqrt(w.pow(2).sum(1, keepdim=True))
        elif info["name"] == "xavier_normal":
            torch.nn.init.xavier_normal_(w, gain=info["gain"])
        elif info["name"] == "xavier_uniform":
            torch.nn.init.xavier_uniform_(w, gain=info["gain"])
        else:
            raise NotImplementedError
        wk, bk = "%s._model.%d._model.0.weight" % (prefix, i), "%s._model.%d._model.0.bias" % (prefix, i)
        params[wk] = w.clone()
        params[bk] = torch.zeros(out)
        if l["activation"] in ("swish", "silu"):            # rllib's Swish module sits at index 1 of the SlimFC's Sequential
            params["%s._model.%d._model.1._beta" % (prefix, i)] = torch.tensor(1.0)
        names.append((wk, bk, l["activation"]))
        prev = out
    acts[prefix] = names


class OracleModel:
    """PhysicsVAE (rllib_model_torch.py:461-950) with the default wiring of the training CLI: encoder sees (body, task),
    decoder sees (body, z), constant log-std, no helper, prior normal_zero_mean_one_std or False."""

    def __init__(self, dim_state_body, dim_action, latent_dim=32, te_layers=None, md_layers=None, wm_layers=None,
                 vf_layers=None, latent_prior_type="normal_zero_mean_one_std", sample_std=0.1):
        self.dsb, self.da, self.z = dim_state_body, dim_action, latent_dim
        self.latent_prior_type = latent_prior_type
        if latent_prior_type not in ("normal_zero_mean_one_std", False):
            raise NotImplementedError("Unknown latent_prior_type:" + str(latent_prior_type))
        self.sample_std = sample_std
        self.latent_prior_noise = True
        te_layers = te_layers or gen_layers(256, 2)
        md_layers = md_layers or gen_layers(512, 3)
        wm_layers = wm_layers or gen_layers(1024, 2)
        vf_layers = vf_layers or gen_layers(256, 2)
        self.params, self.layers = {}, {}
        te_out = 2 * latent_dim if latent_prior_type else latent_dim
        # construction order = RNG order of the reference (rllib_model_torch.py:638-699)
        _init_fc("_task_encoder", 2 * dim_state_body, te_out, te_layers, self.params, self.layers)
        _init_fc("_motor_decoder", dim_state_body + latent_dim, dim_action, md_layers, self.params, self.layers)
        _init_fc("_world_model", dim_state_body + dim_action, dim_state_body, wm_layers, self.params, self.layers)
        _init_fc("_value_branch", 2 * dim_state_body, 1, vf_layers, self.params, self.layers)
        self.learnable = {"_task_encoder": True, "_motor_decoder": True, "_world_model": True, "_value_branch": True}
        self.cur = {}

    # state-dict interchange with the reference / the product
    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.params.items()}

    def load_state_dict(self, sd):
        assert set(sd.keys()) == set(self.params.keys()), (sorted(set(sd) ^ set(self.params)))
        for k in self.params:
            assert tuple(sd[k].shape) == tuple(self.params[k].shape), k
            self.params[k] = sd[k].detach().clone().float()

    def set_learnable(self, net, flag):
        """set_learnable_* (rllib_model_torch.py:930-950)."""
        self.learnable[net] = flag

    def _fc(self, net, x):
        """FC.forward (rllib_model_torch.py:274-275): Linear + bias + activation per layer."""
        for wk, bk, act in self.layers[net]:
            x = activation(act, F.linear(x, self.params[wk], self.params[bk]), self.params.get(wk.replace("._model.0.weight", "._model.1._beta")))
        return x

    def forward_encoder(self, obs, eps=None):
        """forward_encoder + _reparameterize (rllib_model_torch.py:773-820, 734-740).  eps replaces torch.randn_like."""
        z_body = obs[..., :self.dsb]
        h = self._fc("_task_encoder", obs)
        if self.latent_prior_type:
            mu, logvar = h[..., :self.z], h[..., self.z:]
            if self.latent_prior_noise:
                if eps is None:
                    eps = torch.randn_like(mu)
                z_task = mu + eps * torch.exp(0.5 * logvar)
            else:
                z_task = mu
            self.cur["mu"], self.cur["logvar"] = mu, logvar
        else:
            z_task = h
        return z_body, z_task

    def forward_decoder(self, z_body, z_task):
        """forward_decoder (rllib_model_torch.py:822-837) incl. AppendLogStd constant half (:194-206)."""
        a = self._fc("_motor_decoder", torch.cat([z_body, z_task], dim=-1))
        log_std = torch.full_like(a, math.log(self.sample_std))
        return torch.cat([a, log_std], dim=-1)

    def forward_world(self, obs, logits):
        """forward_world (rllib_model_torch.py:839-844)."""
        return self._fc("_world_model", torch.cat([obs[..., :self.dsb], logits[..., :self.da]], dim=-1))

    def forward_value(self, obs):
        """forward_value_branch (rllib_model_torch.py:846-853)."""
        return self._fc("_value_branch", obs)

    def forward(self, obs, eps=None):
        """PhysicsVAE.forward (rllib_model_torch.py:742-771)."""
        obs = obs.float()
        z_body, z_task = self.forward_encoder(obs, eps)
        logits = self.forward_decoder(z_body, z_task)
        future = self.forward_world(obs, logits)
        val = self.forward_value(obs)
        self.cur.update(z_task=z_task, z_body=z_body, future=future, value=val.squeeze(1))
        return logits


# ---------------------------------------------------------------------------------------------------------------------
# loss (train_physics_vae.py:361-435, lookahead == 1)
# ---------------------------------------------------------------------------------------------------------------------
def compute_loss(model, x, y, world, kl_coeff=1.0, a_rec_coeff=1.0, s_rec_coeff=None, cyc_coeff=1e-3, eps=None):
    """x: [B, 2*dsb] (s_t | s_{t+1}), y: [B, da].  Returns (total, dict of parts).
    world=True : coefficients (a, kl, s, cyc) = (0, 0, 1, 0)            read_loss_fn_coeff, train_physics_vae.py:330-335
    world=False: (a_rec_coeff, kl_coeff, world_model_s_rec_coeff=0, cyc_coeff)
    The reference's discarded full forward in the world phase (train_physics_vae.py:377-378) does not affect the value."""
    if world:
        a_c, kl_c, s_c, cyc_c = 0.0, 0.0, 1.0, 0.0
    else:
        a_c, kl_c, s_c, cyc_c = a_rec_coeff, kl_coeff, (0.0 if s_rec_coeff is None else s_rec_coeff), cyc_coeff
    dsb = model.dsb
    s1, s2_gt = x[..., :dsb], x[..., dsb:]
    parts = {"a": 0.0, "kl": 0.0, "s": 0.0, "cyc": 0.0}
    if not world:
        logits = model.forward(torch.cat([s1, s2_gt], dim=-1), eps=eps)
        y_t = logits[..., :logits.shape[1] // 2]                                   # compute_model, :356-359
        if a_c > 0.0:
            parts["a"] = F.mse_loss(y_t, y)                                          # :381-382 (nn.MSELoss, mean)
            if model.latent_prior_type and kl_c > 0.0:
                mu, logvar = model.cur["mu"], model.cur["logvar"]
                parts["kl"] = torch.mean(-0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1), dim=0)   # :384-389
    if s_c > 0:
        s2_pred = model.forward_world(s1, y)                                         # :412-414
        parts["s"] = F.mse_loss(s2_pred, s2_gt)
    if cyc_c > 0:
        parts["cyc"] = F.mse_loss(model.cur["future"], s2_gt)                        # :417-419
    total = a_c * parts["a"] + kl_c * parts["kl"] + s_c * parts["s"] + cyc_c * parts["cyc"]   # :430-434
    return total, parts


def compute_loss_lookahead(model, x, y, world, kl_coeff=1.0, a_rec_coeff=1.0, s_rec_coeff=None, cyc_coeff=1e-3, eps=None):
    """The general form of compute_loss (train_physics_vae.py:361-435) for `lookahead` L >= 1: x [B, L, 2*dsb], y [B, L, da],
    eps [L, B, z] or None.  An autoregressive rollout: the body state of step t + 1 is the world model's prediction from the
    DECODED action of step t (`s1 = self.model._cur_future_state`, :421), so the full forward runs in both phases and gradients
    flow through time; the four terms are averaged over the L steps (:423-428).  The CLI hard-wires L = 1 (:277), for which this
    equals compute_loss; the CUDA path does not implement L > 1 yet -- this pins what it will have to match."""
    if world:
        a_c, kl_c, s_c, cyc_c = 0.0, 0.0, 1.0, 0.0
    else:
        a_c, kl_c, s_c, cyc_c = a_rec_coeff, kl_coeff, (0.0 if s_rec_coeff is None else s_rec_coeff), cyc_coeff
    dsb, L = model.dsb, x.shape[1]
    parts = {"a": 0.0, "kl": 0.0, "s": 0.0, "cyc": 0.0}
    s1 = x[:, 0, :dsb]                                                               # :365
    for t in range(L):
        s2_gt, y_gt = x[:, t, dsb:], y[:, t, :]                                      # :369-375
        logits = model.forward(torch.cat([s1, s2_gt], dim=-1), eps=None if eps is None else eps[t])   # :377-378, always
        y_t = logits[..., :logits.shape[1] // 2]
        if a_c > 0.0:
            parts["a"] = parts["a"] + F.mse_loss(y_t, y_gt)                          # :381-382
            if model.latent_prior_type and kl_c > 0.0:
                mu, logvar = model.cur["mu"], model.cur["logvar"]
                parts["kl"] = parts["kl"] + torch.mean(-0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1), dim=0)
        if s_c > 0:
            parts["s"] = parts["s"] + F.mse_loss(model.forward_world(s1, y_gt), s2_gt)    # :412-414
        if cyc_c > 0:
            parts["cyc"] = parts["cyc"] + F.mse_loss(model.cur["future"], s2_gt)     # :417-419
        s1 = model.cur["future"]                                                     # :421
    if L > 1:
        parts = {k: v / float(L) for k, v in parts.items()}                          # :423-428
    total = a_c * parts["a"] + kl_c * parts["kl"] + s_c * parts["s"] + cyc_c * parts["cyc"]
    return total, parts


def loss_and_grads(model, x, y, world, **kw):
    """compute_loss + loss.backward() (torch_models.py:141-142) for the nets that are learnable in this phase.
    Returns (loss float, parts, {param name: grad}).  Frozen nets are differentiated through but get no grads.
    x with a lookahead axis ([B, L, 2*dsb]) selects compute_loss_lookahead."""
    train_nets = ["_world_model"] if world else ["_task_encoder", "_motor_decoder"]
    saved = model.params
    leaf = {}
    for k, v in saved.items():
        t = v.detach().clone()
        t.requires_grad_(any(k.startswith(n + ".") for n in train_nets))
        leaf[k] = t
    model.params = leaf
    try:
        total, parts = (compute_loss_lookahead if x.dim() == 3 else compute_loss)(model, x, y, world, **kw)
        total.backward()
    finally:
        model.params = saved
    grads = {k: t.grad.detach().clone() for k, t in leaf.items() if t.grad is not None}
    parts = {k: float(v.detach()) if torch.is_tensor(v) else float(v) for k, v in parts.items()}
    return float(total.detach()), parts, grads


class OracleTrainer:
    """torch_models.TrainModel.setup/step + train_physics_vae.TrainModel.setup/step (torch_models.py:110-161;
    train_physics_vae.py:314-351): Adam(lr, betas (0.9, 0.999), eps 1e-8, wd 0) over all parameters, StepLR(50, 0.7)
    stepped once per epoch, sequential mini-batches with a short last batch, two-phase schedule."""

    def __init__(self, model, X, Y, batch_size=256, lr=5e-4, max_iter_world_model=0, kl_coeff=1.0, cyc_coeff=1e-3,
                 a_rec_coeff=1.0, step_size=50, gamma=0.70):
        self.model = model
        self.X = torch.as_tensor(np.asarray(X)).reshape(len(X), -1).float()       # DatasetBase.__getitem__: torch.Tensor(x)
        self.Y = torch.as_tensor(np.asarray(Y)).reshape(len(Y), -1).float()
        self.bs = batch_size
        self.tensors = {k: v.detach().clone().requires_grad_(True) for k, v in model.params.items()}
        self.opt = torch.optim.Adam(list(self.tensors.values()), lr=lr, weight_decay=0.0)
        self.sched = torch.optim.lr_scheduler.StepLR(self.opt, step_size=step_size, gamma=gamma)
        self.iter = 0
        self.max_iter_world_model = max_iter_world_model
        self.kl, self.cyc, self.a = kl_coeff, cyc_coeff, a_rec_coeff
        self.world = True
        self._freeze(["_world_model"])

    def _freeze(self, train_nets):
        for k, t in self.tensors.items():
            t.requires_grad_(any(k.startswith(n + ".") for n in train_nets))

    def batches(self):
        n = len(self.X)
        return [(i, min(i + self.bs, n)) for i in range(0, n, self.bs)]

    def step(self, eps_fn=None):
        if self.iter == self.max_iter_world_model:                                    # train_physics_vae.py:342-350
            self.world = False
            self._freeze(["_task_encoder", "_motor_decoder"])
        self.iter += 1                                                                 # torch_models.py:132
        mean_loss = 0.0
        bl = self.batches()
        for bi, (lo, hi) in enumerate(bl):
            x, y = self.X[lo:hi], self.Y[lo:hi]
            self.opt.zero_grad()
            self.model.params = self.tensors
            eps = eps_fn(self.iter, bi, hi - lo) if (eps_fn and not self.world) else None
            loss, _ = compute_loss(self.model, x, y, self.world, kl_coeff=self.kl, cyc_coeff=self.cyc, a_rec_coeff=self.a, eps=eps)
            loss.backward()
            self.opt.step()
            mean_loss += loss.item()
        mean_loss /= len(bl)                                                           # unweighted mean of batch means
        self.sched.step()
        self.model.params = {k: v.detach() for k, v in self.tensors.items()}
        return {"mean_train_loss": mean_loss, "mean_test_loss": 0.0}


# ---------------------------------------------------------------------------------------------------------------------
# dataset (train_physics_vae.py:94-164; torch_models.py:39-68)
# ---------------------------------------------------------------------------------------------------------------------
def merge_dataset(files):
    """merge_dataset, train_physics_vae.py:94-114."""
    data_all = None
    for i, file in enumerate(files):
        with open(file, "rb") as f:
            data = pickle.load(f)
        if i == 0:
            data_all = data
        else:
            for key in ("iter_per_episode", "dim_state", "dim_state_body", "dim_state_task", "dim_action", "exp_std"):
                assert data_all[key] == data[key]
            data_all["episodes"] = data_all["episodes"] + data["episodes"]
    return data_all


def build_transitions(episodes, num_samples=None, lookahead=1, cond="abs", use_a_gt=False):
    """load_dataset_for_PhysicsVAE's loop (train_physics_vae.py:133-156), literally: one (x, y) per (episode, i)."""
    X, Y = [], []
    assert lookahead >= 1
    for ep in episodes:
        num_tuples = len(ep["time"])
        assert num_tuples >= lookahead
        for i in range(num_tuples - lookahead):
            if num_samples is not None and len(X) >= num_samples:
                break
            x, y = [], []
            for j in range(lookahead):
                s1 = np.asarray(ep["state_body"][i + j])
                s2 = np.asarray(ep["state_body"][i + j + 1])
                a = ep["action_gt"][i + j] if use_a_gt else ep["action"][i + j]
                if cond == "abs":
                    x.append(np.hstack([s1, s2]))
                elif cond == "rel":
                    x.append(np.hstack([s1, s2 - s1]))
                else:
                    raise NotImplementedError
                y.append(a)
            X.append(np.vstack(x))
            Y.append(np.vstack(y))
    return np.array(X), np.array(Y)


def synthetic_episodes(n_episodes, T, dsb, da, seed=0):
    """Synthetic expert demonstrations in the README pickle format (README.md:82-117), SURVEY.md section 8d:
    s_0 ~ N(0,1), s_{t+1} = s_t + 0.05 N(0,1), a_t ~ U(-1,1); float64 states, float32 actions."""
    rng = np.random.default_rng(seed)
    episodes = []
    for _ in range(n_episodes):
        s = np.empty((T, dsb), dtype=np.float64)
        s[0] = rng.standard_normal(dsb)
        s[1:] = s[0] + np.cumsum(0.05 * rng.standard_normal((T - 1, dsb)), axis=0)
        a = rng.uniform(-1, 1, size=(T, da)).astype(np.float32)
        episodes.append({
            "time": [t / 30.0 for t in range(T)],
            "state": [np.hstack([s[t], s[min(t + 1, T - 1)]]) for t in range(T)],
            "action": [a[t] for t in range(T)],
            "action_gt": [a[t] for t in range(T)],
            "reward": [0.0] * T,
            "state_body": [s[t] for t in range(T)],
            "state_task": [s[min(t + 1, T - 1)] for t in range(T)],
        })
    return {"iter_per_episode": 1, "dim_state": 2 * dsb, "dim_state_body": dsb, "dim_state_task": dsb, "dim_action": da,
            "exp_std": 0.05, "episodes": episodes}


def flops_per_transition(dsb, da, z, te, md, wm):
    """Algorithmic FLOPs per transition (2*MAC, unpadded dims), SURVEY.md section 8d / BASELINE.md section 4."""
    def M(i, hidden, o):
        dims = [i] + list(hidden) + [o]
        return sum(dims[k] * dims[k + 1] for k in range(len(dims) - 1)), dims
    mte, dte = M(2 * dsb, te, 2 * z)
    mmd, dmd = M(dsb + z, md, da)
    mwm, dwm = M(dsb + da, wm, dsb)
    m1 = lambda m, d: m - d[0] * d[1]
    world = 2 * (2 * mwm + m1(mwm, dwm))
    vae = 2 * (mte + mmd + mwm) + 2 * (mte + m1(mte, dte)) + 2 * (mmd + m1(mmd, dmd) + z * dmd[1]) + 2 * (m1(mwm, dwm) + da * dwm[1])
    return world, vae
