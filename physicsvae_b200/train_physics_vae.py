"""Host-side mirror of the reference's training script (train_physics_vae.py:30-521): same CLI flags, dataset format,
layer-spec DSL, trainer config, two-phase schedule and checkpoint file set -- over the sm_100a engine.

Differences that are deliberate: the dataset is built by vectorised numpy instead of an O(N) hstack loop;
the discarded full-model forward of the world phase is not executed; `--output` works; Ray Tune's trial loop is replaced
by a local sweep over the grid points when ray is absent.
"""
import argparse
import copy
import itertools
import json
import os
import pickle

import numpy as np
import torch

from . import _abi, parallel
from . import torch_models
from . import rllib_model_torch as policy_models

try:  # pragma: no cover
    from gym.spaces import Box
except Exception:  # noqa
    class Box(object):
        """gym.spaces.Box stand-in: only .shape / .low / .high / .dtype are used (train_physics_vae.py:216-233)."""

        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape if shape is None else tuple(shape)
            self.dtype = np.dtype(dtype)

args = None      # the reference reads a module-global `args` in load_dataset (train_physics_vae.py:339, 471)


def grid_search(values):
    return {"grid_search": values}


def arg_parser():
    """Same flags and defaults as train_physics_vae.py:30-55 (list flags append to their defaults, appendix B.7)."""
    parser = argparse.ArgumentParser()
    parser.add_argument("--max_iter_world_model", type=int, default=0)
    parser.add_argument("--max_iter", type=int, default=100)
    parser.add_argument("--num_cpus", type=int, default=1)
    parser.add_argument("--num_gpus", type=int, default=0)
    parser.add_argument("--data_train", action="append", required=True, type=str, default=None)
    parser.add_argument("--data_test", action="append", type=str, default=None)
    parser.add_argument("--num_data", type=int, default=None)
    parser.add_argument("--output", type=str, default=None)
    parser.add_argument("--lr", type=float, default=0.0005)
    parser.add_argument("--lr_schedule", type=str, default="step")
    parser.add_argument("--batch_size", type=int, default=256)
    parser.add_argument("--checkpoint_freq", type=int, default=100)
    parser.add_argument("--checkpoint", type=str, default=None)
    parser.add_argument("--cluster", action="store_true")
    parser.add_argument("--resume", action="store_true")
    parser.add_argument("--name", type=str, default=None)
    parser.add_argument("--local_dir", type=str, default="~/ray_results")
    parser.add_argument("--world_model", type=str, default=None)
    parser.add_argument("--latent_dim", type=int, default=32)
    parser.add_argument("--vae_kl_coeff", type=float, action="append", default=[1.0])
    parser.add_argument("--vae_cycle_coeff", type=float, action="append", default=[1e-3])
    parser.add_argument("--latent_prior_type", type=str, action="append", default=["normal_zero_mean_one_std"])
    # engine knob (not in the reference): "bf16x3" reproduces fp32 results, "bf16" is the fast path
    parser.add_argument("--precision", type=str, default="bf16x3", choices=["bf16", "bf16x3"])
    # multi-GPU launch (torchrun): "dp" = every trial is data-parallel over all ranks; "replicas" = the grid points of the
    # sweep are spread over the ranks, one independent trainer per GPU (what Ray Tune's parallel trials do upstream)
    parser.add_argument("--sweep_mode", type=str, default="dp", choices=["dp", "replicas"])
    # engine knob (not in the reference): run-to-run bit-identical gradients (no split-K, ordered bias-gradient sums; slower)
    parser.add_argument("--deterministic", action="store_true")
    return parser


def merge_dataset(files):
    """train_physics_vae.py:94-114."""
    data_all = None
    for i, file in enumerate(files):
        with open(file, "rb") as f:
            data = pickle.load(f)
            print(file, "is loaded")
            if i == 0:
                data_all = data
            else:
                for key in ("iter_per_episode", "dim_state", "dim_state_body", "dim_state_task", "dim_action", "exp_std"):
                    assert data_all[key] == data[key]
                data_all["episodes"] = data_all["episodes"] + data["episodes"]
    return data_all


def episodes_to_transitions(episodes, num_samples=None, lookahead=1, cond="abs", use_a_gt=False):
    """Vectorised equivalent of the reference's per-transition loop (train_physics_vae.py:133-156): an episode of T steps
    yields T - lookahead items x_i = [[sb_{i+j}, sb_{i+j+1}] for j < lookahead], y_i = [action_{i+j}]; items never cross
    episode boundaries; `num_samples` caps the total.  Returns X float64 [N, lookahead, 2*dsb], Y [N, lookahead, da]."""
    assert lookahead >= 1
    Xs, Ys = [], []
    count = 0
    for ep in episodes:
        num_tuples = len(ep["time"])
        assert num_tuples >= lookahead
        n = num_tuples - lookahead
        if num_samples is not None:
            n = min(n, num_samples - count)
        if n <= 0:
            continue
        sb = np.asarray(ep["state_body"])
        act = np.asarray(ep["action_gt"] if use_a_gt else ep["action"])
        idx = np.arange(n)[:, None] + np.arange(lookahead)[None, :]            # [n, lookahead]
        s1, s2 = sb[idx], sb[idx + 1]
        if cond == "abs":
            x = np.concatenate([s1, s2], axis=-1)
        elif cond == "rel":
            x = np.concatenate([s1, s2 - s1], axis=-1)
        else:
            raise NotImplementedError
        Xs.append(x)
        Ys.append(act[idx])
        count += n
    if not Xs:
        return np.array([]), np.array([])
    return np.concatenate(Xs, axis=0), np.concatenate(Ys, axis=0)


def episodes_to_index(episodes, num_samples=None, use_a_gt=False):
    """The compact form of the same dataset (lookahead 1, cond "abs") for the device-side builder (`pvae_ingest_episodes`):
    all state_body rows back to back (every state once), the action rows, and per transition the row of s_t -- in exactly the
    order `episodes_to_transitions` emits X / Y rows.  Returns (states float64 [S, dsb], actions float32 [S, da], first int64 [N])."""
    S, A, first = [], [], []
    count, base = 0, 0
    for ep in episodes:
        sb = np.asarray(ep["state_body"], dtype=np.float64)
        act = np.asarray(ep["action_gt"] if use_a_gt else ep["action"], dtype=np.float32)
        T = len(ep["time"])
        n = T - 1
        if num_samples is not None:
            n = min(n, num_samples - count)
        if n > 0:
            first.append(base + np.arange(n, dtype=np.int64))
            count += n
        S.append(sb[:T]); A.append(act[:T])
        base += T
    return (np.concatenate(S, axis=0), np.concatenate(A, axis=0),
            np.concatenate(first) if first else np.zeros((0,), dtype=np.int64))


def load_dataset_for_PhysicsVAE(files, num_samples=None, lookahead=1, cond="abs", use_a_gt=False):
    """train_physics_vae.py:117-164."""
    assert files
    assert len(files) > 0
    data = merge_dataset(files)
    episodes = data["episodes"]
    X, Y = episodes_to_transitions(episodes, num_samples, lookahead, cond, use_a_gt)
    source = episodes_to_index(episodes, num_samples, use_a_gt) if (lookahead == 1 and cond == "abs" and len(X)) else None
    print("------------------Data Loaded------------------")
    print("File:", files)
    print("Num Episodes:", len(episodes))
    print("Num Transitions (Tuples):", len(X))
    print("-----------------------------------------------")
    dataset = torch_models.DatasetBase(X, Y, normalize_x=False, normalize_y=False)
    dataset.episode_source = source          # lets the trainer build the resident buffer on the device from the unique states
    return dataset


def create_model(config):
    """train_physics_vae.py:166-176."""
    model_config = config["model"]
    obs_space = model_config["custom_model_config"]["observation_space"]
    action_space = model_config["custom_model_config"]["action_space"]
    return policy_models.PhysicsVAE(obs_space=obs_space, action_space=action_space, num_outputs=2 * action_space.shape[0],
                                    model_config=model_config, name="physics_vae")


MODEL_CONFIG = copy.deepcopy(policy_models.PhysicsVAE.DEFAULT_CONFIG)


def gen_layers(width, depth, out_size="output", act_hidden="relu", act_out="linear", add_softmax=False):
    """train_physics_vae.py:180-192."""
    assert depth > 0 and width > 0
    layers = []
    for i in range(depth):
        layers.append({"type": "fc", "hidden_size": width, "activation": act_hidden, "init_weight": {"name": "normc", "std": 1.0}})
    layers.append({"type": "fc", "hidden_size": out_size, "activation": act_out, "init_weight": {"name": "normc", "std": 0.01}})
    if add_softmax:
        layers.append({"type": "softmax"})
    return layers


def inspect_dataset(file):
    with open(file, "rb") as f:
        data = pickle.load(f)
        ep0 = data["episodes"][0]
        dim_state_body = len(ep0["state_body"][0])
        dim_action = len(ep0["action"][0])
        return 2 * dim_state_body, dim_state_body, dim_state_body, dim_action


def get_trainer_config(args):
    """train_physics_vae.py:194-288: the task state at t is the body state at t+1, so dim_state = 2 * dim_state_body."""
    assert args.max_iter_world_model <= args.max_iter
    dim_state, dim_state_body, dim_state_task, dim_action = inspect_dataset(args.data_train[0])
    ob_scale, ac_scale = 1000.0, 3.0
    obs_space = Box(low=-ob_scale * np.ones(dim_state), high=ob_scale * np.ones(dim_state), dtype=np.float64)
    obs_space_body = Box(low=-ob_scale * np.ones(dim_state_body), high=ob_scale * np.ones(dim_state_body), dtype=np.float64)
    obs_space_task = Box(low=-ob_scale * np.ones(dim_state_task), high=ob_scale * np.ones(dim_state_task), dtype=np.float64)
    action_space = Box(low=-ac_scale * np.ones(dim_action), high=ac_scale * np.ones(dim_action), dtype=np.float64)

    model_config = {"custom_model": "physics_vae", "custom_model_config": MODEL_CONFIG.copy()}
    custom_model_config = model_config["custom_model_config"]
    custom_model_config["observation_space"] = obs_space
    custom_model_config["observation_space_body"] = obs_space_body
    custom_model_config["observation_space_task"] = obs_space_task
    custom_model_config["action_space"] = action_space
    custom_model_config["world_model_load_weights"] = args.world_model
    custom_model_config["engine_precision"] = getattr(args, "precision", "bf16x3")

    trainer_config = {
        "max_iter_world_model": args.max_iter_world_model,
        "model": model_config,
        "lr": args.lr,
        "lr_schedule_params": {"step_size": 50, "gamma": 0.70},
        "lr_schedule": args.lr_schedule,
        "weight_decay": 0.0,
        "dataset_train": args.data_train,
        "dataset_test": args.data_test,
        "use_gpu": False,
        "loss": "MSE",
        "loss_test": "MSE",
        "batch_size": args.batch_size,
        "suffle_data": True,       # sic: the loader reads "shuffle_data", so nothing is shuffled
        "latent_dim": args.latent_dim,
        "latent_prior_type": grid_search(args.latent_prior_type),
        "act_fn": "relu",
        "MD_width": grid_search([512]),
        "MD_depth": grid_search([3]),
        "TE_width": grid_search([256]),
        "TE_depth": grid_search([2]),
        "lookahead": 1,
        "world_model_width": grid_search([1024]),
        "world_model_depth": grid_search([2]),
        "vae_kl_coeff": grid_search(args.vae_kl_coeff),
        "motor_decoder_a_rec_coeff": 1.0,
        "world_model_s_rec_coeff": 0.0,
        "vae_cycle_coeff": grid_search(args.vae_cycle_coeff),
        "engine_precision": getattr(args, "precision", "bf16x3"),
        "deterministic": bool(getattr(args, "deterministic", False)),
        "num_data": getattr(args, "num_data", None),
    }
    return trainer_config


def resolve_grid(config):
    """All grid points of a config whose leaves may be {"grid_search": [...]} (what tune.run expands)."""
    keys = [k for k, v in config.items() if isinstance(v, dict) and set(v.keys()) == {"grid_search"}]
    points = []
    for combo in itertools.product(*[config[k]["grid_search"] for k in keys]):
        c = copy.copy(config)
        c["model"] = copy.deepcopy(config["model"])
        for k, v in zip(keys, combo):
            c[k] = v
        points.append(c)
    return points


def update_model_config(trainer_config):
    """train_physics_vae.py:290-311."""
    model_config = trainer_config["model"]["custom_model_config"]
    model_config["task_encoder_output_dim"] = trainer_config.get("latent_dim")
    model_config["task_encoder_layers"] = gen_layers(width=trainer_config.get("TE_width"), depth=trainer_config.get("TE_depth"),
                                                     act_hidden=trainer_config.get("act_fn"))
    model_config["motor_decoder_layers"] = gen_layers(width=trainer_config.get("MD_width"), depth=trainer_config.get("MD_depth"),
                                                      act_hidden=trainer_config.get("act_fn"))
    model_config["latent_prior_type"] = trainer_config.get("latent_prior_type")
    model_config["world_model_layers"] = gen_layers(width=trainer_config.get("world_model_width"),
                                                    depth=trainer_config.get("world_model_depth"),
                                                    act_hidden=trainer_config.get("act_fn"))
    if trainer_config.get("engine_precision"):
        model_config["engine_precision"] = trainer_config.get("engine_precision")


class TrainModel(torch_models.TrainModel):
    """train_physics_vae.py:313-467: world-model phase first ((a, kl, s, cyc) = (0, 0, 1, 0), world model trainable), then
    from iteration `max_iter_world_model` on the VAE phase ((1, kl, 0, cyc), encoder + decoder trainable, world model
    frozen but differentiated through)."""

    def setup(self, config):
        update_model_config(config)
        self.config = config
        self.max_iter_world_model = config.get("max_iter_world_model")
        self.latent_prior_type = config.get("latent_prior_type")
        self.lookahead = int(config.get("lookahead"))            # (the CLI hard-wires 1, train_physics_vae.py:277; > 1 = autoregressive rollout)
        assert self.lookahead >= 1
        self.noise_seed = int(config.get("noise_seed", 0))
        self._noise_step = 0
        self.world_phase = True
        super().setup(config)
        self.model.set_learnable_task_encoder(False)
        self.model.set_learnable_motor_decoder(False)
        self.model.set_learnable_world_model(True)
        self.read_loss_fn_coeff(world=True)
        self.sync_replicas()

    def _engine_changed(self, eng):
        if self.config.get("deterministic"):
            eng.set_deterministic(True)             # run-to-run bit-identical gradients (no split-K, ordered bias-gradient sums)
        self._arm_exchange(eng)

    def _arm_exchange(self, eng=None):
        """Data parallel over a symmetric gradient pool: let the engine exchange the part of the gradients that is complete early
        in the backward pass on a side stream, beside the remaining GEMMs; `_reduce` then exchanges only the rest."""
        had = getattr(self, "_early", None) is not None
        self._early = None
        eng = eng or self.engine
        if parallel.world_size() == 1 or not hasattr(self, "world_phase"):
            if had:
                pool = parallel.pool_of(self.model.reduce_range(True))
                if pool is not None:
                    pool.arm_overlap(eng, None)
            return
        early, late = self.model.reduce_ranges(self.world_phase)
        pool = parallel.pool_of(late)
        if pool is not None and pool.arm_overlap(eng, early):
            self._early = early

    def sync_replicas(self):
        """Data-parallel ranks must hold the same parameters: every rank builds its model from its own RNG stream, so the
        flat parameter buffers are broadcast from rank 0 (after construction and after every restore)."""
        if parallel.world_size() > 1:
            self.engine                                    # (makes sure the flat buffers exist)
            parallel.broadcast_([self.model.flat_params(n) for n in policy_models.NET_NAMES])
            self.model.mark_weights_dirty()

    def read_loss_fn_coeff(self, world):
        self.vae_kl_coeff = 0.0 if world else self.config.get("vae_kl_coeff")
        self.a_rec_coeff = 0.0 if world else self.config.get("motor_decoder_a_rec_coeff")
        self.s_rec_coeff = 1.0 if world else self.config.get("world_model_s_rec_coeff")
        self.vae_cycle_coeff = 0.0 if world else self.config.get("vae_cycle_coeff")
        self.world_phase = bool(world)
        if getattr(self, "model", None) is not None and getattr(self.model, "_engine", None) is not None:
            self._arm_exchange()                     # the early / late split of the exchanged range follows the phase

    def load_dataset(self, file):
        num_data = args.num_data if args is not None else self.config.get("num_data")
        return load_dataset_for_PhysicsVAE(file, num_samples=num_data, lookahead=self.lookahead)

    def step(self):
        if self.iter == self.max_iter_world_model:
            # end-to-end learning of the VAE starts (train_physics_vae.py:342-350)
            self._enter_vae_phase()
        return super().step()

    def _enter_vae_phase(self):
        self.model.set_learnable_task_encoder(True)
        self.model.set_learnable_motor_decoder(True)
        self.model.set_learnable_world_model(False)
        self.read_loss_fn_coeff(world=False)

    def load_checkpoint(self, checkpoint_path):
        super().load_checkpoint(checkpoint_path)
        self.sync_replicas()

    def load_trainer_state(self, st):
        """Resume: the phase is a function of the iteration counter (the switch fires when iter == max_iter_world_model)."""
        super().load_trainer_state(st)
        self._noise_step = int(st.get("noise_step", 0))
        if self.iter > self.max_iter_world_model:
            self._enter_vae_phase()
            for p_ in self.model._world_model.parameters():      # a frozen net's stale gradient must not reach Adam
                p_.grad = None

    def trainer_state(self):
        st = super().trainer_state()
        st["noise_step"] = self._noise_step
        return st

    def create_model(self, config):
        return create_model(config)

    def compute_model(self, x):
        logits, _ = self.model(input_dict={"obs": x, "obs_flat": x}, state=None, seq_lens=None)
        return logits[..., :logits.shape[1] // 2]

    # ---- one mini-batch on the engine ---------------------------------------------------------------------------------
    def _shard(self, lo, hi):
        """(first row, rows, loss weight) of this rank's part of the global mini-batch [lo, hi)."""
        world, r = parallel.world_size(), parallel.rank()
        if getattr(self, "dp_local_shards", False):
            # every rank holds its OWN shard of the global batch as rows [lo, hi) of its resident buffer (equal sizes)
            return lo, hi - lo, 1.0
        # every rank holds the whole dataset and takes its contiguous slice of the global batch
        s, e = parallel.shard_rows(lo, hi, r, world)
        return s, e - s, parallel.shard_weight(lo, hi, r, world)

    def _nets(self):
        return ["world_model"] if self.world_phase else ["task_encoder", "motor_decoder"]

    def _engine_step(self, n, w, eps=None, train=True):
        """Forward + loss (+ backward when `train`) for the `n` rows at the device cursor; loss coefficients weighted by `w`
        (n_rank * R / n, see parallel.py).  The Philox offset is (device-side noise counter) + rank."""
        eng = self.engine
        self._forked = False
        if self._early is not None and parallel.world_size() == 1:
            self._arm_exchange(eng)                  # (the ranks stopped sharing this trainer -- replica mode: nobody to exchange with)
        bufs = self._buffers[self._bound[1]][0]
        if isinstance(bufs, list) and len(bufs) > 1:
            # lookahead > 1: autoregressive rollout, gradients through time (pvae_rollout_step); the test pass runs the same call and
            # simply leaves the gradients unused
            kl = self.vae_kl_coeff if self.latent_prior_type else 0.0
            eng.rollout_step(n, self.world_phase, bufs, self._buffers[self._bound[1]][1], eps=eps, seed=self.noise_seed,
                             offset=parallel.rank() + 1000003 * self._noise_step, noise=bool(self.model.latent_prior_noise),
                             a_coeff=self.a_rec_coeff * w, kl_coeff=kl * w, s_coeff=self.s_rec_coeff * w, cyc_coeff=self.vae_cycle_coeff * w)
            if w != 1.0:
                eng.loss[1:5].mul_(w)
            return
        if self.world_phase:
            if self.s_rec_coeff <= 0:
                raise ValueError("world phase needs s_rec_coeff > 0")
            if train:
                eng.world_step(n, s_coeff=self.s_rec_coeff * w)
                self._forked = self._early is not None       # (the step exchanged the early range itself)
            else:
                eng.eval_loss(n, True, s_coeff=self.s_rec_coeff * w)
        else:
            if self.s_rec_coeff and self.s_rec_coeff > 0:
                raise NotImplementedError("world_model_s_rec_coeff > 0 in the VAE phase is not used by the CLI (it is 0.0)")
            kl = self.vae_kl_coeff if self.latent_prior_type else 0.0
            kw = dict(eps=eps, seed=self.noise_seed, offset=parallel.rank(), noise=bool(self.model.latent_prior_noise),
                      a_coeff=self.a_rec_coeff * w, kl_coeff=kl * w, cyc_coeff=self.vae_cycle_coeff * w)
            if train:
                eng.vae_step(n, **kw)
                self._forked = self._early is not None
            else:
                eng.eval_loss(n, False, **kw)
        if w != 1.0:
            eng.loss[1:5].mul_(w)          # the component slots are local means: weight them like slot 0 so that the rank average is the global mean

    def _reduce(self, train=True):
        """The step's only collective: ONE averaging all-reduce over [gradients of the trained nets | loss slots] -- a single
        contiguous range of the model's gradient pool (PhysicsVAE.reduce_range)."""
        if parallel.world_size() > 1:
            if not train:
                rng = self.engine.loss
            elif getattr(self, "_forked", False):
                rng = self.model.reduce_ranges(self.world_phase)[1]      # the engine step exchanged the early part beside its GEMMs
            else:
                rng = self.model.reduce_range(self.world_phase)
            parallel.allreduce_avg_([rng])

    def batch_loss(self, lo, hi, eps=None, train=True):
        """Forward + loss + backward for rows [lo, hi) of the resident buffer; gradients land in `.grad` (all-reduced when
        torch.distributed is initialised).  Returns the global mini-batch loss as a 0-dim device tensor."""
        eng = self.engine
        if self.model._weights_dirty:
            self.model.sync_weights()
        s, n, w = self._shard(lo, hi)
        if not self.world_phase:
            self._noise_step += 1
            eng.noise_counter(True, self._noise_step * parallel.world_size(), parallel.world_size())
        if n > 0:
            eng.set_cursor(s)
            self._engine_step(n, w, eps=eps, train=train)
        else:
            self._forked = False
            if train:
                for name in self._nets():
                    self.model.flat_grads(name).zero_()
                if self._early is not None:          # the other ranks' steps exchange the early range themselves: take part
                    eng.run_exchange()
                    self._forked = True
            eng.loss.zero_()
        self._reduce(train)
        return eng.loss[0].clone()

    def eval_batch_loss(self, lo, hi):
        """The reference's test pass (torch_models.py:147-155): compute_test_loss under no_grad -- forward and loss only; no
        gradient buffer is touched, nothing but the loss slots is exchanged between ranks."""
        return self.batch_loss(lo, hi, train=False)

    def train_on_episodes(self, states, actions, first_state):
        """One whole SGD step (zero_grad -> compute_loss -> backward -> [exchange] -> Adam, torch_models.py:137-143) on a mini-batch
        handed over in the compact form of compute_loss_episodes, as DEVICE tensors (the caller's loader copied them from pinned host
        memory).  The step -- device-side dataset build included -- is ONE captured CUDA graph per staging buffer set, so the host
        side of a step is a graph launch; returns the loss as a 0-dim device tensor (read it with .item(), like the reference)."""
        B = int(first_state.shape[0])
        if B < 2:
            raise ValueError("a mini-batch needs at least 2 transitions (the reference squeezes the batch axis)")
        if not self._graph_ok(B) or self.lookahead != 1:
            loss = self.compute_loss_episodes(states, actions, first_state)
            self.optimizer.step()
            return loss.detach()
        self._engine_rows = max(self._engine_rows, B)
        eng = self.engine
        buf = self._buffers.get("adhoc")
        if buf is None or isinstance(buf[0], list) or buf[1] != B:
            buf = (torch.zeros(eng.transitions_bytes(B), dtype=torch.uint8, device=self.device), B)
            self._buffers["adhoc"] = buf
        self._bind("adhoc")
        key = ("episodes", states.data_ptr(), actions.data_ptr(), first_state.data_ptr(), str(states.dtype), id(buf[0])) + self._graph_key(B)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) > 8:
                self._graphs.clear()
            self._graph_prepare(B)
            if self.model._weights_dirty:
                self.model.sync_weights()
            s0, n, w = self._shard(0, B)
            if n <= 0:
                raise _abi.PvaeError("a captured step needs at least one row per rank")

            def body():
                eng.ingest_episodes(states, actions, first_state, dst_row=0, check=False)
                eng.set_cursor(s0)
                self._engine_step(n, w)
                self._reduce()
                self.optimizer.step()
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            self._graphs[key] = g
            if not self.world_phase:
                eng.noise_counter(True, (self._noise_step + 1) * parallel.world_size(), parallel.world_size())
        g.replay()
        if not self.world_phase:
            self._noise_step += 1
        return eng.loss[0]

    # ---- the captured full-batch step (torch_models.TrainModel._graph_step) ------------------------------------------------
    def _graph_supported(self):
        return self.lookahead == 1          # (a rollout re-carves the workspace per step on the host: eager launches)

    def _graph_key(self, batch_size):
        return super()._graph_key(batch_size) + (self.world_phase, self.s_rec_coeff, self.a_rec_coeff, self.vae_kl_coeff,
                                                 self.vae_cycle_coeff, bool(self.model.latent_prior_noise), self.noise_seed,
                                                 bool(getattr(self, "dp_local_shards", False)))

    def _graph_prepare(self, batch_size):
        self.optimizer.prepare(self._nets())             # Adam moments exist before the capture (no persistent allocation inside it)
        if not self.world_phase:
            # the device-side noise counter must be ON while the step is captured: the library decides at launch time whether the
            # reparameterisation kernel reads it (and the finalisation kernel bumps it)
            self.engine.noise_counter(True, (self._noise_step + 1) * parallel.world_size(), parallel.world_size())

    def _graph_body(self, batch_size):
        _, n, w = self._shard(0, batch_size)
        if n <= 0:
            raise _abi.PvaeError("a captured step needs at least one row per rank (batch_size >= world size)")
        probe = getattr(self, "_graph_probe", None)     # (bench.py) two external CUDA events around the engine's launch sequence
        if probe:
            probe[0].record()
        self._engine_step(n, w)
        if probe:
            probe[1].record()
        self._reduce()

    def _graph_begin(self, batch_size):
        super()._graph_begin(batch_size)
        if not self.world_phase:
            self.engine.noise_counter(True, (self._noise_step + 1) * parallel.world_size(), parallel.world_size())

    def _graph_end(self, n_full, batch_size):
        if not self.world_phase:
            self._noise_step += n_full

    def compute_loss(self, y, x, eps=None):
        """Reference signature (train_physics_vae.py:361-435): x [B, 1, 2*dsb], y [B, 1, da] -> scalar loss.  The batch is
        staged into the resident buffer and run through the engine; gradients are deposited into `.grad`, and the returned
        tensor's backward() is a no-op."""
        B = x.shape[0]
        if B < 2:
            raise ValueError("compute_loss needs at least 2 transitions (the reference squeezes the batch axis)")
        L = x.shape[1] if x.dim() == 3 else 1
        self._engine_rows = max(self._engine_rows, B)
        eng = self.engine                                # (a larger B re-creates the engine: _bind notices and drops the graphs)
        buf = self._buffers.get("adhoc")
        if buf is None or buf[1] != B or (len(buf[0]) if isinstance(buf[0], list) else 1) != L:
            mk = lambda: torch.zeros(eng.transitions_bytes(B), dtype=torch.uint8, device=self.device)
            buf = ([mk() for _ in range(L)], B) if L > 1 else (mk(), B)
            self._buffers["adhoc"] = buf
        self._bind("adhoc")
        if L > 1:                                        # one resident buffer per rollout step: (x[:, t], y[:, t])
            for t in range(L):
                eng.bind_transitions(buf[0][t], B)
                eng.ingest(x[:, t].reshape(B, -1).to(self.device), y[:, t].reshape(B, -1).to(self.device))
            eng.bind_transitions(buf[0][0], B)
        else:
            eng.ingest(x.reshape(B, -1).to(self.device), y.reshape(B, -1).to(self.device))
        loss = self.batch_loss(0, B, eps=eps)
        return torch_models._DepositedLoss.apply(self._anchor, loss)

    def compute_test_loss(self, y, x):
        return self.compute_loss(y, x)

    def compute_loss_episodes(self, states, actions, first_state, eps=None):
        """compute_loss for a mini-batch handed over in the COMPACT form of the dataset (episodes_to_index): states [S, dsb] with every
        state once, actions [S, da], first_state [B] = row of s_t per transition (s_{t+1} is the next row).  Same arithmetic as
        compute_loss(y, x) on the expanded x = [s_t | s_{t+1}], y = a_t; the upload is ~half the bytes because x is never
        materialised -- the device-side builder (pvae_ingest_episodes) pairs the rows."""
        B = int(first_state.shape[0])
        if B < 2:
            raise ValueError("compute_loss needs at least 2 transitions (the reference squeezes the batch axis)")
        self._engine_rows = max(self._engine_rows, B)
        eng = self.engine
        buf = self._buffers.get("adhoc")
        if buf is None or buf[1] != B:
            buf = (torch.zeros(eng.transitions_bytes(B), dtype=torch.uint8, device=self.device), B)
            self._buffers["adhoc"] = buf
        self._bind("adhoc")
        eng.ingest_episodes(states, actions, first_state, dst_row=0, check=False)
        loss = self.batch_loss(0, B, eps=eps)
        return torch_models._DepositedLoss.apply(self._anchor, loss)

    def save_checkpoint(self, checkpoint_dir):
        """train_physics_vae.py:440-467: model.pth + model.pt + task_encoder.pt + motor_decoder.pt + world_model.pt."""
        checkpoint_path = super().save_checkpoint(checkpoint_dir)
        checkpoint = os.path.join(checkpoint_dir, "model.pt")
        self.model.save_weights(checkpoint)
        print("Saved:", checkpoint)
        if self.model._task_encoder:
            checkpoint = os.path.join(checkpoint_dir, "task_encoder.pt")
            self.model.save_weights_task_encoder(checkpoint)
            print("Saved:", checkpoint)
        if self.model._motor_decoder:
            checkpoint = os.path.join(checkpoint_dir, "motor_decoder.pt")
            self.model.save_weights_motor_decoder(checkpoint)
            print("Saved:", checkpoint)
        if self.model._world_model:
            checkpoint = os.path.join(checkpoint_dir, "world_model.pt")
            self.model.save_weights_world_model(checkpoint)
            print("Saved:", checkpoint)
        return checkpoint_path


def latest_checkpoint(trial_dir):
    """model.pth of the newest checkpoint_NNNNNN directory of a trial, or None."""
    if not os.path.isdir(trial_dir):
        return None
    ck = sorted(d for d in os.listdir(trial_dir) if d.startswith("checkpoint_") and os.path.exists(os.path.join(trial_dir, d, "model.pth")))
    return os.path.join(trial_dir, ck[-1], "model.pth") if ck else None


def run_trial(config, max_iter, checkpoint_freq, trial_dir, restore=None):
    """What one Ray Tune trial does (train_physics_vae.py:484-502): train() until training_iteration == max_iter,
    checkpoint every `checkpoint_freq` iterations and at the end; results go to result.json like Tune's JSON logger.
    `restore`: continue from that checkpoint (weights; with its trainer_state.pt also Adam moments, LR schedule, phase and
    iteration counter -- upstream a resumed trial restarts those).
    Data-parallel trials (torchrun, --sweep_mode dp): every rank trains, rank 0 alone writes result.json and the checkpoints;
    all ranks leave a checkpoint iteration together (barrier) and return the same checkpoint path."""
    writer = parallel.rank() == 0
    if writer:
        os.makedirs(trial_dir, exist_ok=True)
    config = dict(config)
    config["save_trainer_state"] = True
    trainer = TrainModel(config)
    if restore:
        trainer.restore(restore)
    last = restore
    log = open(os.path.join(trial_dir, "result.json"), "a") if writer else None
    try:
        for it in range(trainer.training_iteration + 1, max_iter + 1):
            result = trainer.train()
            if log:
                log.write(json.dumps(result) + "\n")
                log.flush()
            if (checkpoint_freq and it % checkpoint_freq == 0) or it == max_iter:
                ckpt_dir = os.path.join(trial_dir, "checkpoint_%06d" % it)
                if writer:
                    trainer.save(ckpt_dir)
                last = os.path.join(ckpt_dir, "model.pth")
                parallel.barrier()                 # nobody runs ahead of (or resumes from) a checkpoint that is still being written
    finally:
        if log:
            log.close()
    return trainer, last


def output_path(path, trial, n_trials):
    """--output of a sweep with several grid points: one file per trial (`model.pt` -> `model.trial_00003.pt`)."""
    if n_trials <= 1:
        return path
    root, ext = os.path.splitext(path)
    return "%s.trial_%05d%s" % (root, trial, ext)


def init_distributed(sweep_mode="dp"):
    """Under torchrun (WORLD_SIZE > 1): one process per GPU, NCCL process group; "replicas" keeps the ranks independent."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        backend = os.environ.get("PVAE_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            if os.environ.get("PVAE_NUMA_BIND", "1") != "0":
                parallel.bind_to_gpu_numa_node(local)     # loader process + its pinned staging memory next to the GPU's PCIe root
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)          # (gloo moves CUDA tensors through the host: tests with several ranks on one GPU)
    parallel.set_replica_mode(sweep_mode == "replicas")


def main(argv=None):
    global args
    args = arg_parser().parse_args(argv)
    trainer_config = get_trainer_config(args)
    checkpoint = args.checkpoint
    init_distributed(args.sweep_mode)
    done = []                                   # (trial index, last checkpoint) of the trials this process ran
    n_points = 1
    if args.checkpoint is None:
        local_dir = os.path.expanduser(args.local_dir)
        name = args.name or "TrainModel"
        points = resolve_grid(trainer_config)
        n_points = len(points)
        mine = parallel.sweep_points(len(points), parallel.job_rank(), parallel.job_world_size()) if args.sweep_mode == "replicas" \
            else list(range(len(points)))
        for i in mine:
            trial_dir = os.path.join(local_dir, name, "trial_%05d" % i)
            restore = latest_checkpoint(trial_dir) if args.resume else None
            _, checkpoint = run_trial(points[i], args.max_iter, args.checkpoint_freq, trial_dir, restore=restore)
            done.append((i, checkpoint))
    else:
        done.append((0, checkpoint))
    if args.output is not None and parallel.rank() == 0:
        # the reference's --output branch instantiates the abstract base trainer and always fails; here it
        # exports the full state dict of the checkpoint -- per trial when the sweep has several grid points; in a
        # data-parallel job rank 0 alone writes
        for i, ck in done:
            if ck is None:
                continue
            out = output_path(args.output, i, n_points)
            torch.save(torch.load(ck, map_location="cpu"), out)
            print("Model Saved:", out)
    return checkpoint


if __name__ == "__main__":
    main()
