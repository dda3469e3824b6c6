"""Host-side mirror of the reference's generic trainer (torch_models.py:21-216): `get_lr_scheduler`, `DatasetBase`,
`get_loss_fn`, `TrainModel` with the Ray Tune `Trainable` protocol (`setup(config)`, `step() -> dict`,
`save_checkpoint(dir) -> path`, `load_checkpoint(path)`), plus the `WorldModel` / `Motor` views the task framing names.

What changed underneath: the mini-batch loop does not collate tensors on the host.  The dataset is converted once into a
GPU-resident transition buffer (engine.ingest) and every mini-batch is a row range of it; forward, loss and backward are
one library call; the loss is accumulated on the device and read back once per epoch instead of `.item()` per batch.
"""
import os

import numpy as np
import torch
import torch.optim as optim
import torch.utils.data as data

from . import _abi, parallel
from .optim import PvaeAdam

try:  # pragma: no cover - ray is not installed in the build image
    from ray import tune
    _TrainableBase = tune.Trainable
except Exception:  # noqa
    import copy as _copy

    class _TrainableBase(object):
        """Stand-in for ray.tune.Trainable: the constructor calls setup(config), train() calls step()."""

        def __init__(self, config=None, logger_creator=None):
            self.config = config or {}
            self._iteration = 0
            self.setup(_copy.deepcopy(self.config))

        def train(self):
            result = dict(self.step())
            self._iteration += 1
            result["training_iteration"] = self._iteration
            return result

        @property
        def training_iteration(self):
            return self._iteration

        def save(self, checkpoint_dir):
            os.makedirs(checkpoint_dir, exist_ok=True)
            return self.save_checkpoint(checkpoint_dir)

        def restore(self, checkpoint_path):
            self.load_checkpoint(checkpoint_path)
            it = getattr(self, "_restored_iteration", None)
            if it is not None:
                self._iteration = it

        def stop(self):
            pass

EPSILON = np.finfo(np.float32).eps
TRAINER_STATE_FILE = "trainer_state.pt"


def get_lr_scheduler(optimizer, name, params):
    """torch_models.py:21-37."""
    lr_scheduler = None
    if name == "cosine":
        lr_scheduler = optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=params["T_max"])
    elif name == "cosine_restart":
        lr_scheduler = optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, T_0=params["T_0"], T_mult=params["T_mult"])
    elif name == "step":
        lr_scheduler = optim.lr_scheduler.StepLR(optimizer, step_size=params["step_size"], gamma=params["gamma"])
    return lr_scheduler


class DatasetBase(data.Dataset):
    """torch_models.py:39-95: holds X [N, lookahead, 2*dsb] float64 and Y [N, lookahead, da]; optional normalisation."""

    def __init__(self, X, Y, normalize_x=True, normalize_y=True):
        self.X = X
        self.Y = Y
        self.normalize_x = normalize_x
        self.normalize_y = normalize_y
        if normalize_x:
            self.X_mean, self.X_std = np.mean(self.X, axis=0), np.std(self.X, axis=0)
        if normalize_y:
            self.Y_mean, self.Y_std = np.mean(self.Y, axis=0), np.std(self.Y, axis=0)

    def __getitem__(self, index):
        return self.preprocess_x(self.X[index]), self.preprocess_y(self.Y[index])

    def __len__(self):
        return len(self.X)

    def preprocess_x(self, x, return_tensor=True):
        x_new = (x - self.X_mean) / (self.X_std + EPSILON) if self.normalize_x else x
        return torch.Tensor(x_new) if return_tensor else x_new

    def postprocess_x(self, x, return_tensor=True):
        x_new = self.X_mean + np.multiply(x, self.X_std) if self.normalize_x else x
        return torch.Tensor(x_new) if return_tensor else x_new

    def preprocess_y(self, y, return_tensor=True):
        y_new = (y - self.Y_mean) / (self.Y_std + EPSILON) if self.normalize_y else y
        return torch.Tensor(y_new) if return_tensor else y_new

    def postprocess_y(self, y, return_tensor=True):
        y_new = self.Y_mean + np.multiply(y, self.Y_std) if self.normalize_y else y
        return torch.Tensor(y_new) if return_tensor else y_new

    def arrays(self):
        """(X, Y) as the engine ingests them: the preprocessed arrays, X float64 / Y float32, [N, L, dim]."""
        X = self.preprocess_x(np.asarray(self.X), return_tensor=False)
        Y = self.preprocess_y(np.asarray(self.Y), return_tensor=False)
        return np.ascontiguousarray(X, dtype=np.float64), np.ascontiguousarray(Y, dtype=np.float32)


def get_loss_fn(loss):
    """torch_models.py:97-107.  The engine fuses the MSE reduction into the last GEMM's epilogue, so only "MSE" is
    executable on the hot path; the others keep the reference's error behaviour for unknown names."""
    if loss == "MSE":
        return torch.nn.MSELoss()
    elif loss == "MAE" or loss == "L1":
        return torch.nn.L1Loss()
    elif loss == "CrossEntropy":
        return torch.nn.CrossEntropyLoss()
    elif loss == "NLLLoss":
        return torch.nn.NLLLoss()
    else:
        raise NotImplementedError


class ResidentLoader(object):
    """What `DataLoader(dataset, batch_size, shuffle=None)` is on the reference path (SequentialSampler, drop_last=False;
    SURVEY.md F4): consecutive row ranges of the GPU-resident transition buffer."""

    def __init__(self, dataset, batch_size, shuffle=None):
        if shuffle:
            raise NotImplementedError("the reference never shuffles on this path (config key typo, SURVEY.md F4)")
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.n = len(dataset)

    def __len__(self):
        return (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for lo in range(0, self.n, self.batch_size):
            yield lo, min(lo + self.batch_size, self.n)


class _DepositedLoss(torch.autograd.Function):
    """The engine deposits gradients straight into `.grad`; `loss.backward()` of reference-style callers is a no-op."""

    @staticmethod
    def forward(ctx, anchor, loss):
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        return None, None


class TrainModel(_TrainableBase):
    """torch_models.py:109-216."""

    def setup(self, config):
        self.model = self.create_model(config)
        self.prepare_data(config)
        if not torch.cuda.is_available():
            raise _abi.PvaeError("physicsvae_b200.TrainModel needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.model = self.model.to(self.device)
        self._engine_rows = self._local_rows(config.get("batch_size"))
        self._engine_precision = config.get("engine_precision", "bf16x3")
        self._buffers = {}             # "train" / "test" / "adhoc" -> (resident transition buffer, rows)
        self._bound = (None, None)     # (engine instance, buffer name) the engine currently reads
        self._graphs = {}              # captured full-batch steps, see _graph_step
        self._use_graphs = bool(config.get("cuda_graph", True))
        self._loss_acc = torch.zeros((), device=self.device)
        # data-parallel ranks on one NVLink node: the gradient pool lives in symmetric (peer-mapped) memory and the exchange is
        # the library's own kernel (parallel.SymmetricPool); None = plain tensor + torch.distributed
        self.model.grad_pool_factory = parallel.symmetric_pool_factory()
        eng = self.engine
        self._engine_changed(eng)
        lr = config.get("lr", 1e-3)
        if config.get("optimizer", "pvae_adam") == "torch_adam":
            capturable = bool(config.get("optimizer_capturable", False))     # CUDA-graph replay of the whole step
            self.optimizer = optim.Adam(self.model.parameters(), lr=torch.tensor(float(lr), device=self.device) if capturable else lr,
                                        weight_decay=config.get("weight_decay", 0.0), fused=True, capturable=capturable)
            self._use_graphs = self._use_graphs and capturable
        else:
            # same interface and arithmetic as optim.Adam, executed by the engine together with the shadow-weight refresh
            self.optimizer = PvaeAdam(self.model, lr=lr, weight_decay=config.get("weight_decay", 0.0))
        self.lr_scheduler = get_lr_scheduler(self.optimizer, config.get("lr_schedule", None), config.get("lr_schedule_params", None))
        self.loss_fn = get_loss_fn(config.get("loss", "MSE"))
        self.loss_fn_test = get_loss_fn(config.get("loss_test", "MSE"))
        if config.get("loss", "MSE") != "MSE" or config.get("loss_test", "MSE") != "MSE":
            raise NotImplementedError("only the MSE loss is fused into the sm_100a engine")
        self.iter = 0
        self._upload(self.train_loader, "train")
        if self.test_loader:
            self._upload(self.test_loader, "test")       # both sets stay resident; a pass only re-binds the engine
        self._bind("train")
        self._anchor = torch.zeros((), device=self.device, requires_grad=True)

    def _local_rows(self, batch_size):
        return max(parallel.max_shard_rows(int(batch_size), parallel.world_size()), 2)

    @property
    def engine(self):
        """Always the model's CURRENT engine: the model re-creates it when a larger batch is asked for (compute_model /
        compute_loss on more rows than batch_size), which invalidates bindings and captured graphs -- see _bind."""
        return self.model.engine(max_batch=self._engine_rows, precision=self._engine_precision)

    # ---- data -------------------------------------------------------------------------------------------------------
    def load_dataset(self, file):
        raise NotImplementedError

    def get_data_loader(self, dataset, batch_size, shuffle):
        return ResidentLoader(dataset, batch_size, shuffle)

    def prepare_data(self, config):
        dataset_train = config.get("dataset_train")
        dataset_test = config.get("dataset_test")
        batch_size = config.get("batch_size")
        shuffle_data = config.get("shuffle_data")      # the CLI sets "suffle_data": this stays None
        dataset_train = self.load_dataset(dataset_train)
        if dataset_test is not None:
            dataset_test = self.load_dataset(dataset_test)
        self.train_loader = self.get_data_loader(dataset_train, batch_size, shuffle_data)
        self.test_loader = self.get_data_loader(dataset_test, batch_size, shuffle_data) if dataset_test is not None else None

    def _bind(self, which):
        """Point the engine at one of the resident buffers.  Re-binding is free (no copy); if the model has re-created its
        engine since the last call the binding is re-established and the captured graphs (which hold the old handle's
        pointers) are dropped."""
        eng = self.engine
        if self._bound[0] is not eng:
            self._graphs = {}
            self._engine_changed(eng)
        if self._bound != (eng, which):
            buf, n = self._buffers[which]
            eng.bind_transitions(buf[0] if isinstance(buf, list) else buf, n)      # (a list: one buffer per rollout step, lookahead > 1)
            self._bound = (eng, which)
        return eng

    def _engine_changed(self, eng):
        """Hook: a fresh engine instance is in use (subclasses re-arm per-engine device state)."""

    def _upload(self, loader, which):
        """DatasetBase -> resident bf16 transition buffer (replaces per-item torch.Tensor + collate, torch_models.py:52-68)."""
        eng = self.engine
        src = getattr(loader.dataset, "episode_source", None)
        if src is not None and not loader.dataset.normalize_x and not loader.dataset.normalize_y and len(src[2]) == len(loader.dataset):
            # dataset build on the device: upload every state once, the ingest kernel pairs (s_t, a_t, s_{t+1}) by index
            states, actions, first = src
            eng.alloc_transitions(len(first))
            st, ac = torch.from_numpy(states).to(self.device), torch.from_numpy(actions).to(self.device)
            chunk = 1 << 20
            for lo in range(0, len(first), chunk):
                eng.ingest_episodes(st, ac, torch.from_numpy(first[lo:lo + chunk]).to(self.device), dst_row=lo)
        elif np.asarray(loader.dataset.X).ndim == 3 and np.asarray(loader.dataset.X).shape[1] > 1:
            # lookahead L > 1 (autoregressive rollout): one resident buffer per rollout step t, holding (x[:, t], y[:, t])
            X, Y = loader.dataset.arrays()
            n, L = X.shape[0], X.shape[1]
            bufs = []
            for t in range(L):
                eng.alloc_transitions(n)
                chunk = 1 << 18
                for lo in range(0, n, chunk):
                    hi = min(lo + chunk, n)
                    eng.ingest(torch.from_numpy(np.ascontiguousarray(X[lo:hi, t])).to(self.device),
                               torch.from_numpy(np.ascontiguousarray(Y[lo:hi, t])).to(self.device), dst_row=lo)
                bufs.append(eng.transitions)
            self._buffers[which] = (bufs, n)
            self._bound = (eng, which)
            return
        else:
            X, Y = loader.dataset.arrays()
            n = X.shape[0]
            eng.alloc_transitions(n)
            chunk = 1 << 18
            for lo in range(0, n, chunk):
                hi = min(lo + chunk, n)
                eng.ingest(torch.from_numpy(X[lo:hi].reshape(hi - lo, -1)).to(self.device),
                           torch.from_numpy(Y[lo:hi].reshape(hi - lo, -1)).to(self.device), dst_row=lo)
        self._buffers[which] = (eng.transitions, eng.n_rows)
        self._bound = (eng, which)

    # ---- the SGD loop (torch_models.py:131-161) -----------------------------------------------------------------------
    def step(self):
        import time as _time
        t_start = _time.perf_counter()
        self.iter += 1
        self.model.train()
        self._bind("train")
        self._loss_acc.zero_()
        bs, n = self.train_loader.batch_size, self.train_loader.n
        n_full = n // bs if self._graph_ok(bs) else 0
        if n_full:
            self._graph_epoch(n_full, bs)               # full mini-batches: one captured graph replayed off the device cursor
        for lo in range(n_full * bs, n, bs):            # the short last batch (or everything, without graphs): eager launches
            self._loss_acc += self.train_batch(lo, min(lo + bs, n))
        mean_train_loss = float(self._loss_acc.item()) / len(self.train_loader)
        train_seconds = _time.perf_counter() - t_start              # (the .item() above waited for the epoch's last kernel)

        mean_test_loss = 0.0
        if self.test_loader:
            # forward + loss only, like the reference's `with torch.no_grad()` test pass (torch_models.py:147-155)
            self._bind("test")
            self._loss_acc.zero_()
            for lo, hi in self.test_loader:
                self._loss_acc += self.eval_batch_loss(lo, hi)
            mean_test_loss = float(self._loss_acc.item()) / len(self.test_loader)
            self._bind("train")

        if self.lr_scheduler:
            self.lr_scheduler.step()
        # the reference's two keys (torch_models.py:160-161) + throughput of the training pass
        return {"mean_train_loss": mean_train_loss, "mean_test_loss": mean_test_loss,
                "transitions_per_sec": n / max(train_seconds, 1e-9), "train_seconds": train_seconds}

    def train_batch(self, lo, hi):
        """zero_grad -> compute_loss -> backward -> [all-reduce] -> Adam, for rows [lo, hi) of the resident buffer."""
        loss = self.batch_loss(lo, hi)
        self.optimizer.step()
        if not isinstance(self.optimizer, PvaeAdam):
            self.model.mark_weights_dirty()      # (PvaeAdam refreshes the shadow operands itself)
        return loss

    # ---- captured full-batch step --------------------------------------------------------------------------------------------
    def _graph_ok(self, batch_size):
        return self._use_graphs and isinstance(self.optimizer, PvaeAdam) and self._graph_supported() and parallel.graph_capturable()

    def _graph_supported(self):
        return False

    def _graph_key(self, batch_size):
        """Everything a captured step bakes in as a host-side constant."""
        lr = self.optimizer.param_groups[0]["lr"]
        return (int(batch_size), float(lr), parallel.world_size(), parallel.rank(), id(getattr(self, "_graph_probe", None)))

    def _shard(self, lo, hi):
        """(first row, rows, loss weight) of this rank's part of the global mini-batch [lo, hi)."""
        s, e = parallel.shard_rows(lo, hi, parallel.rank(), parallel.world_size())
        return s, e - s, parallel.shard_weight(lo, hi, parallel.rank(), parallel.world_size())

    def _graph_body(self, batch_size):
        """One full mini-batch starting at the device cursor: engine step [+ all-reduce]; must not allocate or synchronise."""
        raise NotImplementedError

    def _graph_prepare(self, batch_size):
        """Anything a captured step needs allocated / initialised beforehand (called before the capture)."""

    def _graph_step(self, batch_size):
        """The step as ONE CUDA graph per (phase, batch size, lr, ...): [engine step -> all-reduce -> fused Adam + shadow refresh ->
        loss accumulation -> cursor advance].  Replays walk the resident buffer through the device-side cursor."""
        key = self._graph_key(batch_size)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) > 8:                   # lr schedules change the key every few epochs: keep the cache small
                self._graphs.clear()
            self._graph_prepare(batch_size)
            if self.model._weights_dirty:
                self.model.sync_weights()
            eng = self.engine

            _, n_local, _ = self._shard(0, batch_size)
            n_full = self.train_loader.n // batch_size

            def body():
                self._graph_body(batch_size)
                self.optimizer.step()
                self._loss_acc.add_(eng.loss[0])
                eng.advance_cursor(batch_size, n_local, n_full * batch_size)     # wraps to this rank's first row after the last full batch
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            # the capture pass does not execute anything: the Adam step counters, cursor and noise counter are untouched
            self._graphs[key] = g
        return g

    def _graph_epoch(self, n_full, batch_size):
        g = self._graph_step(batch_size)
        self._streaming = None
        self._graph_begin(batch_size)
        for _ in range(n_full):
            g.replay()
        self._graph_end(n_full, batch_size)

    def _graph_begin(self, batch_size):
        """Position the device-side state (cursor, noise counter) at the first full batch of the epoch."""
        self.engine.set_cursor(self._shard(0, batch_size)[0])

    def train_steps(self, k, batch_size=None, restart=False):
        """Streaming mode: `k` mini-batch SGD steps on consecutive FULL mini-batches of the resident training set, wrapping at
        its end, with no host synchronisation at all -- k replays of the captured step (what `step()` replays per epoch, minus
        the epoch bookkeeping: the loss accumulates on the device in `_loss_acc`, schedulers are not stepped).  `restart`
        repositions the cursor on the first mini-batch.  This is the call bench.py times as `value`."""
        bs = int(batch_size or self.train_loader.batch_size)
        if not self._graph_ok(bs) or self.train_loader.n < bs:
            raise _abi.PvaeError("train_steps needs the captured-graph path (PvaeAdam, cuda_graph=True) and at least one full mini-batch")
        self.model.train()
        self._bind("train")
        g = self._graph_step(bs)
        key = self._graph_key(bs)
        if restart or getattr(self, "_streaming", None) != key:      # first call, or phase / lr / batch size changed
            self._graph_begin(bs)
            self._streaming = key
        for _ in range(int(k)):
            g.replay()
        self._graph_end(int(k), bs)

    def _graph_end(self, n_full, batch_size):
        """Host-side bookkeeping for `n_full` replayed batches."""

    def batch_loss(self, lo, hi):
        raise NotImplementedError

    def eval_batch_loss(self, lo, hi):
        """Loss of rows [lo, hi) of the bound buffer, forward only.  Default: the training pass (subclasses do better)."""
        return self.batch_loss(lo, hi)

    def create_model(self, config):
        return config.get("model")

    def compute_model(self, x):
        return self.model(x)

    def compute_loss(self, y, x):
        raise NotImplementedError

    def compute_test_loss(self, y, x):
        return self.compute_loss(y, x)

    # ---- checkpoints (torch_models.py:209-216) ---------------------------------------------------------------------------
    def save_checkpoint(self, checkpoint_dir):
        print(checkpoint_dir)
        os.makedirs(checkpoint_dir, exist_ok=True)
        checkpoint_path = os.path.join(checkpoint_dir, "model.pth")
        torch.save({k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}, checkpoint_path)
        if self.config.get("save_trainer_state", False):
            # beyond the reference (which restarts Adam's moments and the StepLR counter on resume):
            # what a bit-faithful continuation needs, in a file the reference's loaders never look at
            torch.save(self.trainer_state(), os.path.join(checkpoint_dir, TRAINER_STATE_FILE))
        return checkpoint_path

    def load_checkpoint(self, checkpoint_path):
        self.model.load_state_dict(torch.load(checkpoint_path, map_location="cpu"))
        state_path = os.path.join(os.path.dirname(checkpoint_path), TRAINER_STATE_FILE)
        if os.path.exists(state_path):
            self.load_trainer_state(torch.load(state_path, map_location="cpu"))

    def trainer_state(self):
        opt = self.optimizer
        return {"iter": self.iter, "training_iteration": getattr(self, "_iteration", self.iter),
                "optimizer_kind": type(opt).__name__,
                "optimizer": opt.flat_state_dict() if isinstance(opt, PvaeAdam) else opt.state_dict(),
                "lr_scheduler": self.lr_scheduler.state_dict() if self.lr_scheduler else None}

    def load_trainer_state(self, st):
        self.iter = int(st["iter"])
        self._restored_iteration = int(st.get("training_iteration", st["iter"]))
        if st["optimizer_kind"] != type(self.optimizer).__name__:
            raise ValueError("checkpoint was written with optimizer %s, this trainer uses %s" % (st["optimizer_kind"], type(self.optimizer).__name__))
        if isinstance(self.optimizer, PvaeAdam):
            self.optimizer.load_flat_state_dict(st["optimizer"])
        else:
            self.optimizer.load_state_dict(st["optimizer"])
        if self.lr_scheduler and st.get("lr_scheduler") is not None:
            self.lr_scheduler.load_state_dict(st["lr_scheduler"])


class WorldModel(object):
    """View of a PhysicsVAE's world model: predicts s_{t+1} from (s_t, a_t) (`_world_model` + `forward_world`,
    rllib_model_torch.py:682-689, 839-844).  The reference has no class of this name; this is the thin
    alias the task framing asks for, sharing the PhysicsVAE's parameters."""

    def __init__(self, physics_vae):
        self.vae = physics_vae

    def __call__(self, s, a):
        return self.vae.forward_world(s, a)

    def parameters(self):
        return self.vae._world_model.parameters()

    def state_dict(self):
        return self.vae._world_model.state_dict()


class Motor(object):
    """View of the (s_body, s_task) -> z -> a stack: `_task_encoder` + reparameterisation + `_motor_decoder`
    (rllib_model_torch.py:638-668, 773-837)."""

    def __init__(self, physics_vae):
        self.vae = physics_vae

    def __call__(self, s_body, s_task, eps=None):
        obs = torch.cat([s_body, s_task], dim=-1)
        z_body, z_task, _ = self.vae.forward_encoder(obs, [], None, 0, eps=eps)
        logits, _ = self.vae.forward_decoder(z_body, z_task, [], None, 0)
        return logits[..., :self.vae.dim_action]

    def parameters(self):
        return list(self.vae._task_encoder.parameters()) + list(self.vae._motor_decoder.parameters())
