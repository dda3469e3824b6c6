"""bench.py -- transitions/sec of the PhysicsVAE training step on B200 (BASELINE.json metric), with the CPU reference arm.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--phase world|vae] [--batch B] [--config default|wide|loco]
  python bench.py --impl reference ...            the reference's own files (baseline/_ref, else the oracle port) on the host cores

A "step" is one mini-batch SGD step of the hot path: forward + loss + backward (libpvae_sm100 tcgen05 kernels) +
gradient all-reduce (N > 1) + fused Adam + refresh of the bf16 shadow weights, on a batch of synthetic transitions,
executed through the product's own API: `TrainModel.train_steps(K)` = K replays of the step graph that
`torch_models.TrainModel.step()` captures (physicsvae_b200/torch_models.py).
Headline workload = BASELINE.json configs[1]: world-model-only pretrain, dim_state_body 197 / dim_action 45, batch 65536
per GPU, bf16 operands with fp32 accumulation; the same line carries the VAE phase (configs[2]) under `phases.vae`, the
fp32-accurate bf16x3 mode, and -- on 8 GPUs -- configs[3] / configs[4] under `configs`.  Weak scaling: every rank holds
its own resident shard of 4*B transitions.  Prints ONE JSON line (rank 0).  DESIGN.md section 6 says how each field is computed.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

CPU_THREADS = 0
CPU_KIND = "auto"
CONFIGS = {
    "default": dict(dsb=197, da=45, z=32, te=(256, 2), md=(512, 3), wm=(1024, 2)),
    "wide": dict(dsb=512, da=128, z=32, te=(1024, 3), md=(1024, 3), wm=(1024, 3)),
    "loco": dict(dsb=361, da=54, z=32, te=(256, 2), md=(512, 3), wm=(1024, 2)),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch"],
                    help="b200: this repo; reference: the reference algorithm on the host cores; torch: stock PyTorch-CUDA "
                         "(cuBLAS bf16 nn.Linear autocast + autograd + fused Adam, graph-captured) as the library baseline")
    ap.add_argument("--phase", default="world", choices=["world", "vae"])
    ap.add_argument("--config", default="default", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=65536, help="mini-batch rows PER GPU")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="rows per step of the CPU arm (0: the workload's batch, shrunk only if the run would take too long)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU arm (0 = all cores; 1 with --cpu-sample 256 = configs[0])")
    ap.add_argument("--cpu-kind", default="auto", choices=["auto", "reference", "port"],
                    help="CPU arm: the reference's own files staged in baseline/_ref (auto when present) or the oracle port")
    ap.add_argument("--only-phase", action="store_true", help="measure just --phase (no second phase, no bf16x3 / extra-config legs)")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the power-capped (sustained clock) leg per phase")
    ap.add_argument("--all-configs", action="store_true", help="also run the configs[3] / configs[4] legs when N != 8")
    return ap.parse_args()


def flops_per_transition(dsb, da, z, te, md, wm):
    """Algorithmic FLOPs per transition (2*MAC on unpadded dims; bias / activation / loss elementwise work, the value branch and
    the reference's dead forward excluded), SURVEY.md section 8d.  te / md / wm: hidden widths of the three trained MLPs.
    world step = fwd + wgrad of every layer + dgrad of all but the first; VAE step = fwd of all three nets, wgrad + dgrad of
    encoder and decoder (the decoder's first layer also w.r.t. its z columns), dgrad through the frozen world model (its first
    layer w.r.t. the action columns only)."""
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


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def synthetic_arrays(cfg, n, seed):
    """s_0 ~ N(0,1), s_{t+1} = s_t + 0.05 N(0,1) inside episodes of 129 steps, a ~ U(-1,1);
    X float64 [n, 1, 2*dsb], Y float32 [n, 1, da] exactly like load_dataset_for_PhysicsVAE produces."""
    rng = np.random.default_rng(seed)
    T = 129
    E = (n + T - 2) // (T - 1)
    s = rng.standard_normal((E, 1, cfg["dsb"])) + np.cumsum(
        np.concatenate([np.zeros((E, 1, cfg["dsb"])), 0.05 * rng.standard_normal((E, T - 1, cfg["dsb"]))], axis=1), axis=1)
    X = np.concatenate([s[:, :-1], s[:, 1:]], axis=-1).reshape(-1, 1, 2 * cfg["dsb"])[:n]
    Y = rng.uniform(-1, 1, size=(n, 1, cfg["da"])).astype(np.float32)
    return np.ascontiguousarray(X), Y


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled in-process every ~2 ms (the
    timed region of a short run lasts only tens of milliseconds, `nvidia-smi -lms` cannot resolve that); if NVML is not
    importable the recipe's `nvidia-smi --query-gpu` loop is used instead."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc, self.halt, self.source = index, [], None, threading.Event(), None

    def _nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        self.source = "nvml"
        while not self.halt.is_set():
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            try:
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            self.rows.append((time.time(), sm, mx, pw, mask))
            time.sleep(0.002)

    def _smi(self):
        self.source = "nvidia-smi"
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                      "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        f = lambda v: float(v) if v.replace(".", "", 1).isdigit() else 0.0
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 7:
                mask = sum(bit for (name, bit), v in zip(self.REASONS.items(), c[3:7]) if v.lower().startswith("active"))
                self.rows.append((time.time(), f(c[0]), f(c[1]), f(c[2]), mask))

    def run(self):
        try:
            self._nvml()
        except Exception:
            try:
                self._smi()
            except Exception:
                pass

    def stop(self, t0, t1):
        self.halt.set()
        if self.proc:
            self.proc.terminate()
        self.join(timeout=1.0)
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        rows = inside or [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.05] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        mask = 0
        for r in rows:
            mask |= r[4]
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_min_mhz": min(r[1] for r in rows), "sm_max_mhz": rows[0][2],
                "samples": len(rows), "samples_inside_timed_region": len(inside), "power_w_max": max(r[3] for r in rows),
                "reasons": [n for n, bit in self.REASONS.items() if mask & bit], "source": self.source}


# --------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores -- the reference's OWN files when they are staged in baseline/_ref
# (oracle/stage_ref.py; they import unchanged under oracle/ref_stub), else the oracle port of the same algorithm
# --------------------------------------------------------------------------------------------------------------------
def _reference_stepper(cfg, phase, rows, seed):
    """(step function, kind): one mini-batch of `rows` transitions through compute_loss + backward + Adam + item.
    kind "reference": train_physics_vae.TrainModel.compute_loss of the unmodified reference (incl. its discarded full forward in
    the world phase) on a reference-built model; kind "port": oracle/pvae_oracle.py."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    want = CPU_KIND
    X, Y = synthetic_arrays(cfg, rows, seed + 1)
    if want in ("auto", "reference") and os.path.isfile(os.path.join(ref_dir, "train_physics_vae.py")):
        os.environ["PVAE_REFERENCE"] = ref_dir
        from oracle import refload
        refload.REFERENCE = ref_dir
        tpv, tm, rmt = refload.load()
        torch.manual_seed(seed)
        model = refload.build_reference_model(cfg["dsb"], cfg["da"], cfg["z"], tpv.gen_layers(*cfg["te"]), tpv.gen_layers(*cfg["md"]),
                                              tpv.gen_layers(*cfg["wm"]))
        world = phase == "world"
        model.set_learnable_task_encoder(not world); model.set_learnable_motor_decoder(not world); model.set_learnable_world_model(world)
        opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=0.0)          # torch_models.py:119-122
        x, y = torch.Tensor(X), torch.Tensor(Y)                                          # what DatasetBase + default collate hand over

        def step():
            opt.zero_grad()
            loss = refload.reference_compute_loss(model, x, y, world, kl_coeff=1.0, cyc_coeff=1e-3)
            loss.backward()
            opt.step()
            return loss.item()
        return step, "reference"
    if want == "reference":
        raise RuntimeError("baseline/_ref is not staged (python oracle/stage_ref.py where /root/reference exists)")
    from oracle import pvae_oracle as orc
    torch.manual_seed(seed)
    m = orc.OracleModel(cfg["dsb"], cfg["da"], cfg["z"], orc.gen_layers(*cfg["te"]), orc.gen_layers(*cfg["md"]), orc.gen_layers(*cfg["wm"]))
    tr = orc.OracleTrainer(m, X, Y, batch_size=rows, max_iter_world_model=0 if phase == "vae" else 10 ** 9)
    return tr.step, "port"


def cpu_arm(cfg, phase, rows, steps, warmup, seed=0, min_seconds=0.0, max_seconds=150.0):
    """`steps` timed mini-batch steps of `rows` transitions each on all host cores.  `min_seconds`: keep stepping until that much
    time was measured (the cpu_baseline leg wants 10-30 s of CPU work); `max_seconds`: if the requested run would take longer,
    the per-step sample shrinks (reported) so that the run still ends within minutes."""
    cores = CPU_THREADS or os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = _reference_stepper(cfg, phase, rows, seed)
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0
    while rows > 256 and t1 * (steps + warmup) > max_seconds:
        rows //= 2
        step, kind = _reference_stepper(cfg, phase, rows, seed)
        t0 = time.perf_counter()
        step()
        t1 = time.perf_counter() - t0
    for _ in range(max(warmup - 1, 0)):
        step()
    done, t0 = 0, time.perf_counter()
    while done < steps or (time.perf_counter() - t0) < min_seconds:
        step()
        done += 1
    dt = time.perf_counter() - t0
    return rows * done / dt, dt / done * 1e3, cores, rows, done, kind


KIND_TEXT = {"reference": "the reference's own train_physics_vae.TrainModel.compute_loss (unmodified files, baseline/_ref) + backward + Adam + item",
             "port": "oracle port of compute_loss + backward + Adam + item"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    value, ms, cores, rows, steps, kind = cpu_arm(cfg, args.phase, args.cpu_sample or args.batch, steps, warmup)
    sample = "%d steps of %d transitions each (%s phase, %s dims) of the batch-%d workload, fp32 torch-CPU on %d threads: %s" % (
        steps, rows, args.phase, args.config, args.batch, cores, KIND_TEXT[kind])
    line = {"impl": "reference", "metric": "transitions/sec (world-model+VAE step)", "value": value, "unit": "transitions/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "sample_rows_per_step": rows,
            "config": workload(args, cfg, args.phase, args.batch),
            "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload(args, cfg, phase, batch, dims=None):
    dims = dims or args.config
    return {"workload": "%s-phase train step, dim_state_body=%d dim_action=%d latent=%d, TE %dx%d MD %dx%d WM %dx%d, batch=%d per GPU" % (
        phase, cfg["dsb"], cfg["da"], cfg["z"], cfg["te"][0], cfg["te"][1], cfg["md"][0], cfg["md"][1], cfg["wm"][0], cfg["wm"][1],
        batch), "phase": phase, "dims": dims, "batch_per_gpu": batch, "global_batch": batch * args.gpus,
        "precision": args.precision, "parallelism": "dp%d" % args.gpus, "resident_rows_per_gpu": 4 * batch,
        "l2": "not flushed: the per-step working set (resident transitions + activations, >1 GB at batch 65536) exceeds the 126 MB L2"}


# --------------------------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------------------------
def make_trainer(cfg, B, precision, rank, world, phase="world", n_rows=None, local_shards=True, seed_base=1000):
    """The product's trainer (physicsvae_b200.train_physics_vae.TrainModel) on a synthetic resident dataset of 4*B rows."""
    from physicsvae_b200 import train_physics_vae as tp
    from physicsvae_b200 import torch_models as tm
    n_rows = n_rows or 4 * B

    class BenchTrainer(tp.TrainModel):
        dp_local_shards = local_shards

        def load_dataset(self, file):
            X, Y = synthetic_arrays(cfg, n_rows, seed=seed_base + (rank if local_shards else 0))
            return tm.DatasetBase(X, Y, normalize_x=False, normalize_y=False)

        def _local_rows(self, batch_size):
            return batch_size          # weak scaling: `batch` rows per GPU, each rank owns its shard of the global batch

    box = lambda n: tp.Box(low=-np.ones(n), high=np.ones(n), dtype=np.float64)
    custom = dict(tp.MODEL_CONFIG)
    custom.update(observation_space=box(2 * cfg["dsb"]), observation_space_body=box(cfg["dsb"]), observation_space_task=box(cfg["dsb"]),
                  action_space=box(cfg["da"]), engine_precision=precision, engine_max_batch=B)
    config = {"max_iter_world_model": 10 ** 9, "model": {"custom_model": "physics_vae", "custom_model_config": custom},
              "lr": 5e-4, "lr_schedule": "step", "lr_schedule_params": {"step_size": 50, "gamma": 0.7}, "weight_decay": 0.0,
              "dataset_train": ["synthetic"], "dataset_test": None, "loss": "MSE", "loss_test": "MSE", "batch_size": B,
              "latent_dim": cfg["z"], "latent_prior_type": "normal_zero_mean_one_std", "act_fn": "relu",
              "MD_width": cfg["md"][0], "MD_depth": cfg["md"][1], "TE_width": cfg["te"][0], "TE_depth": cfg["te"][1],
              "lookahead": 1, "world_model_width": cfg["wm"][0], "world_model_depth": cfg["wm"][1], "vae_kl_coeff": 1.0,
              "motor_decoder_a_rec_coeff": 1.0, "world_model_s_rec_coeff": 0.0, "vae_cycle_coeff": 1e-3,
              "engine_precision": precision, "noise_seed": 1234}
    torch.manual_seed(0)                      # same init on every rank (the trainer broadcasts rank 0's parameters anyway)
    tr = BenchTrainer(config)
    if phase == "vae":
        tr._enter_vae_phase()
    return tr


def barrier(world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, dev, world):
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def time_steps(tr, steps, warmup, world, dev, sampler=None):
    """K replays of the product's step graph through TrainModel.train_steps, CUDA events, barrier + synchronize on both sides,
    max over ranks.  Returns (ms for K steps, clock record or None)."""
    tr.train_steps(max(warmup, 3), restart=True)
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    tr.train_steps(steps)
    e1.record()
    barrier(world)
    t1 = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1), dev, world)
    return ms, (sampler.stop(t0, t1) if sampler else None)


def probe_kernel_ms(tr, kev, groups, per_group, step_ms, first=None):
    """Duration of the engine's launch sequence INSIDE the product's step graph: two external CUDA events recorded as nodes of
    the graph around pvae_{world,vae}_step.  The events only ever hold the LAST replay, and the step time is not stationary (the
    board ramps towards its power cap during a long timed region), so one replay cannot stand for the region's average.  What
    is stable is the SHARE of the step the launch sequence takes: each sample = in-graph duration of the last replay of a group
    of `per_group` back-to-back replays / the CUDA-event average step of the same group.  Reported: median share x `step_ms` (the
    timed region's average step).  `first` = the raw in-graph duration of the timed region's own last replay, kept for reference."""
    shares, raw = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(groups):
        e0.record()
        tr.train_steps(per_group)
        e1.record()
        torch.cuda.synchronize()
        v = kev[0].elapsed_time(kev[1])
        g = e0.elapsed_time(e1) / per_group
        if v > 0 and g > 0:
            shares.append(v / g)
            raw.append(v)
    if not shares:
        return first, {"last_replay_of_timed_region_ms": first}
    srt = sorted(shares)
    share = srt[len(srt) // 2] if len(srt) % 2 else 0.5 * (srt[len(srt) // 2 - 1] + srt[len(srt) // 2])
    return share * step_ms, {"share_of_step": {"n": len(srt), "min": srt[0], "median": share, "max": srt[-1], "replays_per_group": per_group},
                             "raw_last_replay_ms": {"timed_region": first, "groups_min": min(raw), "groups_max": max(raw)}}


def measure_phase(args, cfg, tr, phase, B, world, rank, dev, steps, warmup, local, dims=None):
    """Burst-clock leg of one phase: value through the product API (K replays of the trainer's own step graph) and the in-graph
    duration of the engine's launch sequence."""
    from physicsvae_b200 import _abi
    fl_world, fl_vae = flops_per_transition(cfg["dsb"], cfg["da"], cfg["z"], [cfg["te"][0]] * cfg["te"][1],
                                            [cfg["md"][0]] * cfg["md"][1], [cfg["wm"][0]] * cfg["wm"][1])
    flops_step = (fl_world if phase == "world" else fl_vae) * B
    if phase == "vae" and tr.world_phase:
        tr._enter_vae_phase()
    # launches of one step, counted on an eager mini-batch (the graph replays the same sequence)
    tr._bind("train")
    if tr.model._weights_dirty:
        tr.model.sync_weights()                                      # (one-off shadow-operand build: not part of a step)
    n0 = _abi.launch_count()
    tr.train_batch(0, B)
    # the eager path sets the cursor (and, VAE phase, the noise counter) per call; a replay advances the cursor on the device instead
    launches_per_step = _abi.launch_count() - n0 - (2 if phase == "vae" else 1) + 1
    try:
        kev = (torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True))
    except TypeError:
        kev = None
    tr._graph_probe = kev
    ms, clocks = time_steps(tr, steps, warmup, world, dev, ClockSampler(local) if rank == 0 else None)
    # the events now hold the LAST replay of the timed region (barrier() synchronised): the first kernel sample
    first = kev[0].elapsed_time(kev[1]) if kev else None
    kernel_ms, kernel_samples = probe_kernel_ms(tr, kev, 10, max(5, min(steps, 20)), ms / steps, first if first and first > 0 else None) if kev else (None, None)
    # not GEMMs: loss finalisation, fused Adam (one per trained net), cursor advance; VAE phase also reparameterisation fwd / bwd;
    # N > 1 with the symmetric gradient pool: the library's peer-memory all-reduce kernel
    from physicsvae_b200 import parallel as _par
    gemm_launches = launches_per_step - {"world": 3, "vae": 6}[phase] - (1 if world > 1 and _par.allreduce_kind()["kind"].startswith("symm") else 0)
    return {"phase": phase, "B": B, "value": B * world * steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps, "clocks": clocks,
            "kernel_ms": kernel_ms, "kernel_samples": kernel_samples, "kev": kev, "flops_step": flops_step, "launches_per_step": int(launches_per_step),
            "gemm_launches": int(gemm_launches), "sustained": None, "config": workload(args, cfg, phase, B, dims), "dims": dims or args.config}


def measure_sustained(res, tr, world, rank, dev, local, seconds):
    """>= `seconds` of back-to-back steps (the 1 kW power cap pulls the SM clock down), then kernel samples taken as the last replay
    of 50-step groups so that the regime holds."""
    B, kev = res["B"], res["kev"]
    per = res["ms_per_step"]
    n_sus = int(max(res["steps"], seconds * 1e3 / per))
    sampler = ClockSampler(local) if rank == 0 else None
    tr.train_steps(n_sus // 2)                                       # ramp into the power-capped regime
    if sampler:
        sampler.start()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    tr.train_steps(n_sus)
    e1.record()
    barrier(world)
    t1 = time.time()
    sms = max_over_ranks(e0.elapsed_time(e1), dev, world)
    first = kev[0].elapsed_time(kev[1]) if kev else None
    k_sus = probe_kernel_ms(tr, kev, 8, 50, sms / n_sus, first if first and first > 0 else None)[0] if kev else None
    res["sustained"] = {"steps": n_sus, "ms_per_step": sms / n_sus, "value": B * world * n_sus / (sms * 1e-3), "kernel_ms_per_step": k_sus,
                        "clocks": sampler.stop(t0, t1) if sampler else None}


def finish_phase(args, res, tr):
    """Roofline object of a measured phase: fractions against the measured cuBLAS bf16 peak of the matching clock regime."""
    tr._graph_probe = None
    pk, pk_src = peaks()
    kernel_ms, sus, clocks, flops_step = res["kernel_ms"], res["sustained"], res["clocks"], res["flops_step"]
    traffic = None
    try:      # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of the same workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s/%s/%d" % (res["dims"], res["phase"], res["B"])
        if key in tj:
            traffic = tj[key]["dram_bytes_per_launch"]
    except Exception:
        pass
    roof = None
    if kernel_ms:
        gl = max(res["gemm_launches"], 1)
        ach = flops_step / (kernel_ms * 1e-3) / 1e12
        ach_s = flops_step / (sus["kernel_ms_per_step"] * 1e-3) / 1e12 if sus and sus.get("kernel_ms_per_step") else None
        burst_regime = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.95 * clocks["sm_max_mhz"])
        peak = pk["bf16_tflops"] if burst_regime else pk["bf16_tflops_sustained"]
        roof = {"bound": "tensor", "achieved": ach, "unit": "TFLOP/s", "peak": peak, "frac": ach / peak,
                "regime": "burst (median SM clock >= 0.95 max during the timed region)" if burst_regime else "sustained (SM clock below 0.95 max during the timed region)",
                "frac_burst": ach / pk["bf16_tflops"], "peak_burst": pk["bf16_tflops"],
                "achieved_sustained": ach_s, "frac_sustained": (ach_s / pk["bf16_tflops_sustained"]) if ach_s else None,
                "peak_sustained": pk["bf16_tflops_sustained"], "peak_source": pk_src + " (MEASURED_PEAKS.json: cuBLAS bf16 burst / sustained)",
                "traffic": traffic, "kernel": "pvae_gemm_kernel", "launches_per_step": res["gemm_launches"],
                "avg_launch_ms": kernel_ms / gl, "algorithmic_flops_per_launch": flops_step / gl,
                "kernel_ms_per_step": kernel_ms, "kernel_ms_samples": res.get("kernel_samples"), "algorithmic_flops_per_step": flops_step,
                "timing": "share of the step between two external CUDA events recorded as nodes of the product's step graph around the engine's "
                          "launch sequence (median over groups of replays) x the timed region's average step",
                "whole_step_tflops": flops_step / (res["ms_per_step"] * 1e-3) / 1e12}
    out = {k: res[k] for k in ("value", "ms_per_step", "steps", "clocks", "sustained", "launches_per_step", "config")}
    out["roofline"] = roof
    out["loss_after"] = float(tr.engine.loss[0].item())
    return out


def dp_equivalence_check(cfg, world, rank, dev):
    """N > 1, before anything is timed: ONE data-parallel step of the real path (every rank its slice of a 4096-row global
    mini-batch, NCCL all-reduce of the [gradients | loss] range) against the same global batch on rank 0 alone, bf16x3."""
    import torch.distributed as dist
    from physicsvae_b200 import parallel
    Bg = 4096
    out = {}
    for phase in ("world", "vae"):
        # trainer batch = the GLOBAL mini-batch; every rank holds the same Bg rows and takes its slice (engine capacity Bg rows, so
        # that rank 0 can also run the whole batch alone)
        tr = make_trainer(cfg, Bg, "bf16x3", rank, world, phase=phase, n_rows=Bg, local_shards=False, seed_base=555)
        tr.model.latent_prior_noise = False
        tr.batch_loss(0, Bg)
        dp = tr.model.reduce_range(phase == "world").detach().clone()
        torch.cuda.synchronize()
        err = 0.0
        if rank == 0:
            parallel.set_replica_mode(True)                   # world_size() == 1: the whole batch here, no collective
            tr.batch_loss(0, Bg)
            one = tr.model.reduce_range(phase == "world").detach().clone()
            parallel.set_replica_mode(False)
            err = float((dp.double() - one.double()).norm() / one.double().norm())
        t = torch.tensor([err], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[phase] = float(t.item())
        if parallel.allreduce_kind()["kind"].startswith("symm"):
            # the library's peer-memory kernel against NCCL on the same rank-dependent data
            rng = tr.model.reduce_range(phase == "world")
            rng.copy_(torch.randn(rng.numel(), device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank)))
            want = rng.detach().clone()
            dist.all_reduce(want, op=dist.ReduceOp.AVG)
            parallel.allreduce_avg_([rng])
            torch.cuda.synchronize()
            d = torch.tensor([float((rng.double() - want.double()).norm() / want.double().norm())], device=dev, dtype=torch.float64)
            dist.all_reduce(d, op=dist.ReduceOp.MAX)
            out[phase + "_peer_kernel_vs_nccl"] = float(d.item())
        del tr
        torch.cuda.empty_cache()
    return out


def replica_checksum(tr, world, dev):
    """After the timed steps: every rank must hold bit-identical parameters (replicated Adam on all-reduced gradients)."""
    import torch.distributed as dist
    from physicsvae_b200.engine import NET_NAMES
    sums = torch.stack([tr.model.flat_params(n).double().sum() for n in NET_NAMES] +
                       [tr.model.flat_params(n).double().abs().sum() for n in NET_NAMES])
    if world == 1:
        return True
    allv = [torch.empty_like(sums) for _ in range(world)]
    dist.all_gather(allv, sums)
    return all(bool(torch.equal(v, allv[0])) for v in allv)


def run_b200(args, cfg):
    import torch.distributed as dist
    from physicsvae_b200 import _abi, parallel
    from physicsvae_b200 import train_physics_vae as tp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1:
        # one loader process per GPU: keep it (and the pinned staging memory it is about to allocate) on the GPU's NUMA node
        numa = parallel.bind_to_gpu_numa_node(local) if os.environ.get("PVAE_NUMA_BIND", "1") != "0" else {"bound": False, "why": "PVAE_NUMA_BIND=0"}
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    B, phase = args.batch, args.phase
    steps, warmup = args.steps, max(args.warmup, 3)
    n0_launch = _abi.launch_count() if os.path.exists(_abi.LIB_PATH) else 0

    dp_check = None
    if world > 1:
        errs = dp_equivalence_check(cfg, world, rank, dev)
        dp_check = {"grad_rel_l2_vs_single_rank": errs, "ok": all(v < 1e-5 for v in errs.values())}      # (peer-kernel vs NCCL: < 1e-5 too)

    # burst-clock legs of both phases first (a sustained leg leaves the GPU power-capped), then the sustained legs
    tr = make_trainer(cfg, B, args.precision, rank, world, phase=phase)
    res = {phase: measure_phase(args, cfg, tr, phase, B, world, rank, dev, steps, warmup, local)}
    trainers = {phase: tr}
    if not args.only_phase:
        other = "vae" if phase == "world" else "world"
        time.sleep(1.0)
        trainers[other] = make_trainer(cfg, B, args.precision, rank, world, phase=other)
        res[other] = measure_phase(args, cfg, trainers[other], other, B, world, rank, dev, steps, warmup, local)
    if args.sustained_seconds > 0:
        for ph in res:
            measure_sustained(res[ph], trainers[ph], world, rank, dev, local, args.sustained_seconds)
    done = {ph: finish_phase(args, res[ph], trainers[ph]) for ph in res}
    main = done[phase]
    phases = {ph: v for ph, v in done.items() if ph != phase}
    for ph in list(trainers):
        if ph != phase:
            del trainers[ph]
    torch.cuda.empty_cache()
    replicas_identical = replica_checksum(tr, world, dev)
    if dp_check is not None:
        dp_check["replicas_identical_after_timed_steps"] = replicas_identical
        dp_check["ok"] = dp_check["ok"] and replicas_identical
    gpu_launches_timed = int(main["launches_per_step"] * steps)
    exchange_us = None
    if world > 1:
        # the gradient exchange on its own: the full [gradients | loss] range of the headline phase, 50 back-to-back calls
        rng_ = tr.model.reduce_range(phase == "world")
        for _ in range(5):
            parallel.allreduce_avg_([rng_])
        barrier(world)
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(50):
            parallel.allreduce_avg_([rng_])
        x1.record()
        barrier(world)
        exchange_us = {"us_per_call": max_over_ranks(x0.elapsed_time(x1), dev, world) / 50 * 1e3, "bytes": int(rng_.numel() * 4),
                       "what": "averaging exchange of the whole [gradients | loss] range of the %s phase, alone, back to back" % phase}

    # ---- end-to-end legs (world phase of the headline workload unless --phase vae) ------------------------------------------
    # (1) "e2e": every step uploads ITS inputs from pinned host memory and reads the loss back.  The mini-batch travels in the
    #     compact form of the dataset (every state once + per-transition index, TrainModel.compute_loss_episodes): half the bytes
    #     of the expanded x = [s_t | s_{t+1}] the reference's DataLoader hands over.  Double-buffered on a copy stream.
    tr_e = tr
    T = 129
    E = (B + T - 2) // (T - 1)
    rng = np.random.default_rng(77 + rank)
    st = (rng.standard_normal((E, 1, cfg["dsb"])) + np.cumsum(np.concatenate(
        [np.zeros((E, 1, cfg["dsb"])), 0.05 * rng.standard_normal((E, T - 1, cfg["dsb"]))], axis=1), axis=1)).reshape(E * T, cfg["dsb"])
    first = (np.arange(E)[:, None] * T + np.arange(T - 1)[None, :]).reshape(-1)[:B].astype(np.int64)
    sh = torch.from_numpy(st.astype(np.float32))
    ah = torch.from_numpy(rng.uniform(-1, 1, size=(E * T, cfg["da"])).astype(np.float32))
    if args.precision == "bf16":
        # the loader keeps the dataset pinned in the engine's operand precision (converted once, when the dataset is loaded: the same
        # round-to-nearest bf16 values the device-side ingest produces from fp32) -- half the bytes per step again
        sh, ah = sh.bfloat16(), ah.bfloat16()
    sh, ah = sh.pin_memory(), ah.pin_memory()
    fh = torch.from_numpy(first).pin_memory()
    h2d = sh.numel() * sh.element_size() + ah.numel() * ah.element_size() + fh.numel() * 8
    copy_stream = torch.cuda.Stream()
    bufs = [tuple(torch.empty_like(t, device=dev) for t in (sh, ah, fh)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            for d, s_ in zip(bufs[i % 2], (sh, ah, fh)):
                d.copy_(s_, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_step(i, last=False):
        if not last:
            prefetch(i + 1)                  # buffer (i + 1) % 2 was last read by step i - 1, which loss.item() has retired
        torch.cuda.current_stream().wait_event(ready[i % 2])
        s_, a_, f_ = bufs[i % 2]
        return tr_e.train_on_episodes(s_, a_, f_).item()       # dataset build + forward + loss + backward + [exchange] + Adam: one graph launch
    e2e_steps = max(3, min(steps, 20))
    prefetch(0)
    for i in range(3):
        e2e_step(i)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 3 + e2e_steps):
        e2e_step(i, last=(i == 2 + e2e_steps))
    e1.record()
    barrier(world)
    ems = max_over_ranks(e0.elapsed_time(e1), dev, world)
    e2e = {"value": B * world * e2e_steps / (ems * 1e-3), "unit": "transitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "steps": e2e_steps, "ms_per_step": ems / e2e_steps,
           "upload_dtype": str(sh.dtype).replace("torch.", ""),
           "api": "double-buffered pinned-host loader (unique states + index, %s) -> TrainModel.train_on_episodes (one captured graph per staging buffer: "
                  "device-side dataset build + compute_loss + backward + Adam) + loss.item()" % str(sh.dtype).replace("torch.", "")}
    # (2) "e2e_resident": what the reference's Trainable.step() does -- whole epochs over the dataset -- with the dataset uploaded
    #     once at setup: TrainModel.step() = graph replays + one loss read-back per epoch (+ the LR scheduler)
    epochs = max(2, min(steps // 4, 10))
    tr_e.step()
    barrier(world)
    e0.record()
    for _ in range(epochs):
        r = tr_e.step()
    e1.record()
    barrier(world)
    rms = max_over_ranks(e0.elapsed_time(e1), dev, world)
    e2e_res = {"value": B * world * 4 * epochs / (rms * 1e-3), "unit": "transitions/s", "epochs": epochs, "steps_per_epoch": 4,
               "ms_per_step": rms / (4 * epochs), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 1, "mean_train_loss": r["mean_train_loss"],
               "api": "TrainModel.step() over the resident dataset (uploaded once by setup()); one loss read-back per epoch"}
    if tr_e is not tr:
        del tr_e

    # ---- secondary legs ----------------------------------------------------------------------------------------------------
    extra = {}
    if not args.only_phase and args.precision == "bf16":
        # the fp32-accurate arithmetic (hi/lo split bf16, three tensor-core passes): the mode that meets rtol=1e-3 / atol=1e-5
        del tr
        torch.cuda.empty_cache()
        t3 = make_trainer(cfg, B, "bf16x3", rank, world, phase=phase)
        k3 = max(5, min(steps, 20))
        ms3, _ = time_steps(t3, k3, 3, world, dev)
        extra["bf16x3"] = {"value": B * world * k3 / (ms3 * 1e-3), "ms_per_step": ms3 / k3, "steps": k3, "phase": phase,
                           "note": "fp32-accurate mode (tests: rtol=1e-3 / atol=1e-5 against the fp32 oracle at this batch size)"}
        del t3
        torch.cuda.empty_cache()
        tr = None
    if not args.only_phase:
        # the CLI's default mini-batch (train_physics_vae.py --batch_size 256 = BASELINE.json configs[0]'s batch): a launch-bound step,
        # which is what the captured graph is for -- the same trainer timed replaying its graph and launching every kernel eagerly
        tr = None
        torch.cuda.empty_cache()
        tb = make_trainer(cfg, 256, args.precision, rank, world, phase=phase, n_rows=256 * 64)
        kb = 500
        msb, _ = time_steps(tb, kb, 20, world, dev)
        for i in range(10):
            tb.train_batch(256 * (i % 64), 256 * (i % 64) + 256)
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200):
            tb.train_batch(256 * (i % 64), 256 * (i % 64) + 256)
        e1.record()
        barrier(world)
        mse = max_over_ranks(e0.elapsed_time(e1), dev, world)
        extra["batch256"] = {"value": 256 * world * kb / (msb * 1e-3), "ms_per_step": msb / kb, "steps": kb, "phase": phase,
                             "eager_ms_per_step": mse / 200, "eager_value": 256 * world * 200 / (mse * 1e-3),
                             "note": "CLI default batch 256 per GPU: TrainModel.train_steps (captured graph) vs TrainModel.train_batch (eager launches)"}
        del tb
        torch.cuda.empty_cache()
    cfgs = {}
    if not args.only_phase and (world == 8 or args.all_configs):
        # BASELINE.json configs[3]: full VAE, global batch 262144 on 8 GPUs (32768 rows per GPU);
        # configs[4]: wide dims (dsb 512, da 128, hidden 3x1024), global batch 131072 on 8 GPUs (16384 rows per GPU), both phases
        tr = None
        torch.cuda.empty_cache()
        kc = max(20, min(steps, 100))
        for name, cname, ph, rows in (("cfg4_vae_262144_global", "default", "vae", 262144 // 8), ("cfg5_wide_131072_global_world", "wide", "world", 131072 // 8),
                                      ("cfg5_wide_131072_global_vae", "wide", "vae", 131072 // 8)):
            c = CONFIGS[cname]
            tc = make_trainer(c, rows, args.precision, rank, world, phase=ph)
            r_ = finish_phase(args, measure_phase(args, c, tc, ph, rows, world, rank, dev, kc, 5, local, dims=cname), tc)
            cfgs[name] = {k: r_[k] for k in ("value", "ms_per_step", "steps", "clocks", "roofline", "config", "launches_per_step")}
            del tc
            torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": "transitions/sec (world-model+VAE step)", "value": main["value"], "unit": "transitions/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3(fp32-accurate)",
                "data": "synthetic", "config": main["config"], "clocks": main["clocks"], "e2e": e2e, "e2e_resident": e2e_res,
                "gpu_launches": gpu_launches_timed, "launches_per_step": main["launches_per_step"],
                "api": "physicsvae_b200.train_physics_vae.TrainModel.train_steps (replays of the step graph TrainModel.step() captures)",
                "cuda_graph": True, "roofline": main["roofline"], "sustained": main["sustained"], "loss_after": main["loss_after"],
                "phases": {k: {kk: v[kk] for kk in ("value", "ms_per_step", "steps", "clocks", "roofline", "sustained", "launches_per_step", "config", "loss_after")}
                           for k, v in phases.items()},
                "library_launches_total": int(_abi.launch_count() - n0_launch)}
        line.update(extra)
        if cfgs:
            line["configs"] = cfgs
        line["allreduce"] = dict(parallel.allreduce_kind(), exchange=exchange_us, overlapped_with_backward=bool(os.environ.get("PVAE_OVERLAP", "0") == "1" and world > 1
                                                                                          and parallel.allreduce_kind()["kind"].startswith("symm")))
        if dp_check is not None:
            line["dp_check"] = dp_check
        if numa is not None:
            line["numa"] = numa          # (rank 0's record: where its loader process and pinned staging memory live)
        if not args.no_cpu_baseline:
            v, cms, cores, rows, nsteps, kind = cpu_arm(cfg, phase, args.cpu_sample or 4096, 3, 1, min_seconds=10.0)
            line["cpu_baseline"] = {"value": v, "unit": "transitions/s", "cores": cores, "kind": kind, "ms_per_step": cms,
                                    "sample": "%d steps of %d transitions (%s phase, ~10 s) of the batch-%d workload, fp32 torch-CPU on %d threads: %s" % (
                                        nsteps, rows, phase, B, cores, KIND_TEXT[kind])}
        print(json.dumps(line), flush=True)
    # Teardown: a CUDA graph that captured NCCL work keeps the communicator busy and destroy_process_group() can block on
    # it forever; every rank is done with collectives here, so leave without the collective teardown.
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


# --------------------------------------------------------------------------------------------------------------------
# library baseline: the same training step written with stock PyTorch on the same GPU (not part of the driver contract)
# --------------------------------------------------------------------------------------------------------------------
def run_torch(args, cfg):
    """World / VAE step with nn.Linear under bf16 autocast (cuBLASLt), autograd, torch.optim.Adam(fused, capturable), the whole
    step captured in one CUDA graph; inputs pre-concatenated and resident in bf16 (the friendliest setting for the library)."""
    import torch.nn as nn
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, phase = args.batch, args.phase
    dsb, da, z = cfg["dsb"], cfg["da"], cfg["z"]

    def mlp(i, w, d, o):
        layers, k = [], i
        for _ in range(d):
            layers += [nn.Linear(k, w), nn.ReLU()]
            k = w
        return nn.Sequential(*layers, nn.Linear(k, o)).to(dev)
    torch.manual_seed(0)
    te, md, wm = mlp(2 * dsb, *cfg["te"], 2 * z), mlp(dsb + z, *cfg["md"], da), mlp(dsb + da, *cfg["wm"], dsb)
    s1 = torch.randn(B, dsb, device=dev, dtype=torch.bfloat16)
    s2 = (s1.float() + 0.05 * torch.randn(B, dsb, device=dev)).bfloat16()
    a = (torch.rand(B, da, device=dev) * 2 - 1).bfloat16()
    s1a, s12 = torch.cat([s1, a], 1), torch.cat([s1, s2], 1)
    train = list(wm.parameters()) if phase == "world" else list(te.parameters()) + list(md.parameters())
    if phase == "vae":
        for p_ in wm.parameters():
            p_.requires_grad_(False)
    opt = torch.optim.Adam(train, lr=torch.tensor(5e-4, device=dev), fused=True, capturable=True)
    mse = nn.functional.mse_loss

    def step():
        opt.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if phase == "world":
                loss = mse(wm(s1a).float(), s2.float())
            else:
                h = te(s12).float()
                mu, lv = h[:, :z], h[:, z:]
                zz = mu + torch.randn_like(mu) * torch.exp(0.5 * lv)
                act = md(torch.cat([s1, zz.bfloat16()], 1))
                fut = wm(torch.cat([s1, act], 1))
                kl = torch.mean(-0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp(), dim=1))
                loss = mse(act.float(), a.float()) + kl + 1e-3 * mse(fut.float(), s2.float())
        loss.backward()
        opt.step()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            step()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(max(args.warmup, 3)):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    fl = flops_per_transition(dsb, da, z, [cfg["te"][0]] * cfg["te"][1], [cfg["md"][0]] * cfg["md"][1], [cfg["wm"][0]] * cfg["wm"][1])
    flops = (fl[0] if phase == "world" else fl[1]) * B
    pk, _ = peaks()
    print(json.dumps({"impl": "torch", "metric": "transitions/sec (world-model+VAE step)", "value": B / (ms * 1e-3), "unit": "transitions/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                      "dtype": "bf16 autocast (cuBLASLt) + fp32 masters", "data": "synthetic", "config": workload(args, cfg, args.phase, args.batch), "cuda_graph": True,
                      "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": pk["bf16_tflops_sustained"],
                                   "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                                   "note": "whole step (library kernels are not separable by event inside the graph)"}}), flush=True)


if __name__ == "__main__":
    a = parse()
    c = CONFIGS[a.config]
    CPU_THREADS = a.cpu_threads
    CPU_KIND = a.cpu_kind
    if a.impl == "torch":
        run_torch(a, c)
    elif a.impl == "reference":
        run_reference(a, c)
    else:
        run_b200(a, c)
