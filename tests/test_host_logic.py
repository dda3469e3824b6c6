"""Host-side mirror of the reference's Python surface: CLI flags, layer-spec DSL, trainer config, dataset builder, module
signatures and state-dict keys, seeded init, freezing, checkpoint file formats, loaders, data-parallel sharding.  CPU only."""
import argparse
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import pvae_oracle as orc
from physicsvae_b200 import _abi, parallel
from physicsvae_b200 import rllib_model_torch as pm
from physicsvae_b200 import torch_models as tm
from physicsvae_b200 import train_physics_vae as tp

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _small_model(prior="normal_zero_mean_one_std"):
    g = np.load(os.path.join(G, "ref_small_step.npz"))
    dsb, da, z = int(g["dsb"]), int(g["da"]), int(g["z"])
    te, md, wm = tp.gen_layers(48, 2), tp.gen_layers(64, 3), tp.gen_layers(96, 2)
    for l in (te, md, wm):
        l[-1]["init_weight"] = {"name": "normc", "std": 0.3}
    box = lambda n: tp.Box(low=-np.ones(n), high=np.ones(n), dtype=np.float64)
    custom = dict(pm.PhysicsVAE.DEFAULT_CONFIG)
    custom.update(observation_space=box(2 * dsb), observation_space_body=box(dsb), observation_space_task=box(dsb), action_space=box(da),
                  task_encoder_output_dim=z, task_encoder_layers=te, motor_decoder_layers=md, world_model_layers=wm,
                  value_fn_layers=tp.gen_layers(48, 2), latent_prior_type=prior)
    torch.manual_seed(0)
    m = pm.PhysicsVAE(box(2 * dsb), box(da), 2 * da, {"custom_model_config": custom}, "physics_vae")
    return g, m


def test_cli_flags_and_defaults():
    a = tp.arg_parser().parse_args(["--data_train", "d.pkl"])
    assert (a.max_iter, a.max_iter_world_model, a.lr, a.lr_schedule, a.batch_size, a.checkpoint_freq, a.latent_dim) == \
        (100, 0, 0.0005, "step", 256, 100, 32)
    assert a.vae_kl_coeff == [1.0] and a.vae_cycle_coeff == [1e-3] and a.latent_prior_type == ["normal_zero_mean_one_std"]
    assert a.local_dir == "~/ray_results" and a.data_test is None and a.num_data is None
    # list flags append to their defaults
    b = tp.arg_parser().parse_args(["--data_train", "a", "--data_train", "b", "--vae_kl_coeff", "0.5"])
    assert b.data_train == ["a", "b"] and b.vae_kl_coeff == [1.0, 0.5]
    with pytest.raises(SystemExit):
        tp.arg_parser().parse_args([])


def test_gen_layers_dsl():
    assert tp.gen_layers(256, 2) == orc.gen_layers(256, 2) == pm.DEFAULT_FC_256X2
    l = tp.gen_layers(8, 1, act_hidden="elu", add_softmax=True)
    assert l[-1] == {"type": "softmax"} and l[0]["activation"] == "elu" and l[1]["hidden_size"] == "output"
    with pytest.raises(AssertionError):
        tp.gen_layers(0, 1)


def test_activation_registry():
    assert pm.get_activation_fn("linear") is None and pm.get_activation_fn(None) is None
    assert pm.get_activation_fn("relu") is torch.nn.ReLU and pm.get_activation_fn("elu") is torch.nn.ELU
    with pytest.raises(ValueError):
        pm.get_activation_fn("gelu")
    with pytest.raises(NotImplementedError):
        pm.get_initializer({"name": "orthogonal"})
    with pytest.raises(NotImplementedError):
        tm.get_loss_fn("Huber")


def test_state_dict_keys_and_seeded_init_match_reference():
    g, m = _small_model()
    ref = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())            # same names, same order, no log_std entry (constant type)
    for k in ref:
        assert np.array_equal(sd[k].numpy(), ref[k]), k    # same RNG consumption as the reference constructor
    assert len(list(m.parameters())) == 2 * (3 + 4 + 3 + 3)


def test_loco_checkpoint_key_layout():
    g = np.load(os.path.join(G, "ref_loco_ckpt.npz"))
    box = lambda n: tp.Box(low=-np.ones(n), high=np.ones(n), dtype=np.float64)
    custom = dict(pm.PhysicsVAE.DEFAULT_CONFIG)
    custom.update(observation_space=box(722), observation_space_body=box(361), observation_space_task=box(361), action_space=box(54))
    m = pm.PhysicsVAE(box(722), box(54), 108, {"custom_model_config": custom}, "physics_vae")
    assert list(m.state_dict().keys()) == g["keys"].tolist()
    assert sum(p.numel() for p in m.parameters()) == 3118816


def test_model_attributes_and_errors():
    g, m = _small_model()
    assert (m.dim_state, m.dim_state_body, m.dim_state_task, m.dim_action) == (74, 37, 37, 11)
    assert m.latent_prior_noise is True and m.get_initial_state() == []
    with pytest.raises(AssertionError):
        m.value_function()
    # no eager fallback: CPU tensors cannot be run
    x = torch.zeros(4, 74)
    with pytest.raises(_abi.PvaeError):
        m(input_dict={"obs": x, "obs_flat": x}, state=None, seq_lens=None)
    with pytest.raises(_abi.PvaeError):
        m._world_model(torch.zeros(2, 48))
    # AppendLogStd semantics
    app = m._motor_decoder._model[-1]
    out = app(torch.zeros(3, 11))
    assert out.shape == (3, 22) and torch.allclose(out[:, 11:], torch.full((3, 11), float(np.log(0.1))))
    m.set_exploration_std(0.05)
    assert torch.allclose(app(torch.zeros(1, 11))[:, 11:], torch.full((1, 11), float(np.log(0.05))))
    with pytest.raises(AssertionError):
        box = tp.Box(low=-np.ones(4), high=np.ones(4))
        pm.PhysicsVAE(box, box, 3, {}, "x")


def test_broken_upstream_priors_raise_clearly():
    for prior in ("normal_state_mean_one_std", "hypersphere_uniform", "bogus"):
        with pytest.raises(NotImplementedError):
            _small_model(prior)
    g, m = _small_model(False)
    assert m._task_encoder.layer_spec()[-1][0] == 8        # no prior: the encoder emits z directly


def test_freezing_and_partial_checkpoints(tmp_path):
    g, m = _small_model()
    m.set_learnable_task_encoder(False)
    m.set_learnable_motor_decoder(False)
    assert all(not p.requires_grad for p in m._task_encoder.parameters())
    assert all(p.requires_grad for p in m._world_model.parameters())
    f = {n: str(tmp_path / (n + ".pt")) for n in ("model", "task_encoder", "motor_decoder", "world_model")}
    m.save_weights(f["model"]); m.save_weights_task_encoder(f["task_encoder"])
    m.save_weights_motor_decoder(f["motor_decoder"]); m.save_weights_world_model(f["world_model"])
    assert list(torch.load(f["task_encoder"]).keys()) == ["task_encoder"]
    assert list(torch.load(f["world_model"]).keys())[0] == "_model.0._model.0.weight"
    torch.manual_seed(1)
    g2, m2 = _small_model()
    torch.manual_seed(2)
    for p in m2.parameters():
        p.data.normal_()
    m2.load_weights_world_model(f["world_model"]); m2.load_weights_task_encoder(f["task_encoder"]); m2.load_weights_motor_decoder(f["motor_decoder"])
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        if not k.startswith("_value_branch"):
            assert torch.equal(a, b), k
    assert not m2._world_model.training            # load_weights* switch the sub-module to eval (appendix B.3)
    m2.load_weights(f["model"])
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_dataset_builder_matches_reference_fixture(tmp_path):
    g = np.load(os.path.join(G, "ref_dataset.npz"))
    data = pickle.loads(g["pickle_bytes"].tobytes())
    f = str(tmp_path / "demo.pkl")
    pickle.dump(data, open(f, "wb"))
    ds = tp.load_dataset_for_PhysicsVAE([f])
    assert ds.X.dtype == np.float64 and np.array_equal(ds.X, g["X"]) and np.array_equal(ds.Y, g["Y"]) and ds.Y.dtype == g["Y"].dtype
    ds = tp.load_dataset_for_PhysicsVAE([f], num_samples=20)
    assert np.array_equal(ds.X, g["X_cap"]) and np.array_equal(ds.Y, g["Y_cap"])
    ds = tp.load_dataset_for_PhysicsVAE([f], lookahead=2, cond="rel", use_a_gt=True)
    assert np.array_equal(ds.X, g["X_rel"]) and np.array_equal(ds.Y, g["Y_rel"])
    # merge of two files concatenates episodes; mismatching metadata asserts
    ds2 = tp.load_dataset_for_PhysicsVAE([f, f])
    assert len(ds2) == 2 * len(g["X"])
    bad = dict(data); bad["dim_action"] = 99
    f2 = str(tmp_path / "bad.pkl")
    pickle.dump(bad, open(f2, "wb"))
    with pytest.raises(AssertionError):
        tp.merge_dataset([f, f2])
    with pytest.raises(AssertionError):
        tp.load_dataset_for_PhysicsVAE([])
    with pytest.raises(NotImplementedError):
        tp.episodes_to_transitions(data["episodes"], cond="delta")
    # __getitem__ hands out float32 tensors like the reference's DatasetBase
    x0, y0 = tp.load_dataset_for_PhysicsVAE([f])[0]
    assert x0.dtype == torch.float32 and x0.shape == (1, 10) and y0.shape == (1, 3)
    # sequential loader with a short last batch
    loader = tm.ResidentLoader(tp.load_dataset_for_PhysicsVAE([f]), 7, None)
    assert [hi - lo for lo, hi in loader] == g["batch_sizes"].tolist() and len(loader) == 7
    with pytest.raises(NotImplementedError):
        tm.ResidentLoader(ds, 7, True)


def test_dataset_normalisation_roundtrip():
    rng = np.random.default_rng(0)
    X, Y = rng.standard_normal((50, 1, 6)) * 3 + 1, rng.standard_normal((50, 1, 2)).astype(np.float32)
    ds = tm.DatasetBase(X, Y, normalize_x=True, normalize_y=True)
    x, y = ds[3]
    assert np.allclose(ds.postprocess_x(x.numpy(), return_tensor=False), X[3], atol=1e-5)
    assert np.allclose(ds.postprocess_y(y.numpy(), return_tensor=False), Y[3], atol=1e-5)
    Xa, Ya = ds.arrays()
    assert abs(Xa.mean()) < 1e-9 and Xa.dtype == np.float64 and Ya.dtype == np.float32


def test_trainer_config_and_grid(tmp_path):
    data = orc.synthetic_episodes(2, 9, 5, 3, seed=1)
    f = str(tmp_path / "demo.pkl")
    pickle.dump(data, open(f, "wb"))
    a = tp.arg_parser().parse_args(["--data_train", f, "--vae_kl_coeff", "0.5", "--max_iter_world_model", "3"])
    cfg = tp.get_trainer_config(a)
    assert cfg["lr_schedule_params"] == {"step_size": 50, "gamma": 0.70} and cfg["batch_size"] == 256 and cfg["lookahead"] == 1
    assert "shuffle_data" not in cfg and cfg["suffle_data"] is True          # the reference's typo is part of the behaviour
    assert cfg["vae_kl_coeff"] == {"grid_search": [1.0, 0.5]}
    mc = cfg["model"]["custom_model_config"]
    assert mc["observation_space"].shape == (10,) and mc["observation_space_body"].shape == (5,) and mc["action_space"].shape == (3,)
    pts = tp.resolve_grid(cfg)
    assert len(pts) == 2 and sorted(p["vae_kl_coeff"] for p in pts) == [0.5, 1.0] and pts[0]["MD_width"] == 512
    tp.update_model_config(pts[0])
    mc = pts[0]["model"]["custom_model_config"]
    assert [l["hidden_size"] for l in mc["world_model_layers"]] == [1024, 1024, "output"]
    assert [l["hidden_size"] for l in mc["motor_decoder_layers"]] == [512, 512, 512, "output"]
    assert mc["task_encoder_output_dim"] == 32 and mc["latent_prior_type"] == "normal_zero_mean_one_std"
    with pytest.raises(AssertionError):
        tp.get_trainer_config(argparse.Namespace(max_iter_world_model=5, max_iter=2, data_train=[f]))
    model = tp.create_model(pts[0])
    assert model.dim_state_body == 5 and model.dim_action == 3 and model.num_outputs == 6


def test_lr_scheduler_factory():
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.Adam(p, lr=5e-4)
    s = tm.get_lr_scheduler(opt, "step", {"step_size": 50, "gamma": 0.7})
    assert isinstance(s, torch.optim.lr_scheduler.StepLR) and s.step_size == 50
    assert tm.get_lr_scheduler(opt, None, None) is None
    assert isinstance(tm.get_lr_scheduler(opt, "cosine", {"T_max": 3}), torch.optim.lr_scheduler.CosineAnnealingLR)


def test_shard_rows_partition_and_weights():
    for lo, hi, world in ((0, 256, 2), (300, 424, 8), (0, 5, 8), (10, 11, 2), (0, 65536, 4)):
        parts = [parallel.shard_rows(lo, hi, r, world) for r in range(world)]
        assert parts[0][0] == lo and parts[-1][1] == hi
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        sizes = [e - s for s, e in parts]
        assert max(sizes) - min(sizes) <= 1 and max(sizes) <= parallel.max_shard_rows(hi - lo, world)
        w = [parallel.shard_weight(lo, hi, r, world) for r in range(world)]
        assert abs(sum(w) / world - 1.0) < 1e-12


def test_sweep_replicas_partition_and_latest_checkpoint(tmp_path):
    """`--sweep_mode replicas`: every grid point runs on exactly one rank; in that mode a trainer sees a world of one."""
    from physicsvae_b200 import parallel
    from physicsvae_b200 import train_physics_vae as tp
    for n, world in [(1, 8), (5, 2), (8, 8), (9, 4)]:
        parts = [parallel.sweep_points(n, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    parallel.set_replica_mode(True)
    try:
        assert parallel.world_size() == 1 and parallel.rank() == 0
    finally:
        parallel.set_replica_mode(False)
    a = tp.arg_parser().parse_args(["--data_train", "x.pkl"])
    assert a.sweep_mode == "dp" and a.resume is False
    assert tp.latest_checkpoint(str(tmp_path / "missing")) is None
    for it in (2, 10):
        d = tmp_path / "trial" / ("checkpoint_%06d" % it)
        d.mkdir(parents=True)
        (d / "model.pth").write_bytes(b"")
    (tmp_path / "trial" / "checkpoint_000011").mkdir()            # incomplete checkpoint directory: ignored
    assert tp.latest_checkpoint(str(tmp_path / "trial")).endswith(os.path.join("checkpoint_000010", "model.pth"))


def test_episode_index_matches_transition_arrays():
    """The compact dataset form of the device-side builder (unique states + per-transition index) describes exactly the X / Y
    rows the reference's loop produces (train_physics_vae.py:133-156), including the --num_data cap and episode boundaries."""
    from physicsvae_b200 import train_physics_vae as tp
    data = orc.synthetic_episodes(4, 17, 7, 3, seed=11)
    data["episodes"][2] = {k: (v[:5] if hasattr(v, "__len__") else v) for k, v in data["episodes"][2].items()}   # a short episode
    for cap in (None, 1, 16, 23, 40, 10 ** 6):
        X, Y = tp.episodes_to_transitions(data["episodes"], num_samples=cap)
        states, actions, first = tp.episodes_to_index(data["episodes"], num_samples=cap)
        assert states.dtype == np.float64 and actions.dtype == np.float32 and first.dtype == np.int64
        assert len(first) == len(X) and states.shape[0] == actions.shape[0] == sum(len(e["time"]) for e in data["episodes"])
        assert np.array_equal(X[:, 0, :7], states[first]) and np.array_equal(X[:, 0, 7:], states[first + 1])
        assert np.array_equal(np.asarray(Y[:, 0, :], dtype=np.float32), actions[first])
    Xg, Yg = tp.episodes_to_transitions(data["episodes"], use_a_gt=True)
    _, ag, fg = tp.episodes_to_index(data["episodes"], use_a_gt=True)
    assert np.array_equal(np.asarray(Yg[:, 0, :], dtype=np.float32), ag[fg])


def test_bench_reference_arm_and_flop_model():
    """`bench.py --impl reference` needs no GPU: it must print one JSON line with the contract's keys; the FLOP model of the roofline
    leg must give SURVEY.md 8d's numbers."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-sample", "256",
                          "--cpu-threads", "1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "transitions/s" and d["higher_is_better"] is True and d["steps"] == 2
    # the reference's own files when they are staged in baseline/_ref (build container, GPU box), else the oracle port
    staged = os.path.isfile(os.path.join(root, "baseline", "_ref", "train_physics_vae.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"] > 0 and d["sample_rows_per_step"] == 256
    out2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "256",
                           "--cpu-kind", "port"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out2.returncode == 0 and json.loads([l for l in out2.stdout.splitlines() if l.startswith("{")][0])["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    import importlib.util
    spec = importlib.util.spec_from_file_location("pvae_bench", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.flops_per_transition(197, 45, 32, [256] * 2, [512] * 3, [1024] * 2) == (8493056, 10269696)
    assert b.flops_per_transition(512, 128, 32, [1024] * 3, [1024] * 3, [1024] * 3) == (18350080, 44892160)
    assert b.flops_per_transition(197, 45, 32, [256] * 2, [512] * 3, [1024] * 2) == orc.flops_per_transition(197, 45, 32, [256] * 2, [512] * 3, [1024] * 2)


def test_round2_host_logic_cli_flags_output_paths_swish_keys():
    """Host-side pieces added in round 2 that need no GPU: the extra CLI flags keep the reference's defaults, a sweep's --output
    names one file per trial, rllib's Swish contributes a `_beta` key per swish layer (state-dict interchange), and FC / PhysicsVAE
    still refuse to compute on the CPU."""
    from physicsvae_b200 import train_physics_vae as tp
    from physicsvae_b200 import rllib_model_torch as pm
    from physicsvae_b200 import parallel, _abi
    a = tp.arg_parser().parse_args(["--data_train", "x.pkl"])
    assert a.deterministic is False and a.sweep_mode == "dp" and a.precision == "bf16x3" and a.batch_size == 256 and a.lr == 0.0005
    assert tp.arg_parser().parse_args(["--data_train", "x.pkl", "--deterministic"]).deterministic is True
    assert tp.output_path("out/model.pt", 3, 1) == "out/model.pt"
    assert tp.output_path("out/model.pt", 3, 4) == "out/model.trial_00003.pt"
    assert parallel.world_size() == 1 and parallel.graph_capturable() and parallel.symmetric_pool_factory() is None
    assert parallel.allreduce_kind()["kind"] == "nccl"
    layers = tp.gen_layers(8, 2, act_hidden="swish")
    fc = pm.FC(size_in=5, size_out=3, layers=layers)
    keys = list(fc.state_dict().keys())
    assert keys == ["_model.0._model.0.weight", "_model.0._model.0.bias", "_model.0._model.1._beta", "_model.1._model.0.weight",
                    "_model.1._model.0.bias", "_model.1._model.1._beta", "_model.2._model.0.weight", "_model.2._model.0.bias"]
    assert float(fc.state_dict()["_model.0._model.1._beta"]) == 1.0 and fc._model[0]._model[1]._beta.requires_grad
    assert fc.layer_spec() == [(8, "swish"), (8, "swish"), (3, "linear")]
    with pytest.raises(_abi.PvaeError):
        fc(torch.zeros(2, 5))                      # CPU tensors: no eager fallback
    with pytest.raises(ValueError):
        pm.get_activation_fn("gelu")


def test_bind_to_gpu_numa_node_reads_sysfs(tmp_path):
    """parallel.bind_to_gpu_numa_node: PCI address -> NUMA node -> CPU list, intersected with the CPUs the process may use."""
    import os
    from physicsvae_b200 import parallel
    bdf = "0000:1b:00.0"
    d = tmp_path / "bus" / "pci" / "devices" / bdf
    d.mkdir(parents=True)
    allowed = sorted(os.sched_getaffinity(0))
    try:
        (d / "numa_node").write_text("1\n")
        n = tmp_path / "devices" / "system" / "node" / "node1"
        n.mkdir(parents=True)
        keep = allowed[: max(1, len(allowed) // 2)]
        (n / "cpulist").write_text("%s,%d-%d\n" % (",".join(map(str, keep)), 100000, 100003))
        r = parallel.bind_to_gpu_numa_node(0, sysfs=str(tmp_path), bdf=bdf)
        assert r["bound"] and r["node"] == 1 and r["cpus"] == len(keep) and r["of"] == len(allowed)
        assert sorted(os.sched_getaffinity(0)) == keep
        (d / "numa_node").write_text("-1\n")
        assert not parallel.bind_to_gpu_numa_node(0, sysfs=str(tmp_path), bdf=bdf)["bound"]
        (d / "numa_node").write_text("1\n")
        (n / "cpulist").write_text("100000-100003\n")
        assert not parallel.bind_to_gpu_numa_node(0, sysfs=str(tmp_path), bdf=bdf)["bound"]       # no usable CPU there: untouched
        assert not parallel.bind_to_gpu_numa_node(0, sysfs=str(tmp_path / "nope"), bdf=bdf)["bound"]
    finally:
        os.sched_setaffinity(0, allowed)
