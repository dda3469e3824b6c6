"""Shared parity harness: builds the CPU oracle and the sm_100a product on identical weights / transitions and compares
them.  Used by tests/ (-m gpu), __graft_entry__.smoke() and bench.py.  The oracle is only ever the checker here."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pvae_oracle as orc  # noqa: E402

# north_star tolerance for fp32 parity: outputs within rtol=1e-3 / atol=1e-5 of the reference's PyTorch-CPU results
RTOL, ATOL = 1e-3, 1e-5

SMALL = dict(dsb=37, da=11, z=8, te=(48, 2), md=(64, 3), wm=(96, 2))
DEFAULT = dict(dsb=197, da=45, z=32, te=(256, 2), md=(512, 3), wm=(1024, 2))
LOCO = dict(dsb=361, da=54, z=32, te=(256, 2), md=(512, 3), wm=(1024, 2))
WIDE = dict(dsb=512, da=128, z=32, te=(1024, 3), md=(1024, 3), wm=(1024, 3))


def oracle_model(cfg, seed=0, prior="normal_zero_mean_one_std", act="relu", out_std=None):
    torch.manual_seed(seed)
    layers = {k: orc.gen_layers(cfg[k][0], cfg[k][1], act_hidden=act) for k in ("te", "md", "wm")}
    if out_std is not None:      # a larger output-layer init makes the loss landscape less degenerate for gradient checks
        for l in layers.values():
            l[-1]["init_weight"] = {"name": "normc", "std": out_std}
    m = orc.OracleModel(cfg["dsb"], cfg["da"], cfg["z"], layers["te"], layers["md"], layers["wm"],
                        orc.gen_layers(cfg["te"][0], cfg["te"][1], act_hidden=act), latent_prior_type=prior)
    return m, layers


def product_model(cfg, layers, state_dict, prior="normal_zero_mean_one_std", precision="bf16x3", max_batch=1024, act="relu"):
    """The product's PhysicsVAE on cuda:0 carrying the oracle's weights."""
    from physicsvae_b200 import rllib_model_torch as pm
    from physicsvae_b200 import train_physics_vae as tp
    dsb, da = cfg["dsb"], cfg["da"]
    box = lambda n: tp.Box(low=-np.ones(n), high=np.ones(n), dtype=np.float64)
    custom = dict(pm.PhysicsVAE.DEFAULT_CONFIG)
    custom.update(observation_space=box(2 * dsb), observation_space_body=box(dsb), observation_space_task=box(dsb),
                  action_space=box(da), task_encoder_output_dim=cfg["z"], task_encoder_layers=layers["te"],
                  motor_decoder_layers=layers["md"], world_model_layers=layers["wm"],
                  value_fn_layers=orc.gen_layers(cfg["te"][0], cfg["te"][1], act_hidden=act),
                  latent_prior_type=prior, engine_precision=precision, engine_max_batch=max_batch)
    model = pm.PhysicsVAE(box(2 * dsb), box(da), 2 * da, {"custom_model_config": custom}, "physics_vae")
    sd = dict(state_dict)
    for k, v in model.state_dict().items():          # rllib's Swish carries a `_beta` parameter the oracle does not model (beta = 1)
        if k.endswith("._beta") and k not in sd:
            sd[k] = v.detach().clone()
    model.load_state_dict(sd)
    return model.to("cuda:0")


def fast_transitions(cfg, n, seed=0):
    """Vectorised synthetic transitions for the full-size tests (same recipe as oracle.synthetic_episodes):
    X float64 [n, 2*dsb], Y float32 [n, da]."""
    rng = np.random.default_rng(seed)
    T = 129
    E = (n + T - 2) // (T - 1)
    dsb = cfg["dsb"]
    s = rng.standard_normal((E, 1, dsb)) + np.cumsum(
        np.concatenate([np.zeros((E, 1, dsb)), 0.05 * rng.standard_normal((E, T - 1, dsb))], axis=1), axis=1)
    X = np.concatenate([s[:, :-1], s[:, 1:]], axis=-1).reshape(-1, 2 * dsb)[:n]
    Y = rng.uniform(-1, 1, size=(n, cfg["da"])).astype(np.float32)
    return np.ascontiguousarray(X), Y


def make_trainer(cfg, X, Y, batch_size, precision="bf16x3", world_epochs=10 ** 9, state_dict=None, X_test=None, Y_test=None, act="relu", **extra):
    """The product's train_physics_vae.TrainModel on given transition arrays (no pickle): DatasetBase(X, Y) straight in."""
    from physicsvae_b200 import train_physics_vae as tp
    from physicsvae_b200 import torch_models as tm
    sets = {"train": (X, Y), "test": (X_test, Y_test)}

    class ArrayTrainer(tp.TrainModel):
        def load_dataset(self, file):
            Xa, Ya = sets[file]
            return tm.DatasetBase(np.asarray(Xa).reshape(len(Xa), 1, -1), np.asarray(Ya).reshape(len(Ya), 1, -1), normalize_x=False, normalize_y=False)

    box = lambda n: tp.Box(low=-np.ones(n), high=np.ones(n), dtype=np.float64)
    custom = dict(tp.MODEL_CONFIG)
    custom.update(observation_space=box(2 * cfg["dsb"]), observation_space_body=box(cfg["dsb"]), observation_space_task=box(cfg["dsb"]),
                  action_space=box(cfg["da"]), engine_precision=precision, engine_max_batch=batch_size,
                  value_fn_layers=orc.gen_layers(cfg["te"][0], cfg["te"][1], act_hidden=act))          # (oracle_model builds the value branch like this)
    config = {"max_iter_world_model": world_epochs, "model": {"custom_model": "physics_vae", "custom_model_config": custom},
              "lr": 5e-4, "lr_schedule": "step", "lr_schedule_params": {"step_size": 50, "gamma": 0.7}, "weight_decay": 0.0,
              "dataset_train": "train", "dataset_test": "test" if X_test is not None else None, "loss": "MSE", "loss_test": "MSE",
              "batch_size": batch_size, "latent_dim": cfg["z"], "latent_prior_type": "normal_zero_mean_one_std", "act_fn": act,
              "MD_width": cfg["md"][0], "MD_depth": cfg["md"][1], "TE_width": cfg["te"][0], "TE_depth": cfg["te"][1],
              "lookahead": 1, "world_model_width": cfg["wm"][0], "world_model_depth": cfg["wm"][1], "vae_kl_coeff": 1.0,
              "motor_decoder_a_rec_coeff": 1.0, "world_model_s_rec_coeff": 0.0, "vae_cycle_coeff": 1e-3,
              "engine_precision": precision}
    config.update(extra)
    tr = ArrayTrainer(config)
    if state_dict is not None:
        sd = dict(state_dict)
        for k, v in tr.model.state_dict().items():
            if k.endswith("._beta") and k not in sd:
                sd[k] = v.detach().clone()
        tr.model.load_state_dict(sd)
    return tr


def transitions(cfg, n, seed=0):
    """Synthetic transition batch in the dataset's dtypes: X float64 [n, 2*dsb] = (s_t | s_{t+1}), Y float32 [n, da]."""
    T = 129
    eps = orc.synthetic_episodes((n + T - 2) // (T - 1), T, cfg["dsb"], cfg["da"], seed=seed)["episodes"]
    X, Y = orc.build_transitions(eps, num_samples=n)
    return X.reshape(n, -1), Y.reshape(n, -1).astype(np.float32)


def grads_by_net(model):
    """{net name: {state-dict key: grad}} from the product module's .grad views."""
    out = {}
    for k, p in model.named_parameters():
        if p.grad is not None:
            out[k] = p.grad.detach().cpu().clone()
    return out


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def assert_close(name, got, ref, rtol=RTOL, atol=ATOL):
    got, ref = got.detach().cpu().float(), ref.detach().cpu().float()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    bad = (got - ref).abs() > atol + rtol * ref.abs()
    assert not bool(bad.any()), "%s: %d / %d elements outside rtol=%g atol=%g (max abs err %.3e, rel L2 %.3e)" % (
        name, int(bad.sum()), bad.numel(), rtol, atol, float((got - ref).abs().max()), rel_l2(got, ref))


def assert_close_scaled(name, got, ref, rtol=RTOL, atol=ATOL):
    """assert_close with the absolute term scaled to the tensor's magnitude: |err| <= atol * max(1, max|ref|) + rtol * |ref|.
    For outputs far from unit scale (xavier-initialised nets, the shipped checkpoint on N(0, 1) inputs): the fp32-accurate mode
    carries 16 mantissa bits per operand (hi + lo bf16), i.e. errors of ~1e-5 of the tensor's scale, which an element that
    happens to be near zero cannot meet relative to itself."""
    got, ref = got.detach().cpu().float(), ref.detach().cpu().float()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = max(1.0, float(ref.abs().max()))
    bad = (got - ref).abs() > atol * scale + rtol * ref.abs()
    assert not bool(bad.any()), "%s: %d / %d elements outside rtol=%g atol=%g x %.3g (max abs err %.3e, rel L2 %.3e)" % (
        name, int(bad.sum()), bad.numel(), rtol, atol, scale, float((got - ref).abs().max()), rel_l2(got, ref))
    assert rel_l2(got, ref) < 1e-4, (name, rel_l2(got, ref))


def step_pair(cfg, B, world, prior="normal_zero_mean_one_std", precision="bf16x3", act="relu", seed=0, out_std=None,
              kl_coeff=1.0, cyc_coeff=1e-3, n_rows=None, cursor=0):
    """Run one training step (forward + loss + backward) on both sides.  Returns dicts (oracle, product)."""
    om, layers = oracle_model(cfg, seed=seed, prior=prior, act=act, out_std=out_std)
    n_rows = n_rows or B
    X, Y = transitions(cfg, n_rows, seed=seed + 1)
    x = torch.from_numpy(X[cursor:cursor + B]).float()
    y = torch.from_numpy(Y[cursor:cursor + B])
    eps = torch.randn(B, cfg["z"], generator=torch.Generator().manual_seed(seed + 2))
    o_loss, o_parts, o_grads = orc.loss_and_grads(om, x, y, world, kl_coeff=kl_coeff, cyc_coeff=cyc_coeff, eps=eps)
    pm_ = product_model(cfg, layers, om.state_dict(), prior=prior, precision=precision, max_batch=max(B, 128), act=act)
    eng = pm_.engine()
    eng.alloc_transitions(n_rows)
    eng.ingest(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
    pm_.sync_weights()
    pm_.set_learnable_task_encoder(not world)
    pm_.set_learnable_motor_decoder(not world)
    pm_.set_learnable_world_model(world)
    eng.set_cursor(cursor)
    if world:
        loss = eng.world_step(B)
    else:
        loss = eng.vae_step(B, eps=eps.cuda(), kl_coeff=kl_coeff if prior else 0.0, cyc_coeff=cyc_coeff)
    torch.cuda.synchronize()
    p = {"loss": float(loss[0]), "parts": {"a": float(loss[1]), "kl": float(loss[2]), "s": float(loss[3]), "cyc": float(loss[4])},
         "grads": grads_by_net(pm_), "model": pm_}
    o = {"loss": o_loss, "parts": o_parts, "grads": o_grads, "model": om, "x": x, "y": y, "eps": eps}
    return o, p


def check_step(o, p, B, grad_rel_l2=5e-3):
    """Loss within the north_star tolerance; every gradient tensor within rtol/atol, and -- because a mean-reduced loss
    makes raw gradients ~1e-6 and atol=1e-5 nearly vacuous -- also by relative L2 norm.  The L2 bound
    leaves room for a ReLU mask flipping at |pre-activation| ~ 1e-6, which moves dW by ~1e-3 relative."""
    assert abs(p["loss"] - o["loss"]) <= ATOL + RTOL * abs(o["loss"]), (p["loss"], o["loss"])
    for k in ("a", "kl", "s", "cyc"):
        assert abs(p["parts"][k] - o["parts"][k]) <= ATOL + RTOL * abs(o["parts"][k]), (k, p["parts"], o["parts"])
    assert set(p["grads"].keys()) == set(o["grads"].keys()), sorted(set(p["grads"]) ^ set(o["grads"]))
    worst = 0.0
    for k, g in o["grads"].items():
        assert_close("grad " + k, p["grads"][k], g)
        r = rel_l2(p["grads"][k], g)
        worst = max(worst, r)
        assert r < grad_rel_l2, "grad %s: relative L2 error %.3e" % (k, r)
    return worst


def run_smoke():
    """__graft_entry__.smoke(): one small world step + one VAE step on cuda:0, checked against the oracle."""
    from physicsvae_b200 import _abi
    n0 = _abi.launch_count() if os.path.exists(_abi.LIB_PATH) else 0
    for world in (True, False):
        o, p = step_pair(SMALL, 200, world, out_std=0.3, cyc_coeff=0.05, n_rows=512, cursor=100)
        worst = check_step(o, p, 200)
        print("[smoke] %s step: loss %.6f (oracle %.6f), worst grad rel-L2 %.2e" % ("world" if world else "vae", p["loss"], o["loss"], worst))
    print("[smoke] ok, %d kernel launches from libpvae_sm100.so" % (_abi.launch_count() - n0))
