"""Generate tests/golden/*.npz from the REFERENCE ITSELF (imported unchanged under oracle/ref_stub) in this container.
Run: python oracle/make_golden.py      (needs /root/reference; the fixtures are committed, this script documents them)

Fixtures
  ref_small_step.npz   reduced-width model (dsb 37, da 11, z 8): seeded normc init state dict, a 64-row transition batch,
                       eps, and the reference's forward outputs, world/VAE losses and every parameter gradient.
  ref_loco_ckpt.npz    outputs of the shipped checkpoint data/pretrained/loco_modelV1.pt on a seeded input
                        -- weights are NOT copied (12.5 MB), only sha256 + outputs.
  ref_dataset.npz      load_dataset_for_PhysicsVAE on a synthetic README-format pickle: X / Y arrays and batch sizes.
  ref_small_variants.npz  tiny model; ELU / tanh / sigmoid hidden activations and latent_prior_type=False: init, forward outputs, losses, gradients.
  ref_small_variants2.npz  the same recipe for swish hidden activations (rllib's Swish: x * sigmoid(beta x), beta a parameter initialised
                       to 1) and for the xavier_normal / xavier_uniform initialisers of get_initializer (rllib_model_torch.py:220-232).
  ref_lookahead.npz    compute_loss with lookahead 3 (autoregressive rollout): losses and gradients of both phases.
  ref_trajectory.npz   four epochs of the reference's own TrainModel.step() (two world-model epochs, phase switch, two VAE epochs;
                       Adam lr 5e-4, StepLR, batch 32 with a short last batch, latent_prior_noise False so that no RNG is involved):
                       the pickle, the seeded initial state dict, the per-epoch mean_train_loss and the final state dict.
"""
import hashlib
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload, pvae_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def small_step():
    tpv, tm, rmt = refload.load()
    dsb, da, z = 37, 11, 8
    te, md, wm = tpv.gen_layers(48, 2), tpv.gen_layers(64, 3), tpv.gen_layers(96, 2)
    for l in (te, md, wm):
        l[-1]["init_weight"] = {"name": "normc", "std": 0.3}
    torch.manual_seed(0)
    model = refload.build_reference_model(dsb, da, z, te, md, wm, vf_layers=tpv.gen_layers(48, 2))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B = 64
    data = orc.synthetic_episodes(1, 129, dsb, da, seed=5)
    X, Y = orc.build_transitions(data["episodes"], num_samples=B)
    x = torch.Tensor(X)
    y = torch.Tensor(Y)
    out = {"B": B, "dsb": dsb, "da": da, "z": z, "X": X, "Y": Y}
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    # forward with the noise draw pinned: eps is the first RNG draw of forward()
    torch.manual_seed(7)
    eps = torch.randn(B, z)
    out["eps"] = eps.numpy()
    torch.manual_seed(7)
    logits, state = model(input_dict={"obs": x[:, 0, :], "obs_flat": x[:, 0, :]}, state=None, seq_lens=None)
    out["logits"] = logits.detach().numpy()
    out["mu"] = model._cur_task_encoder_mu.detach().numpy()
    out["logvar"] = model._cur_task_encoder_logvar.detach().numpy()
    out["z_task"] = model._cur_task_encoder_variable.detach().numpy()
    out["future"] = model._cur_future_state.detach().numpy()
    out["value"] = model._cur_value.detach().numpy()
    for world in (True, False):
        model.zero_grad()
        model.set_learnable_task_encoder(not world)
        model.set_learnable_motor_decoder(not world)
        model.set_learnable_world_model(world)
        torch.manual_seed(7)
        loss = refload.reference_compute_loss(model, x, y, world, kl_coeff=1.0, cyc_coeff=0.05)
        loss.backward()
        tag = "world" if world else "vae"
        out[tag + "/loss"] = float(loss)
        for k, p in model.named_parameters():
            if p.grad is not None:
                out[tag + "/grad/" + k] = p.grad.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "ref_small_step.npz"), **out)
    print("ref_small_step.npz: world loss %.8f vae loss %.8f" % (out["world/loss"], out["vae/loss"]))


def small_variants():
    """ref_small_variants.npz: the same recipe as ref_small_step on a tiny model for two more configurations the reference
    supports -- ELU hidden activations (`gen_layers(..., act_hidden="elu")`) and `latent_prior_type=False` (z = encoder output,
    no KL term)."""
    tpv, tm, rmt = refload.load()
    dsb, da, z, B = 13, 5, 4, 48
    data = orc.synthetic_episodes(1, 65, dsb, da, seed=6)
    X, Y = orc.build_transitions(data["episodes"], num_samples=B)
    x, y = torch.Tensor(X), torch.Tensor(Y)
    out = {"B": B, "dsb": dsb, "da": da, "z": z, "X": X, "Y": Y}
    torch.manual_seed(7)
    out["eps"] = torch.randn(B, z).numpy()
    for tag, act, prior in (("elu", "elu", "normal_zero_mean_one_std"), ("noprior", "relu", False),
                            ("tanh", "tanh", "normal_zero_mean_one_std"), ("sigmoid", "sigmoid", "normal_zero_mean_one_std")):
        te, md, wm, vf = (tpv.gen_layers(16, 2, act_hidden=act), tpv.gen_layers(24, 3, act_hidden=act), tpv.gen_layers(32, 2, act_hidden=act),
                          tpv.gen_layers(16, 2, act_hidden=act))
        for l in (te, md, wm):
            l[-1]["init_weight"] = {"name": "normc", "std": 0.3}
        torch.manual_seed(3)
        model = refload.build_reference_model(dsb, da, z, te, md, wm, vf_layers=vf, prior=prior)
        for k, v in model.state_dict().items():
            out[tag + "/sd/" + k] = v.detach().numpy().copy()
        torch.manual_seed(7)
        logits, _ = model(input_dict={"obs": x[:, 0, :], "obs_flat": x[:, 0, :]}, state=None, seq_lens=None)
        out[tag + "/logits"] = logits.detach().numpy()
        out[tag + "/z_task"] = model._cur_task_encoder_variable.detach().numpy()
        out[tag + "/future"] = model._cur_future_state.detach().numpy()
        out[tag + "/value"] = model._cur_value.detach().numpy()
        for world in (True, False):
            model.zero_grad()
            model.set_learnable_task_encoder(not world)
            model.set_learnable_motor_decoder(not world)
            model.set_learnable_world_model(world)
            torch.manual_seed(7)
            loss = refload.reference_compute_loss(model, x, y, world, kl_coeff=1.0, cyc_coeff=0.05, prior=prior)
            loss.backward()
            ph = tag + ("/world" if world else "/vae")
            out[ph + "/loss"] = float(loss)
            for k, p_ in model.named_parameters():
                if p_.grad is not None:
                    out[ph + "/grad/" + k] = p_.grad.detach().numpy().copy()
        print("ref_small_variants.npz %s: world loss %.8f vae loss %.8f" % (tag, out[tag + "/world/loss"], out[tag + "/vae/loss"]))
    np.savez_compressed(os.path.join(OUT, "ref_small_variants.npz"), **out)


def small_variants2():
    """ref_small_variants2.npz: swish hidden layers, and the two xavier initialisers (seeded init + losses + gradients)."""
    tpv, tm, rmt = refload.load()
    dsb, da, z, B = 13, 5, 4, 48
    data = orc.synthetic_episodes(1, 65, dsb, da, seed=6)
    X, Y = orc.build_transitions(data["episodes"], num_samples=B)
    x, y = torch.Tensor(X), torch.Tensor(Y)
    out = {"B": B, "dsb": dsb, "da": da, "z": z, "X": X, "Y": Y}
    torch.manual_seed(7)
    out["eps"] = torch.randn(B, z).numpy()
    prior = "normal_zero_mean_one_std"
    for tag, act, init in (("swish", "swish", None), ("xavier_normal", "relu", {"name": "xavier_normal", "gain": 1.0}),
                           ("xavier_uniform", "relu", {"name": "xavier_uniform", "gain": 1.4})):
        te, md, wm, vf = (tpv.gen_layers(16, 2, act_hidden=act), tpv.gen_layers(24, 3, act_hidden=act), tpv.gen_layers(32, 2, act_hidden=act),
                          tpv.gen_layers(16, 2, act_hidden=act))
        for l in (te, md, wm):
            l[-1]["init_weight"] = {"name": "normc", "std": 0.3}
        if init is not None:
            for l in (te, md, wm, vf):
                for layer in l:
                    layer["init_weight"] = dict(init)
        torch.manual_seed(3)
        model = refload.build_reference_model(dsb, da, z, te, md, wm, vf_layers=vf, prior=prior)
        for k, v in model.state_dict().items():
            out[tag + "/sd/" + k] = v.detach().numpy().copy()
        torch.manual_seed(7)
        logits, _ = model(input_dict={"obs": x[:, 0, :], "obs_flat": x[:, 0, :]}, state=None, seq_lens=None)
        out[tag + "/logits"] = logits.detach().numpy()
        out[tag + "/z_task"] = model._cur_task_encoder_variable.detach().numpy()
        out[tag + "/future"] = model._cur_future_state.detach().numpy()
        out[tag + "/value"] = model._cur_value.detach().numpy()
        for world in (True, False):
            model.zero_grad()
            model.set_learnable_task_encoder(not world)
            model.set_learnable_motor_decoder(not world)
            model.set_learnable_world_model(world)
            torch.manual_seed(7)
            loss = refload.reference_compute_loss(model, x, y, world, kl_coeff=1.0, cyc_coeff=0.05, prior=prior)
            loss.backward()
            ph = tag + ("/world" if world else "/vae")
            out[ph + "/loss"] = float(loss)
            for k, p_ in model.named_parameters():
                if p_.grad is not None:
                    out[ph + "/grad/" + k] = p_.grad.detach().numpy().copy()
        print("ref_small_variants2.npz %s: world loss %.8f vae loss %.8f" % (tag, out[tag + "/world/loss"], out[tag + "/vae/loss"]))
    np.savez_compressed(os.path.join(OUT, "ref_small_variants2.npz"), **out)


def lookahead():
    """ref_lookahead.npz: compute_loss with lookahead 3 (the autoregressive rollout, train_physics_vae.py:367-428) on a tiny model:
    init, X [B, 3, 2*dsb], Y [B, 3, da], the three eps draws, both phases' losses and every gradient."""
    tpv, tm, rmt = refload.load()
    dsb, da, z, B, L = 11, 4, 3, 24, 3
    te, md, wm, vf = tpv.gen_layers(16, 2), tpv.gen_layers(24, 3), tpv.gen_layers(32, 2), tpv.gen_layers(16, 2)
    for l in (te, md, wm):
        l[-1]["init_weight"] = {"name": "normc", "std": 0.3}
    torch.manual_seed(11)
    model = refload.build_reference_model(dsb, da, z, te, md, wm, vf_layers=vf)
    data = orc.synthetic_episodes(2, 30, dsb, da, seed=12)
    X, Y = orc.build_transitions(data["episodes"], num_samples=B, lookahead=L)
    x, y = torch.Tensor(X), torch.Tensor(Y)
    out = {"B": B, "L": L, "dsb": dsb, "da": da, "z": z, "X": X, "Y": Y}
    for k, v in model.state_dict().items():
        out["sd/" + k] = v.detach().numpy().copy()
    torch.manual_seed(5)
    out["eps"] = torch.stack([torch.randn(B, z) for _ in range(L)]).numpy()
    for world in (True, False):
        model.zero_grad()
        model.set_learnable_task_encoder(not world)
        model.set_learnable_motor_decoder(not world)
        model.set_learnable_world_model(world)
        torch.manual_seed(5)
        loss = refload.reference_compute_loss(model, x, y, world, kl_coeff=1.0, cyc_coeff=0.05, lookahead=L)
        loss.backward()
        tag = "world" if world else "vae"
        out[tag + "/loss"] = float(loss)
        for k, p_ in model.named_parameters():
            if p_.grad is not None:
                out[tag + "/grad/" + k] = p_.grad.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "ref_lookahead.npz"), **out)
    print("ref_lookahead.npz: world loss %.8f vae loss %.8f" % (out["world/loss"], out["vae/loss"]))


def loco_ckpt():
    tpv, tm, rmt = refload.load()
    path = os.path.join(refload.REFERENCE, "data", "pretrained", "loco_modelV1.pt")
    sha = hashlib.sha256(open(path, "rb").read()).hexdigest()
    model = refload.build_reference_model(361, 54, 32, tpv.gen_layers(256, 2), tpv.gen_layers(512, 3), tpv.gen_layers(1024, 2))
    model.load_weights(path)
    x = torch.randn(4, 722, generator=torch.Generator().manual_seed(1234))
    model.latent_prior_noise = False
    logits, _ = model(input_dict={"obs": x, "obs_flat": x}, state=None, seq_lens=None)
    np.savez_compressed(os.path.join(OUT, "ref_loco_ckpt.npz"), sha256=sha, x=x.numpy(), logits=logits.detach().numpy(),
                        mu=model._cur_task_encoder_mu.detach().numpy(), logvar=model._cur_task_encoder_logvar.detach().numpy(),
                        future=model._cur_future_state.detach().numpy(), value=model._cur_value.detach().numpy(),
                        keys=np.array(list(model.state_dict().keys())))
    print("ref_loco_ckpt.npz: sha256 %s a[0,:4] %s" % (sha[:12], logits[0, :4].tolist()))


def dataset():
    tpv, tm, rmt = refload.load()
    data = orc.synthetic_episodes(3, 17, 5, 3, seed=9)
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "demo.pkl")
        with open(f, "wb") as fh:
            pickle.dump(data, fh)
        ds = tpv.load_dataset_for_PhysicsVAE([f], num_samples=None, lookahead=1)
        ds_cap = tpv.load_dataset_for_PhysicsVAE([f], num_samples=20, lookahead=1)
        ds_rel = tpv.load_dataset_for_PhysicsVAE([f], num_samples=None, lookahead=2, cond="rel", use_a_gt=True)
        loader = torch.utils.data.DataLoader(ds, batch_size=7, shuffle=None)
        sizes = [int(xb.shape[0]) for xb, yb in loader]
        x0, y0 = next(iter(loader))
    np.savez_compressed(os.path.join(OUT, "ref_dataset.npz"), pickle_bytes=np.frombuffer(pickle.dumps(data), dtype=np.uint8),
                        X=ds.X, Y=ds.Y, X_cap=ds_cap.X, Y_cap=ds_cap.Y, X_rel=ds_rel.X, Y_rel=ds_rel.Y,
                        batch_sizes=np.array(sizes), x0=x0.numpy(), y0=y0.numpy())
    print("ref_dataset.npz: X %s %s Y %s %s batches %s" % (ds.X.shape, ds.X.dtype, ds.Y.shape, ds.Y.dtype, sizes))


def trajectory():
    import argparse
    tpv, tm, rmt = refload.load()
    dsb, da, z = 13, 5, 4
    data = orc.synthetic_episodes(3, 41, dsb, da, seed=2)
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "demo.pkl")
        with open(f, "wb") as fh:
            pickle.dump(data, fh)
        args = argparse.Namespace(max_iter_world_model=2, max_iter=4, data_train=[f], data_test=None, world_model=None, lr=5e-4,
                                  lr_schedule="step", batch_size=32, latent_dim=z, latent_prior_type=["normal_zero_mean_one_std"],
                                  vae_kl_coeff=[1.0], vae_cycle_coeff=[1e-3], num_data=None)
        tpv.args = args
        cfg = tpv.get_trainer_config(args)
        for k, v in list(cfg.items()):
            if isinstance(v, dict) and set(v) == {"grid_search"}:
                cfg[k] = v["grid_search"][0]
        cfg.update(TE_width=16, MD_width=24, world_model_width=32)
        torch.manual_seed(5)
        ref = tpv.TrainModel(cfg)
    ref.model.latent_prior_noise = False
    out = {"pickle_bytes": np.frombuffer(pickle.dumps(data), dtype=np.uint8), "dsb": dsb, "da": da, "z": z, "batch_size": 32, "lr": 5e-4,
           "max_iter_world_model": 2, "TE_width": 16, "MD_width": 24, "world_model_width": 32,
           "lr_step_size": cfg["lr_schedule_params"]["step_size"], "lr_gamma": cfg["lr_schedule_params"]["gamma"]}
    for k, v in ref.model.state_dict().items():
        out["init/" + k] = v.detach().numpy().copy()
    losses = []
    for it in range(4):
        r = ref.step()
        losses.append(r["mean_train_loss"])
    out["losses"] = np.array(losses, dtype=np.float64)
    for k, v in ref.model.state_dict().items():
        if k.startswith("_value_branch"):                  # never trained (no loss term): identical to init, checked here, not stored
            assert np.array_equal(v.detach().numpy(), out["init/" + k]), k
            continue
        out["final/" + k] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "ref_trajectory.npz"), **out)
    print("ref_trajectory.npz: losses %s" % losses)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    small_step()
    loco_ckpt()
    dataset()
    trajectory()
    small_variants()
    small_variants2()
    lookahead()
