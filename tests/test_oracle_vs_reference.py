"""Live check of the oracle against the reference's own files imported unchanged under oracle/ref_stub.  Runs only where
/root/reference is mounted (the build container); skipped on the GPU box.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import pvae_oracle as orc
from oracle import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference is not mounted here")


def test_default_dims_step_matches_live_reference():
    tpv, tm, rmt = refload.load()
    dsb, da, z, B = 197, 45, 32, 32
    torch.manual_seed(3)
    ref = refload.build_reference_model(dsb, da, z, tpv.gen_layers(256, 2), tpv.gen_layers(512, 3), tpv.gen_layers(1024, 2))
    torch.manual_seed(3)
    m = orc.OracleModel(dsb, da, z)
    for k, v in ref.state_dict().items():
        assert torch.equal(m.params[k], v), k
    X, Y = orc.build_transitions(orc.synthetic_episodes(1, 129, dsb, da, seed=1)["episodes"], num_samples=B)
    x, y = torch.Tensor(X), torch.Tensor(Y)
    for world in (True, False):
        ref.zero_grad()
        ref.set_learnable_task_encoder(not world)
        ref.set_learnable_motor_decoder(not world)
        ref.set_learnable_world_model(world)
        torch.manual_seed(11)
        eps = torch.randn(B, z)
        torch.manual_seed(11)
        loss = refload.reference_compute_loss(ref, x, y, world)
        loss.backward()
        o_loss, _, grads = orc.loss_and_grads(m, x[:, 0, :], y[:, 0, :], world, eps=eps)
        assert abs(o_loss - float(loss)) <= 1e-6 * abs(float(loss))
        ref_grads = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
        assert set(ref_grads) == set(grads)
        for k in grads:
            np.testing.assert_allclose(grads[k].numpy(), ref_grads[k].numpy(), rtol=1e-5, atol=1e-10, err_msg=k)


def test_shipped_checkpoint_golden_vector():
    # strict load of data/pretrained/loco_modelV1.pt into the oracle + the committed outputs
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_loco_ckpt.npz"))
    sd = torch.load(os.path.join(refload.REFERENCE, "data", "pretrained", "loco_modelV1.pt"))
    m = orc.OracleModel(361, 54, 32)
    m.load_state_dict(sd)
    assert list(sd.keys()) == g["keys"].tolist()
    m.latent_prior_noise = False
    logits = m.forward(torch.from_numpy(g["x"]))
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(m.cur["future"].numpy(), g["future"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(m.cur["value"].numpy(), g["value"], rtol=1e-5, atol=1e-6)
    assert abs(float(logits[0, 0]) - (-0.24995048)) < 1e-6 and abs(float(logits[0, -1]) - np.log(0.1)) < 1e-6


def test_trainer_trajectory_matches_live_reference():
    """A few epochs of the reference's own TrainModel.step (Adam + StepLR + phase switch) vs OracleTrainer."""
    import pickle
    import tempfile
    tpv, tm, rmt = refload.load()
    dsb, da = 13, 5
    data = orc.synthetic_episodes(2, 41, dsb, da, seed=2)
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "demo.pkl")
        pickle.dump(data, open(f, "wb"))
        import argparse
        args = argparse.Namespace(max_iter_world_model=2, max_iter=4, data_train=[f], data_test=None, world_model=None, lr=5e-4,
                                  lr_schedule="step", batch_size=32, latent_dim=4, latent_prior_type=["normal_zero_mean_one_std"],
                                  vae_kl_coeff=[1.0], vae_cycle_coeff=[1e-3], num_data=None)
        tpv.args = args
        cfg = tpv.get_trainer_config(args)
        for k, v in list(cfg.items()):
            if isinstance(v, dict) and set(v) == {"grid_search"}:
                cfg[k] = v["grid_search"][0]
        cfg.update(TE_width=16, MD_width=24, world_model_width=32)
        torch.manual_seed(5)
        ref = tpv.TrainModel(cfg)
    torch.manual_seed(5)
    m = orc.OracleModel(dsb, da, 4, orc.gen_layers(16, 2), orc.gen_layers(24, 3), orc.gen_layers(32, 2))
    for k, v in ref.model.state_dict().items():
        assert torch.equal(m.params[k], v), k
    X, Y = orc.build_transitions(data["episodes"])
    tr = orc.OracleTrainer(m, X, Y, batch_size=32, lr=5e-4, max_iter_world_model=2)
    for it in range(4):
        # the reference draws eps with torch.randn_like inside forward: replay the same global-RNG stream on both sides.
        # In the world phase the reference's discarded full forward (train_physics_vae.py:377-378) also consumes one draw.
        torch.manual_seed(100 + it)
        r = ref.step()
        torch.manual_seed(100 + it)
        torch.empty((), dtype=torch.int64).random_()      # the DataLoader iterator draws its base seed from the global RNG
        def eps_fn(i, b, n):
            return torch.randn(n, 4)
        o = tr.step(eps_fn=eps_fn)
        assert abs(o["mean_train_loss"] - r["mean_train_loss"]) <= 2e-6 * abs(r["mean_train_loss"]), (it, o, r)
    for k, v in ref.model.state_dict().items():
        np.testing.assert_allclose(m.params[k].numpy(), v.numpy(), rtol=2e-5, atol=2e-7, err_msg=k)


def test_lookahead_rollout_matches_live_reference():
    """lookahead 3: the autoregressive rollout of compute_loss (train_physics_vae.py:367-428) -- both phases, losses and every
    gradient (gradients flow through the predicted states), against the reference's own code."""
    tpv, tm, rmt = refload.load()
    dsb, da, z, B, L = 11, 4, 3, 24, 3
    te, md, wm, vf = tpv.gen_layers(16, 2), tpv.gen_layers(24, 3), tpv.gen_layers(32, 2), tpv.gen_layers(16, 2)
    for l in (te, md, wm):
        l[-1]["init_weight"] = {"name": "normc", "std": 0.3}
    torch.manual_seed(11)
    ref = refload.build_reference_model(dsb, da, z, te, md, wm, vf_layers=vf)
    m = orc.OracleModel(dsb, da, z, te, md, wm, vf)
    m.load_state_dict({k: v.detach().clone() for k, v in ref.state_dict().items()})
    data = orc.synthetic_episodes(2, 30, dsb, da, seed=12)
    X, Y = orc.build_transitions(data["episodes"], num_samples=B, lookahead=L)
    assert X.shape == (B, L, 2 * dsb) and Y.shape == (B, L, da)
    x, y = torch.Tensor(X), torch.Tensor(Y)
    for world in (True, False):
        ref.zero_grad()
        ref.set_learnable_task_encoder(not world); ref.set_learnable_motor_decoder(not world); ref.set_learnable_world_model(world)
        torch.manual_seed(5)
        loss = refload.reference_compute_loss(ref, x, y, world, kl_coeff=1.0, cyc_coeff=0.05, lookahead=L)
        loss.backward()
        torch.manual_seed(5)
        eps = torch.stack([torch.randn(B, z) for _ in range(L)])           # one randn_like draw per forward, in step order
        o_loss, parts, grads = orc.loss_and_grads(m, x, y, world, kl_coeff=1.0, cyc_coeff=0.05, eps=eps)
        assert abs(o_loss - float(loss)) <= 1e-6 * abs(float(loss)), (world, o_loss, float(loss))
        want = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
        assert set(grads) == set(want)
        for k in want:
            np.testing.assert_allclose(grads[k].numpy(), want[k].numpy(), rtol=2e-5, atol=1e-9, err_msg=k)
    # lookahead 1 through the general form == the specialised compute_loss
    x1, y1 = x[:, :1, :], y[:, :1, :]
    a = orc.loss_and_grads(m, x1, y1, False, cyc_coeff=0.05, eps=eps[:1])
    b = orc.loss_and_grads(m, x1[:, 0, :], y1[:, 0, :], False, cyc_coeff=0.05, eps=eps[0])
    assert abs(a[0] - b[0]) <= 1e-7 * abs(b[0]) and all(torch.allclose(a[2][k], b[2][k], rtol=1e-6, atol=1e-9) for k in b[2])


def test_product_checkpoint_loads_strictly_into_the_live_reference_class(tmp_path):
    """Interchange, product -> reference: `model.pt` / `world_model.pt` / `motor_decoder.pt` / `task_encoder.pt` written by the
    product's save_weights* (train_physics_vae.py:440-467's file set) load with strict=True into the reference's own PhysicsVAE
    through the reference's own loaders (rllib_model_torch.py:870-928), tensor for tensor."""
    from physicsvae_b200 import rllib_model_torch as pm
    from physicsvae_b200 import train_physics_vae as tp
    tpv, tm, rmt = refload.load()
    dsb, da, z = 23, 7, 6
    box = lambda n: tp.Box(low=-np.ones(n), high=np.ones(n), dtype=np.float64)
    custom = dict(pm.PhysicsVAE.DEFAULT_CONFIG)
    custom.update(observation_space=box(2 * dsb), observation_space_body=box(dsb), observation_space_task=box(dsb), action_space=box(da),
                  task_encoder_output_dim=z, task_encoder_layers=tp.gen_layers(16, 2), motor_decoder_layers=tp.gen_layers(24, 3),
                  world_model_layers=tp.gen_layers(32, 2), value_fn_layers=tp.gen_layers(16, 2))
    torch.manual_seed(21)
    ours = pm.PhysicsVAE(box(2 * dsb), box(da), 2 * da, {"custom_model_config": custom}, "physics_vae")
    files = {n: str(tmp_path / (n + ".pt")) for n in ("model", "task_encoder", "motor_decoder", "world_model")}
    ours.save_weights(files["model"])
    ours.save_weights_task_encoder(files["task_encoder"])
    ours.save_weights_motor_decoder(files["motor_decoder"])
    ours.save_weights_world_model(files["world_model"])
    torch.manual_seed(99)                                   # a differently initialised reference model
    ref = refload.build_reference_model(dsb, da, z, tpv.gen_layers(16, 2), tpv.gen_layers(24, 3), tpv.gen_layers(32, 2),
                                        vf_layers=tpv.gen_layers(16, 2))
    assert any(not torch.equal(v, ours.state_dict()[k]) for k, v in ref.state_dict().items())
    missing = ref.load_state_dict(torch.load(files["model"]), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for k, v in ref.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k
    # the per-part files through the reference's own loaders
    torch.manual_seed(100)
    ref2 = refload.build_reference_model(dsb, da, z, tpv.gen_layers(16, 2), tpv.gen_layers(24, 3), tpv.gen_layers(32, 2),
                                         vf_layers=tpv.gen_layers(16, 2))
    ref2.load_weights_task_encoder(files["task_encoder"])
    ref2.load_weights_motor_decoder(files["motor_decoder"])
    ref2.load_weights_world_model(files["world_model"])
    for k, v in ref2.state_dict().items():
        if not k.startswith("_value_branch"):
            assert torch.equal(v, ours.state_dict()[k]), k
    # and the other way round: the reference's file into the product
    ref2.save_weights(str(tmp_path / "ref_model.pt"))
    ours2 = pm.PhysicsVAE(box(2 * dsb), box(da), 2 * da, {"custom_model_config": custom}, "physics_vae")
    ours2.load_weights(str(tmp_path / "ref_model.pt"))
    for k, v in ours2.state_dict().items():
        assert torch.equal(v, ref2.state_dict()[k]), k
