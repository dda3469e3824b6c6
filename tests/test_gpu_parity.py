"""Parity tests proper (-m gpu): the sm_100a path, called through the C ABI, against the CPU oracle on the same seeded
inputs, against the reference-generated golden fixtures, and -- at BASELINE.json's full sizes -- through size-independent
properties.  Tolerance: north_star's rtol=1e-3 / atol=1e-5 for the fp32-accurate mode (bf16x3)."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from tests import parity as P
from oracle import pvae_oracle as orc

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _planes(t, planes):
    hi = t.to(torch.bfloat16)
    if planes == 1:
        return hi.contiguous(), hi.double()
    lo = (t - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo]).contiguous(), hi.double() + lo.double()


# ---- kernel level: D = A . B^T in every operand-major combination ---------------------------------------------------------
@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K,planes,splits", [(128, 256, 64, 1, 1), (328, 200, 248, 1, 1), (520, 456, 392, 2, 1),
                                                  (8, 16, 8, 2, 1), (1024, 512, 4104, 1, 8), (264, 1000, 2008, 2, 5)])
def test_tcgen05_gemm(a_major, b_major, M, N, K, planes, splits):
    from physicsvae_b200.engine import gemm_bf16
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    Ap, Ar = _planes(A if a_major == 0 else A.t().contiguous(), planes)
    Bp, Br = _planes(B if b_major == 0 else B.t().contiguous(), planes)
    Ar = Ar.t() if a_major else Ar
    Br = Br.t() if b_major else Br
    ref = (Ar @ Br.t()).float()
    D = gemm_bf16(Ap, Bp, M, N, K, a_major, b_major, planes, splits)
    torch.cuda.synchronize()
    # bf16x1: exact products, fp32 accumulation; bf16x3: drops lo*lo terms (~2^-16 relative per product)
    tol = 2e-5 * K ** 0.5 * (1 if planes == 1 else 4)
    assert float((D - ref).abs().max()) <= tol + 1e-5 * float(ref.abs().max()), float((D - ref).abs().max())


def test_gemm_rejects_misaligned_operands():
    from physicsvae_b200.engine import gemm_bf16
    A = torch.zeros(16, 12, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        gemm_bf16(A, A, 16, 16, 12)          # row stride 12 elements is not a multiple of 16 bytes


# ---- training steps vs the oracle ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [True, False])
@pytest.mark.parametrize("B", [2, 200, 255])
def test_small_step_matches_oracle(world, B):
    o, p = P.step_pair(P.SMALL, B, world, out_std=0.3, cyc_coeff=0.05, n_rows=max(B, 300), cursor=min(45, max(B, 300) - B))
    P.check_step(o, p, B)


@pytest.mark.parametrize("world", [True, False])
def test_default_dims_step_matches_oracle(world):
    # BASELINE.json configs[0] dims and batch: dim_state_body 197, dim_action 45, latent 32, batch 256, normc(0.01) outputs
    o, p = P.step_pair(P.DEFAULT, 256, world, n_rows=1024, cursor=256)
    P.check_step(o, p, 256)


@pytest.mark.parametrize("world", [True, False])
def test_loco_dims_step_matches_oracle(world):
    # the shipped data's dims (361 / 54): K and N that are not multiples of 8
    o, p = P.step_pair(P.LOCO, 130, world, out_std=0.1, cyc_coeff=0.01, n_rows=400, cursor=270)
    P.check_step(o, p, 130)


@pytest.mark.parametrize("world", [True, False])
def test_wide_dims_step_matches_oracle(world):
    # BASELINE.json configs[4]: dim_state 512, dim_action 128, hidden [1024, 1024, 1024] for all three trained MLPs
    o, p = P.step_pair(P.WIDE, 160, world, out_std=0.1, cyc_coeff=0.01, n_rows=200, cursor=40)
    P.check_step(o, p, 160)


@pytest.mark.parametrize("world", [True, False])
def test_smooth_activation_gradients_are_tight(world):
    # ELU has a continuous derivative, so there are no ReLU-mask flips and the backward machinery can be held to 1e-4
    o, p = P.step_pair(P.SMALL, 200, world, act="elu", out_std=0.3, cyc_coeff=0.05)
    assert P.check_step(o, p, 200, grad_rel_l2=1e-4) < 1e-4


@pytest.mark.parametrize("act", ["tanh", "sigmoid"])
def test_other_activations(act):
    o, p = P.step_pair(P.SMALL, 96, False, act=act, out_std=0.3, cyc_coeff=0.05)
    P.check_step(o, p, 96, grad_rel_l2=1e-4)


def test_no_prior_vae_step():
    o, p = P.step_pair(P.SMALL, 150, False, prior=False, out_std=0.3, cyc_coeff=0.05)
    assert o["parts"]["kl"] == 0.0
    P.check_step(o, p, 150)


def test_zero_cycle_coefficient_skips_world_model():
    o, p = P.step_pair(P.SMALL, 150, False, out_std=0.3, cyc_coeff=0.0)
    P.check_step(o, p, 150)


def test_bf16_mode_is_close():
    # performance mode: bf16 operands, fp32 accumulation.  Not held to the fp32 tolerance; report-level closeness only.
    o, p = P.step_pair(P.DEFAULT, 512, True, precision="bf16", n_rows=512)
    assert abs(p["loss"] - o["loss"]) < 2e-3 * abs(o["loss"])
    for k, g in o["grads"].items():
        assert P.rel_l2(p["grads"][k], g) < 3e-2, k


@pytest.mark.parametrize("world", [True, False])
@pytest.mark.parametrize("cfg_name,B", [("DEFAULT", 300), ("LOCO", 130), ("SMALL", 77)])
def test_tma_epilogue_equals_direct_epilogue(world, cfg_name, B, monkeypatch):
    """bf16 mode, the two epilogue implementations of the GEMM kernel (shared-memory staged TMA stores / aux loads vs direct
    row accesses) must agree: same loss, same activations and gradients; only the bias gradients differ, by the bf16
    rounding of the summed values."""
    cfg = getattr(P, cfg_name)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PVAE_TMA_EPILOGUE", mode)
        o, p = P.step_pair(cfg, B, world, precision="bf16", out_std=0.2, cyc_coeff=0.05, n_rows=B + 70, cursor=33)
        res[mode] = p
    a, b = res["1"], res["0"]
    assert abs(a["loss"] - b["loss"]) <= 1e-6 * abs(b["loss"])
    for k in b["grads"]:
        tol = 5e-3 if k.endswith("bias") else 2e-5
        assert P.rel_l2(a["grads"][k], b["grads"][k]) < tol, (k, P.rel_l2(a["grads"][k], b["grads"][k]))
    # and both are bf16-close to the fp32 oracle (report-level bound; the exact check is the comparison above)
    assert abs(a["loss"] - o["loss"]) < 3e-3 * abs(o["loss"])
    for k, g in o["grads"].items():
        assert P.rel_l2(a["grads"][k], g) < 0.15, (k, P.rel_l2(a["grads"][k], g), P.rel_l2(b["grads"][k], g))


@pytest.mark.parametrize("world", [True, False])
@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_cta_pair_multicast_equals_single_cta(world, precision, monkeypatch):
    """CTA pairs that share the B tile through TMA multicast must reproduce the single-CTA kernel (same products, same
    accumulation order inside a tile; only the order of the fp32 split-K atomics differs)."""
    res = {}
    for mode in ("2", "1"):
        monkeypatch.setenv("PVAE_CLUSTER", mode)
        o, p = P.step_pair(P.DEFAULT, 700, world, precision=precision, out_std=0.2, cyc_coeff=0.05, n_rows=800, cursor=50)
        res[mode] = p
    a, b = res["2"], res["1"]
    assert abs(a["loss"] - b["loss"]) <= 1e-6 * abs(b["loss"])
    for k in b["grads"]:
        assert P.rel_l2(a["grads"][k], b["grads"][k]) < 2e-5, (k, P.rel_l2(a["grads"][k], b["grads"][k]))


# ---- golden fixtures generated by the reference itself ---------------------------------------------------------------------
def test_reference_fixture_forward_losses_gradients():
    g = np.load(os.path.join(G, "ref_small_step.npz"))
    cfg = dict(dsb=int(g["dsb"]), da=int(g["da"]), z=int(g["z"]), te=(48, 2), md=(64, 3), wm=(96, 2))
    layers = {k: orc.gen_layers(*cfg[k]) for k in ("te", "md", "wm")}
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    m = P.product_model(cfg, layers, sd, max_batch=128)
    x = torch.Tensor(g["X"])[:, 0, :].cuda()
    eps = torch.from_numpy(g["eps"]).cuda()
    logits, state = m(input_dict={"obs": x, "obs_flat": x, "eps": eps}, state=None, seq_lens=None)
    assert state == [] and logits.shape == g["logits"].shape
    P.assert_close("logits", logits, torch.from_numpy(g["logits"]))
    P.assert_close("mu", m._cur_task_encoder_mu, torch.from_numpy(g["mu"]))
    P.assert_close("logvar", m._cur_task_encoder_logvar, torch.from_numpy(g["logvar"]))
    P.assert_close("z", m._cur_task_encoder_variable, torch.from_numpy(g["z_task"]))
    P.assert_close("future", m._cur_future_state, torch.from_numpy(g["future"]))
    P.assert_close("value", m.value_function(), torch.from_numpy(g["value"]))
    eng = m.engine()
    B = int(g["B"])
    eng.alloc_transitions(B)
    eng.ingest(torch.from_numpy(g["X"]).cuda(), torch.from_numpy(g["Y"]).cuda())
    for world in (True, False):
        tag = "world" if world else "vae"
        m.set_learnable_task_encoder(not world); m.set_learnable_motor_decoder(not world); m.set_learnable_world_model(world)
        eng.set_cursor(0)
        loss = eng.world_step(B) if world else eng.vae_step(B, eps=eps, kl_coeff=1.0, cyc_coeff=0.05)
        torch.cuda.synchronize()
        ref_loss = float(g[tag + "/loss"])
        assert abs(float(loss[0]) - ref_loss) <= P.ATOL + P.RTOL * abs(ref_loss)
        ref = {k[len(tag) + 6:]: g[k] for k in g.files if k.startswith(tag + "/grad/")}
        got = P.grads_by_net(m)
        assert set(got) == set(ref)
        for k in ref:
            P.assert_close(tag + " grad " + k, got[k], torch.from_numpy(ref[k]))
            assert P.rel_l2(got[k], torch.from_numpy(ref[k])) < 5e-3, k


@pytest.mark.parametrize("tag,act,prior", [("elu", "elu", "normal_zero_mean_one_std"), ("noprior", "relu", False),
                                           ("tanh", "tanh", "normal_zero_mean_one_std"), ("sigmoid", "sigmoid", "normal_zero_mean_one_std")])
def test_reference_variant_fixture(tag, act, prior):
    """tests/golden/ref_small_variants.npz (generated by the reference itself): ELU hidden activations, and latent_prior_type=False
    (z = encoder output, no KL term) -- forward outputs, both phases' losses and every gradient."""
    g = np.load(os.path.join(G, "ref_small_variants.npz"))
    cfg = dict(dsb=int(g["dsb"]), da=int(g["da"]), z=int(g["z"]), te=(16, 2), md=(24, 3), wm=(32, 2))
    layers = {k: orc.gen_layers(cfg[k][0], cfg[k][1], act_hidden=act) for k in ("te", "md", "wm")}
    sd = {k[len(tag) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + "/sd/")}
    m = P.product_model(cfg, layers, sd, prior=prior, max_batch=128, act=act)
    x = torch.Tensor(g["X"])[:, 0, :].cuda()
    eps = torch.from_numpy(g["eps"]).cuda()
    logits, state = m(input_dict={"obs": x, "obs_flat": x, "eps": eps}, state=None, seq_lens=None)
    P.assert_close("logits", logits, torch.from_numpy(g[tag + "/logits"]))
    P.assert_close("z", m._cur_task_encoder_variable, torch.from_numpy(g[tag + "/z_task"]))
    P.assert_close("future", m._cur_future_state, torch.from_numpy(g[tag + "/future"]))
    P.assert_close("value", m.value_function(), torch.from_numpy(g[tag + "/value"]))
    eng = m.engine()
    B = int(g["B"])
    eng.alloc_transitions(B)
    eng.ingest(torch.from_numpy(g["X"]).cuda(), torch.from_numpy(g["Y"]).cuda())
    for world in (True, False):
        ph = tag + ("/world" if world else "/vae")
        m.set_learnable_task_encoder(not world); m.set_learnable_motor_decoder(not world); m.set_learnable_world_model(world)
        eng.set_cursor(0)
        loss = eng.world_step(B) if world else eng.vae_step(B, eps=eps, kl_coeff=1.0, cyc_coeff=0.05)
        torch.cuda.synchronize()
        ref_loss = float(g[ph + "/loss"])
        assert abs(float(loss[0]) - ref_loss) <= P.ATOL + P.RTOL * abs(ref_loss), (ph, float(loss[0]), ref_loss)
        ref = {k[len(ph) + 6:]: g[k] for k in g.files if k.startswith(ph + "/grad/")}
        got = P.grads_by_net(m)
        assert set(got) == set(ref)
        for k in ref:
            assert P.rel_l2(got[k], torch.from_numpy(ref[k])) < 5e-3, (ph, k, P.rel_l2(got[k], torch.from_numpy(ref[k])))


# ---- forward / inference API ----------------------------------------------------------------------------------------------
def test_forward_parts_and_aliases():
    from physicsvae_b200 import torch_models as tm
    om, layers = P.oracle_model(P.SMALL, out_std=0.3)
    m = P.product_model(P.SMALL, layers, om.state_dict(), max_batch=256)
    X, Y = P.transitions(P.SMALL, 77)
    x = torch.from_numpy(X).float()
    eps = torch.randn(77, 8, generator=torch.Generator().manual_seed(1))
    logits = om.forward(x, eps=eps)
    zb, zt, _ = m.forward_encoder(x.cuda(), [], None, 0, eps=eps.cuda())
    P.assert_close("z_task", zt, om.cur["z_task"])
    assert torch.equal(zb.cpu(), x[:, :37])
    lg, _ = m.forward_decoder(zb, zt, [], None, 0)
    P.assert_close("logits", lg, logits)
    P.assert_close("world", m.forward_world(x.cuda(), lg), om.cur["future"])
    P.assert_close("world(s1, a)", m.forward_world(x[:, :37].cuda(), torch.from_numpy(Y).cuda()),
                   om.forward_world(x[:, :37], torch.from_numpy(Y)))
    v, _ = m.forward_value_branch(x.cuda(), [], None, 0)
    P.assert_close("value", v[:, 0], om.cur["value"])
    # latent_prior_noise False -> z = mu (rllib_model_torch.py:734-740); pass-through decoding from z ~ N(0, I)
    m.latent_prior_noise = False
    _, zt2, _ = m.forward_encoder(x.cuda(), [], None, 0)
    P.assert_close("z = mu", zt2, om.cur["mu"])
    P.assert_close("Motor", tm.Motor(m)(x[:, :37].cuda(), x[:, 37:].cuda()), om._fc("_motor_decoder", torch.cat([x[:, :37], om.cur["mu"]], 1)))
    P.assert_close("WorldModel", tm.WorldModel(m)(x[:, :37].cuda(), torch.from_numpy(Y).cuda()), om.forward_world(x[:, :37], torch.from_numpy(Y)))
    # Philox noise mode: deterministic in (seed, offset), standard normal
    eng = m.engine()
    a = eng.forward(x.cuda(), 1, noise=True, seed=5, offset=9)
    b = eng.forward(x.cuda(), 1, noise=True, seed=5, offset=9)
    c = eng.forward(x.cuda(), 1, noise=True, seed=5, offset=10)
    assert torch.equal(a["z"], b["z"]) and not torch.equal(a["z"], c["z"])


def test_philox_noise_is_standard_normal():
    om, layers = P.oracle_model(P.SMALL)
    m = P.product_model(P.SMALL, layers, om.state_dict(), max_batch=16384)
    x = torch.zeros(16384, 74, device="cuda")
    out = m.engine().forward(x, 1, noise=True, seed=3, offset=0)
    e = (out["z"] - out["mu"]) / torch.exp(0.5 * out["logvar"])
    assert abs(float(e.mean())) < 0.02 and abs(float(e.std()) - 1.0) < 0.02
    assert abs(float((e ** 4).mean()) - 3.0) < 0.2


# ---- trainer level: trajectories, phase switch, checkpoints -----------------------------------------------------------------
def _pickle(tmp_path, n_ep=3, T=41, dsb=13, da=5, seed=2):
    data = orc.synthetic_episodes(n_ep, T, dsb, da, seed=seed)
    f = str(tmp_path / "demo.pkl")
    pickle.dump(data, open(f, "wb"))
    return f, data


def test_trainer_trajectory_phase_switch_and_checkpoints(tmp_path):
    from physicsvae_b200 import train_physics_vae as tp
    f, data = _pickle(tmp_path)
    tp.args = tp.arg_parser().parse_args(["--data_train", f, "--max_iter_world_model", "2", "--max_iter", "4", "--batch_size", "32",
                                          "--latent_dim", "4"])
    cfg = tp.resolve_grid(tp.get_trainer_config(tp.args))[0]
    cfg.update(TE_width=16, MD_width=24, world_model_width=32, noise_seed=1)
    torch.manual_seed(5)
    tr = tp.TrainModel(cfg)
    torch.manual_seed(5)
    om = orc.OracleModel(13, 5, 4, orc.gen_layers(16, 2), orc.gen_layers(24, 3), orc.gen_layers(32, 2))
    for k, v in tr.model.state_dict().items():
        assert torch.equal(v.cpu(), om.params[k]), k
    X, Y = orc.build_transitions(data["episodes"])
    ot = orc.OracleTrainer(om, X, Y, batch_size=32, lr=5e-4, max_iter_world_model=2)
    assert len(tr.train_loader) == len(ot.batches()) == 4          # 120 transitions -> 32, 32, 32, 24
    tr.model.latent_prior_noise = False                             # deterministic z = mu on both sides
    om.latent_prior_noise = False
    for it in range(4):
        r = tr.train()
        o = ot.step()
        assert r["training_iteration"] == it + 1
        assert abs(r["mean_train_loss"] - o["mean_train_loss"]) <= P.ATOL + P.RTOL * abs(o["mean_train_loss"]), (it, r, o)
        assert tr.world_phase == ot.world == (it < 2)
    sd = tr.model.state_dict()
    for k, v in om.params.items():
        P.assert_close("param " + k, sd[k], v, rtol=2e-3, atol=2e-5)
    # frozen-phase bookkeeping: the value branch never trains; Adam state exists only for parameters that saw a gradient
    assert all(p.grad is None for p in tr.model._value_branch.parameters())
    assert len(tr.optimizer.state) == 2 * (3 + 3 + 4)
    # checkpoint file set + interchange with the oracle (= reference key layout)
    path = tr.save(str(tmp_path / "ckpt"))
    assert sorted(os.listdir(tmp_path / "ckpt")) == ["model.pt", "model.pth", "motor_decoder.pt", "task_encoder.pt", "world_model.pt"]
    assert path.endswith("model.pth")
    om2 = orc.OracleModel(13, 5, 4, orc.gen_layers(16, 2), orc.gen_layers(24, 3), orc.gen_layers(32, 2))
    om2.load_state_dict(torch.load(path))
    tr2 = tp.TrainModel(cfg)
    tr2.restore(path)
    x = torch.from_numpy(X[:16, 0, :]).float()
    tr2.model.latent_prior_noise = False
    om2.latent_prior_noise = False
    P.assert_close("restored forward", tr2.compute_model(x.cuda()), om2.forward(x)[:, :5])


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_device_side_dataset_build_equals_array_ingest(precision, tmp_path):
    """pvae_ingest_episodes (unique states + index) must fill the resident buffer bit-identically to pvae_ingest
    on the reference-format X / Y arrays, and the trainer must take that path for a pickled dataset."""
    from physicsvae_b200 import train_physics_vae as tp
    from physicsvae_b200.engine import Engine
    f, data = _pickle(tmp_path, n_ep=5, T=23, dsb=13, da=5, seed=4)
    X, Y = tp.episodes_to_transitions(data["episodes"])
    states, actions, first = tp.episodes_to_index(data["episodes"])
    nets = {"task_encoder": [(8, "relu"), (8, "linear")], "motor_decoder": [(8, "relu"), (5, "linear")],
            "world_model": [(8, "relu"), (13, "linear")], "value_branch": [(8, "relu"), (1, "linear")]}
    eng = Engine(13, 5, 4, nets, precision=precision, max_batch=64)
    n = len(X)
    buf = eng.alloc_transitions(n)
    eng.ingest(torch.from_numpy(X[:, 0, :]).cuda(), torch.from_numpy(Y[:, 0, :]).cuda())
    torch.cuda.synchronize()
    a = buf.clone()
    buf.zero_()
    half = n // 2 + 3                                       # two calls with a destination offset, like the chunked upload
    eng.ingest_episodes(torch.from_numpy(states).cuda(), torch.from_numpy(actions).cuda(), torch.from_numpy(first[:half]).cuda())
    eng.ingest_episodes(torch.from_numpy(states).cuda(), torch.from_numpy(actions).cuda(), torch.from_numpy(first[half:]).cuda(), dst_row=half)
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.uint8).flatten(), buf.view(torch.uint8).flatten())
    with pytest.raises(ValueError):
        eng.ingest_episodes(torch.from_numpy(states).cuda(), torch.from_numpy(actions).cuda(), torch.tensor([len(states) - 1]).cuda())
    # a loader that keeps the dataset in bf16 (the engine's operand precision): same resident bytes as converting fp32 on the device
    s32, a32 = torch.from_numpy(states).float(), torch.from_numpy(actions).float()
    if precision == "bf16":
        buf.zero_()
        eng.ingest_episodes(s32.cuda(), a32.cuda(), torch.from_numpy(first).cuda())
        torch.cuda.synchronize()
        b32 = buf.clone()
        buf.zero_()
        eng.ingest_episodes(s32.bfloat16().cuda(), a32.bfloat16().cuda(), torch.from_numpy(first).cuda())
        torch.cuda.synchronize()
        assert torch.equal(b32.view(torch.uint8).flatten(), buf.view(torch.uint8).flatten())
    else:
        with pytest.raises(ValueError):
            eng.ingest_episodes(s32.bfloat16().cuda(), a32.bfloat16().cuda(), torch.from_numpy(first).cuda())
    ds = tp.load_dataset_for_PhysicsVAE([f])
    assert ds.episode_source is not None and len(ds.episode_source[2]) == len(ds) == n


def test_trainer_trajectory_matches_reference_fixture(tmp_path):
    """tests/golden/ref_trajectory.npz holds four epochs of the REFERENCE's own TrainModel.step (oracle/make_golden.py): two
    world-model epochs, the phase switch, two VAE epochs, Adam + StepLR, batches 32 / 32 / 32 / 24.  The CUDA trainer, started
    from the fixture's initial weights, must reproduce the per-epoch losses and the final parameters."""
    from physicsvae_b200 import train_physics_vae as tp
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_trajectory.npz"))
    f = str(tmp_path / "demo.pkl")
    with open(f, "wb") as fh:
        fh.write(g["pickle_bytes"].tobytes())
    tp.args = tp.arg_parser().parse_args(["--data_train", f, "--max_iter_world_model", str(int(g["max_iter_world_model"])), "--max_iter", "4",
                                          "--batch_size", str(int(g["batch_size"])), "--latent_dim", str(int(g["z"])), "--lr", str(float(g["lr"]))])
    cfg = tp.resolve_grid(tp.get_trainer_config(tp.args))[0]
    cfg.update(TE_width=int(g["TE_width"]), MD_width=int(g["MD_width"]), world_model_width=int(g["world_model_width"]), noise_seed=1)
    assert cfg["lr_schedule_params"] == {"step_size": int(g["lr_step_size"]), "gamma": float(g["lr_gamma"])}
    tr = tp.TrainModel(cfg)
    tr.model.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init/")})
    tr.model.latent_prior_noise = False
    assert len(tr.train_loader) == 4
    for it, want in enumerate(g["losses"]):
        r = tr.train()
        assert abs(r["mean_train_loss"] - want) <= P.ATOL + P.RTOL * abs(want), (it, r, float(want))
    sd = tr.model.state_dict()
    for k in g.files:
        if k.startswith("final/"):
            P.assert_close("param " + k[6:], sd[k[6:]], torch.from_numpy(g[k]), rtol=2e-3, atol=2e-5)


def test_resume_continues_the_trajectory(tmp_path):
    """run_trial with --resume semantics: 2 iterations + checkpoint + a NEW trainer restored from it + 3 more iterations ==
    5 iterations straight (weights, Adam moments, StepLR counter, phase switch at iteration 3 and the iteration counter all
    travel in trainer_state.pt; upstream a resumed trial restarts the optimizer)."""
    from physicsvae_b200 import train_physics_vae as tp
    f, data = _pickle(tmp_path)
    tp.args = tp.arg_parser().parse_args(["--data_train", f, "--max_iter_world_model", "3", "--max_iter", "5", "--batch_size", "32",
                                          "--latent_dim", "4"])
    cfg = tp.resolve_grid(tp.get_trainer_config(tp.args))[0]
    cfg.update(TE_width=16, MD_width=24, world_model_width=32, noise_seed=1, lr_schedule_params={"step_size": 2, "gamma": 0.5})
    torch.manual_seed(5)
    straight, _ = tp.run_trial(cfg, 5, 0, str(tmp_path / "straight"))
    torch.manual_seed(5)
    first, ck = tp.run_trial(cfg, 2, 2, str(tmp_path / "resumed"))
    assert ck == tp.latest_checkpoint(str(tmp_path / "resumed")) and ck.endswith(os.path.join("checkpoint_000002", "model.pth"))
    assert "trainer_state.pt" in os.listdir(os.path.dirname(ck))
    torch.manual_seed(99)                                   # a different init: everything must come from the checkpoint
    second, _ = tp.run_trial(cfg, 5, 0, str(tmp_path / "resumed"), restore=ck)
    assert second.training_iteration == 5 and second.iter == 5 and not second.world_phase
    assert abs(second.optimizer.param_groups[0]["lr"] - straight.optimizer.param_groups[0]["lr"]) < 1e-12
    a, b = straight.model.state_dict(), second.model.state_dict()
    for k in a:
        P.assert_close("resumed " + k, b[k], a[k], rtol=1e-4, atol=1e-6)
    lines = open(tmp_path / "resumed" / "result.json").read().strip().splitlines()
    assert [json.loads(l)["training_iteration"] for l in lines] == [1, 2, 3, 4, 5]


def test_cli_main_grid_checkpoints_output_and_resume(tmp_path):
    """The CLI end to end with the reference's flags (train_physics_vae.py:30-55, 469-521): a two-point grid (list flags append
    to their defaults), per-trial result.json + checkpoint directories with the reference's five files, a working --output
    export (it always fails upstream), and --resume continuing every trial from its newest checkpoint."""
    from physicsvae_b200 import train_physics_vae as tp
    f, data = _pickle(tmp_path)
    out = str(tmp_path / "exported.pt")
    base = ["--data_train", f, "--max_iter_world_model", "1", "--batch_size", "32", "--latent_dim", "4", "--vae_kl_coeff", "0.5",
            "--local_dir", str(tmp_path / "results"), "--name", "cli", "--checkpoint_freq", "1"]
    ck = tp.main(base + ["--max_iter", "2", "--output", out])
    root = tmp_path / "results" / "cli"
    assert sorted(os.listdir(root)) == ["trial_00000", "trial_00001"]          # vae_kl_coeff grid [1.0, 0.5]
    assert ck == str(root / "trial_00001" / "checkpoint_000002" / "model.pth")
    for t in ("trial_00000", "trial_00001"):
        res = [json.loads(l) for l in open(root / t / "result.json").read().strip().splitlines()]
        assert [r["training_iteration"] for r in res] == [1, 2] and all(np.isfinite(r["mean_train_loss"]) for r in res)
        for c in ("checkpoint_000001", "checkpoint_000002"):
            assert sorted(os.listdir(root / t / c)) == ["model.pt", "model.pth", "motor_decoder.pt", "task_encoder.pt", "trainer_state.pt",
                                                        "world_model.pt"]
    # a sweep with several grid points exports one file per trial (a single path would hold whichever trial finished last)
    assert not os.path.exists(out)
    ref = torch.load(ck)
    for t in (0, 1):
        sd = torch.load(str(tmp_path / ("exported.trial_%05d.pt" % t)))
        assert len(sd) == 26 and all(k.split(".")[0] in ("_task_encoder", "_motor_decoder", "_world_model", "_value_branch") for k in sd)
        want = torch.load(str(root / ("trial_%05d" % t) / "checkpoint_000002" / "model.pth"))
        assert all(torch.equal(sd[k], want[k]) for k in want)
    # a single grid point keeps the path as given
    one = str(tmp_path / "single.pt")
    tp.main(["--data_train", f, "--max_iter_world_model", "1", "--batch_size", "32", "--latent_dim", "4", "--local_dir", str(tmp_path / "results1"),
             "--name", "one", "--max_iter", "1", "--output", one])
    assert os.path.exists(one)
    ck3 = tp.main(base + ["--max_iter", "3", "--resume"])
    assert ck3 == str(root / "trial_00001" / "checkpoint_000003" / "model.pth")
    for t in ("trial_00000", "trial_00001"):
        res = [json.loads(l) for l in open(root / t / "result.json").read().strip().splitlines()]
        assert [r["training_iteration"] for r in res] == [1, 2, 3]
    moved = [k for k in ref if k.startswith("_motor_decoder") and not torch.equal(torch.load(ck3)[k], ref[k])]
    assert moved, "the resumed VAE-phase iteration must have trained the decoder"


@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_fused_adam_matches_torch_adam_and_refreshes_shadow(weight_decay):
    """physicsvae_b200.optim.PvaeAdam (one kernel per layer: Adam update + bf16 shadow refresh) vs torch.optim.Adam on the same
    gradients, including the lazy per-net state and frozen sub-nets; afterwards the engine's forward must see the new weights."""
    from physicsvae_b200.optim import PvaeAdam
    om, layers = P.oracle_model(P.SMALL, out_std=0.3)
    m = P.product_model(P.SMALL, layers, om.state_dict(), max_batch=128)
    eng = m.engine()
    m.sync_weights()
    ref = {k: v.detach().clone().cuda().requires_grad_(True) for k, v in m.state_dict().items()}
    topt = torch.optim.Adam(list(ref.values()), lr=3e-3, weight_decay=weight_decay)
    opt = PvaeAdam(m, lr=3e-3, weight_decay=weight_decay)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for it in range(5):
        world = it < 2                       # phase switch: world model first, then encoder + decoder
        m.set_learnable_task_encoder(not world); m.set_learnable_motor_decoder(not world); m.set_learnable_world_model(world)
        for k, p in m.named_parameters():
            if p.grad is not None:
                p.grad.copy_(torch.randn(p.shape, generator=gen, device="cuda") * 1e-2)
                ref[k].grad = p.grad.clone()
            else:
                ref[k].grad = None
        opt.step()
        topt.step()
    for k, p in m.named_parameters():
        P.assert_close("adam " + k, p, ref[k], rtol=1e-5, atol=1e-7)
    assert len(opt.state) == len(topt.state) == 2 * (3 + 3 + 4)
    x = torch.randn(64, 74, device="cuda")
    a = eng.forward(x, 15, noise=False)          # shadow operands as refreshed by the optimizer kernels
    m.sync_weights()                             # full refresh from the fp32 masters
    b = eng.forward(x, 15, noise=False)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_compute_loss_reference_signature(tmp_path):
    from physicsvae_b200 import train_physics_vae as tp
    f, data = _pickle(tmp_path)
    tp.args = tp.arg_parser().parse_args(["--data_train", f, "--batch_size", "32", "--latent_dim", "4"])
    cfg = tp.resolve_grid(tp.get_trainer_config(tp.args))[0]
    cfg.update(TE_width=16, MD_width=24, world_model_width=32)
    torch.manual_seed(7)
    tr = tp.TrainModel(cfg)
    torch.manual_seed(7)
    om = orc.OracleModel(13, 5, 4, orc.gen_layers(16, 2), orc.gen_layers(24, 3), orc.gen_layers(32, 2))
    X, Y = orc.build_transitions(data["episodes"], num_samples=50)
    x, y = torch.Tensor(X), torch.Tensor(Y)                         # [B, 1, 2*dsb], [B, 1, da] like the DataLoader's batches
    loss = tr.compute_loss(y, x)
    loss.backward()                                                  # no-op: gradients are already deposited
    o_loss, _, o_grads = orc.loss_and_grads(om, x[:, 0, :], y[:, 0, :], True)
    assert abs(float(loss) - o_loss) <= P.ATOL + P.RTOL * abs(o_loss)
    got = P.grads_by_net(tr.model)
    assert set(got) == set(o_grads)
    for k in o_grads:
        P.assert_close(k, got[k], o_grads[k])
    with pytest.raises(ValueError):
        tr.compute_loss(y[:1], x[:1])                                # the reference fails on B == 1 too (squeeze)


def test_fake_multi_rank_equals_single_rank():
    """Data-parallel equivalence on one GPU: per-shard steps with weighted coefficients, summed / R == the full batch."""
    from physicsvae_b200 import parallel
    om, layers = P.oracle_model(P.SMALL, out_std=0.3)
    m = P.product_model(P.SMALL, layers, om.state_dict(), max_batch=256)
    X, Y = P.transitions(P.SMALL, 203)
    eng = m.engine()
    eng.alloc_transitions(203)
    eng.ingest(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda())
    m.sync_weights()
    m.set_learnable_task_encoder(False); m.set_learnable_motor_decoder(False)
    eng.set_cursor(0)
    eng.world_step(203)
    full, full_loss = m.flat_grads("world_model").clone(), float(eng.loss[0])
    R = 4
    acc, loss = torch.zeros_like(full), 0.0
    for r in range(R):
        s, e = parallel.shard_rows(0, 203, r, R)
        eng.set_cursor(s)
        eng.world_step(e - s, s_coeff=parallel.shard_weight(0, 203, r, R))
        acc += m.flat_grads("world_model")
        loss += float(eng.loss[0])
    assert abs(loss / R - full_loss) <= 1e-5 * abs(full_loss)
    assert P.rel_l2(acc / R, full) < 1e-4


# ---- full-size properties (BASELINE.json configs[1] / [2] sizes; the oracle would take minutes here) ---------------------------
@pytest.mark.parametrize("world", [True, False])
def test_full_size_batch_additivity_and_loss_consistency(world):
    """At batch 65536 (bf16): (i) the reported loss equals the MSE recomputed from the forward API's outputs; (ii) the
    gradient of the whole batch equals the weighted mean of the gradients of its two halves (linearity of the mean loss in
    the batch); (iii) a repeated step gives the same loss bit for bit."""
    B = 65536
    cfg = P.DEFAULT
    om, layers = P.oracle_model(cfg, out_std=0.1)
    m = P.product_model(cfg, layers, om.state_dict(), precision="bf16", max_batch=B)
    X, Y = P.transitions(cfg, B, seed=3)
    eng = m.engine()
    eng.alloc_transitions(B)
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    eng.ingest(Xd, Yd)
    m.sync_weights()
    m.set_learnable_task_encoder(not world); m.set_learnable_motor_decoder(not world); m.set_learnable_world_model(world)
    nets = ["world_model"] if world else ["task_encoder", "motor_decoder"]

    def run(lo, n, w=1.0):
        eng.set_cursor(lo)
        if world:
            eng.world_step(n, s_coeff=w)
        else:
            eng.vae_step(n, noise=False, a_coeff=w, kl_coeff=w, cyc_coeff=0.05 * w)
        return torch.cat([m.flat_grads(k) for k in nets]).clone(), eng.loss.clone()
    g_full, l_full = run(0, B)
    g_again, l_again = run(0, B)
    assert torch.equal(l_full, l_again)
    g_a, l_a = run(0, B // 2)
    g_b, l_b = run(B // 2, B // 2)
    assert P.rel_l2(0.5 * (g_a + g_b), g_full) < 2e-3
    assert abs(float(0.5 * (l_a[0] + l_b[0]) - l_full[0])) <= 1e-4 * abs(float(l_full[0]))
    xb = Xd.float().bfloat16().float()
    if world:
        fut = m.forward_world(xb[:, :cfg["dsb"]], Yd.bfloat16().float())
        mse = float(((fut - xb[:, cfg["dsb"]:]) ** 2).mean())
        assert abs(mse - float(l_full[3])) <= 2e-3 * abs(mse)
    else:
        m.latent_prior_noise = False
        logits, _ = m(input_dict={"obs": xb, "obs_flat": xb}, state=None, seq_lens=None)
        mse_a = float(((logits[:, :cfg["da"]] - Yd.bfloat16().float()) ** 2).mean())
        mu, lv = m._cur_task_encoder_mu, m._cur_task_encoder_logvar
        kl = float(torch.mean(-0.5 * torch.sum(1 + lv - mu ** 2 - lv.exp(), dim=1)))
        assert abs(mse_a - float(l_full[1])) <= 2e-3 * abs(mse_a) and abs(kl - float(l_full[2])) <= 2e-3 * abs(kl) + 1e-6
