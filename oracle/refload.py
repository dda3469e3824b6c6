"""Import the reference's own three hot-path files UNCHANGED from /root/reference under oracle/ref_stub (ray / gym are
not installed).  Only possible in the build container -- /root/reference does not exist on the GPU box -- so this is
used by oracle/make_golden.py and by the `-m "not gpu"` test that pins the oracle to the live reference.
TEST INFRASTRUCTURE ONLY."""
import argparse
import importlib
import os
import sys

import numpy as np

REFERENCE = os.environ.get("PVAE_REFERENCE", "/root/reference")
STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_stub")


def available():
    return os.path.isfile(os.path.join(REFERENCE, "train_physics_vae.py"))


_mods = None


def load():
    """Returns (train_physics_vae, torch_models, rllib_model_torch) reference modules."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference not found at %s" % REFERENCE)
    if not hasattr(np, "product"):
        np.product = np.prod               # removed in NumPy 2; used at rllib_model_torch.py:384, 599-604
    saved_path = list(sys.path)
    saved = {k: sys.modules.pop(k) for k in ("train_physics_vae", "torch_models", "rllib_model_torch") if k in sys.modules}
    for k in [k for k in sys.modules if k == "ray" or k.startswith("ray.") or k == "gym" or k.startswith("gym.")]:
        del sys.modules[k]
    sys.path[:0] = [REFERENCE, STUB]
    try:
        rmt = importlib.import_module("rllib_model_torch")
        tm = importlib.import_module("torch_models")
        tpv = importlib.import_module("train_physics_vae")
    finally:
        sys.path[:] = saved_path
    # keep the reference modules importable under private names only; give the public names back
    for k in ("train_physics_vae", "torch_models", "rllib_model_torch"):
        sys.modules["_pvae_ref_" + k] = sys.modules.pop(k)
    sys.modules.update(saved)
    tpv.args = argparse.Namespace(num_data=None)     # module global read at train_physics_vae.py:339
    _mods = (tpv, tm, rmt)
    return _mods


def build_reference_model(dsb, da, z, te_layers, md_layers, wm_layers, vf_layers=None, prior="normal_zero_mean_one_std"):
    """Construct the reference PhysicsVAE exactly like train_physics_vae.create_model does."""
    tpv, tm, rmt = load()
    from gym.spaces import Box  # the stub
    box = lambda n: Box(low=-np.ones(n), high=np.ones(n), dtype=np.float64)
    cfg = dict(rmt.PhysicsVAE.DEFAULT_CONFIG)
    cfg.update(observation_space=box(2 * dsb), observation_space_body=box(dsb), observation_space_task=box(dsb),
               action_space=box(da), task_encoder_output_dim=z, task_encoder_layers=te_layers,
               motor_decoder_layers=md_layers, world_model_layers=wm_layers, latent_prior_type=prior)
    if vf_layers is not None:
        cfg["value_fn_layers"] = vf_layers
    return rmt.PhysicsVAE(obs_space=box(2 * dsb), action_space=box(da), num_outputs=2 * da,
                          model_config={"custom_model_config": cfg}, name="physics_vae")


class _Shell(object):
    """Just enough of train_physics_vae.TrainModel to call the reference's unbound compute_loss on a model."""


def reference_compute_loss(model, x, y, world, kl_coeff=1.0, cyc_coeff=1e-3, prior="normal_zero_mean_one_std", lookahead=1):
    """Calls the reference's own TrainModel.compute_loss (train_physics_vae.py:361-435) on x [B,1,2dsb], y [B,1,da]."""
    import torch
    tpv, tm, rmt = load()
    sh = _Shell()
    sh.model = model
    sh.lookahead = lookahead
    sh.latent_prior_type = prior
    sh.loss_fn = tm.get_loss_fn("MSE")
    sh.vae_kl_coeff = 0.0 if world else kl_coeff
    sh.a_rec_coeff = 0.0 if world else 1.0
    sh.s_rec_coeff = 1.0 if world else 0.0
    sh.vae_cycle_coeff = 0.0 if world else cyc_coeff
    sh.compute_model = lambda x_: tpv.TrainModel.compute_model(sh, x_)
    return tpv.TrainModel.compute_loss(sh, y, x)
