"""Throw-away stand-in for ray[rllib]==1.11.0 (not installed, no network) so that the reference's own files import
UNCHANGED from /root/reference.  Only what train_physics_vae.py / torch_models.py / rllib_model_torch.py touch at import
and on the hot path is provided; semantics restated from the ray 1.11 sources cited in SURVEY.md section 8c.
TEST INFRASTRUCTURE ONLY -- never imported by the physicsvae_b200 package."""
from . import tune  # noqa: F401


def init(*args, **kwargs):
    return None


def shutdown(*args, **kwargs):
    return None
