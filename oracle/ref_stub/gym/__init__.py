"""Stand-in for gym: only spaces.Box(low, high, dtype) with .shape is used on the hot path (train_physics_vae.py:216-233)."""
from . import spaces  # noqa: F401
