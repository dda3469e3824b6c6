"""physicsvae_b200 -- B200-native (sm_100a) implementation of the PhysicsVAE training hot path.

Host-side mirror of the reference's Python surface (rllib_model_torch.FC / PhysicsVAE, torch_models.TrainModel,
train_physics_vae CLI) over libpvae_sm100.so (include/pvae_sm100.h).  See DESIGN.md.
"""
__version__ = "0.1.0"
