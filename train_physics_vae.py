"""CLI entry point with the reference's name and flags: `python train_physics_vae.py --data_train demo.pkl ...`"""
from physicsvae_b200.train_physics_vae import main

if __name__ == "__main__":
    main()
