"""Test infrastructure: the CPU oracle of the PhysicsVAE hot path and the harness that pins it to the reference.
Nothing under physicsvae_b200/ may import this package."""
