"""Stage what the GPU box needs from the reference into the git-ignored baseline/_ref/ (it travels with gpurun, /root/reference
does not): the three hot-path files -- UNCHANGED, byte for byte -- for bench.py's `--impl reference` arm (they import under
oracle/ref_stub exactly like in this container), and the shipped checkpoint for the known-answer test on the CUDA path.
Nothing here enters the git history; the product never reads baseline/_ref/.  TEST / BENCH INFRASTRUCTURE ONLY.

  python oracle/stage_ref.py        (no-op where /root/reference does not exist)
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("PVAE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["train_physics_vae.py", "torch_models.py", "rllib_model_torch.py", os.path.join("data", "pretrained", "loco_modelV1.pt")]


def stage(verbose=True):
    if not os.path.isfile(os.path.join(REFERENCE, FILES[0])):
        if verbose:
            print("[stage_ref] %s not present: nothing staged" % REFERENCE)
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        src, dst = os.path.join(REFERENCE, f), os.path.join(DST, os.path.basename(f))
        if not os.path.exists(dst) or hashlib.sha256(open(src, "rb").read()).digest() != hashlib.sha256(open(dst, "rb").read()).digest():
            shutil.copyfile(src, dst)
        if verbose:
            print("[stage_ref] %s -> %s" % (src, os.path.relpath(dst, ROOT)))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() or True else 1)
