"""The data-parallel CLI path with TWO ranks on one GPU (gloo moves the CUDA tensors through the host -- NCCL refuses two ranks
on one device): parameters broadcast from rank 0 although every rank seeds its own RNG, sharded mini-batches + averaged gradients
== the single-rank trainer, rank 0 alone writes result.json / checkpoints / --output.  (The NCCL / peer-memory path itself is
checked on real multi-GPU boxes by bench.py's dp_check.)"""
import json
import os
import pickle
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pvae_oracle as orc  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, mode):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      PVAE_DIST_BACKEND="gloo")
    import torch
    from physicsvae_b200 import train_physics_vae as tp
    torch.manual_seed(1000 + 17 * rank)                    # every rank draws DIFFERENT initial weights
    f = os.path.join(tmp, "demo.pkl")
    base = ["--data_train", f, "--max_iter_world_model", "1", "--batch_size", "33", "--latent_dim", "4", "--local_dir",
            os.path.join(tmp, "results"), "--name", "dp", "--checkpoint_freq", "1", "--max_iter", "2"]
    if mode == "trainer":
        tp.args = tp.arg_parser().parse_args(base)
        tp.init_distributed("dp")
        cfg = tp.resolve_grid(tp.get_trainer_config(tp.args))[0]
        cfg.update(TE_width=16, MD_width=24, world_model_width=32, noise_seed=3)
        tr = tp.TrainModel(cfg)
        tr.model.latent_prior_noise = False
        losses = [tr.train()["mean_train_loss"] for _ in range(3)]
        torch.save({"losses": losses, "sd": {k: v.detach().cpu() for k, v in tr.model.state_dict().items()}}, os.path.join(tmp, "rank%d.pt" % rank))
    else:
        ck = tp.main(base + ["--output", os.path.join(tmp, "exported.pt")])
        with open(os.path.join(tmp, "main_rank%d.json" % rank), "w") as fh:
            json.dump({"checkpoint": ck}, fh)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def _spawn(tmp, mode):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp), mode), nprocs=2, join=True)


def _write_pickle(tmp):
    data = orc.synthetic_episodes(3, 41, 13, 5, seed=2)
    with open(os.path.join(str(tmp), "demo.pkl"), "wb") as fh:
        pickle.dump(data, fh)
    return data


def test_two_rank_dp_trainer_equals_single_rank(tmp_path):
    _write_pickle(tmp_path)
    _spawn(tmp_path, "trainer")
    r0, r1 = torch.load(str(tmp_path / "rank0.pt")), torch.load(str(tmp_path / "rank1.pt"))
    for k in r0["sd"]:
        assert torch.equal(r0["sd"][k], r1["sd"][k]), k              # replicas stay bit-identical
    assert r0["losses"] == r1["losses"]
    # the same training on one rank, started from rank 0's seed
    from physicsvae_b200 import train_physics_vae as tp
    torch.manual_seed(1000)
    base = ["--data_train", str(tmp_path / "demo.pkl"), "--max_iter_world_model", "1", "--batch_size", "33", "--latent_dim", "4"]
    tp.args = tp.arg_parser().parse_args(base)
    cfg = tp.resolve_grid(tp.get_trainer_config(tp.args))[0]
    cfg.update(TE_width=16, MD_width=24, world_model_width=32, noise_seed=3)
    tr = tp.TrainModel(cfg)
    tr.model.latent_prior_noise = False
    losses = [tr.train()["mean_train_loss"] for _ in range(3)]
    for a, b in zip(r0["losses"], losses):
        assert abs(a - b) <= 1e-4 * abs(b), (r0["losses"], losses)
    sd = {k: v.detach().cpu() for k, v in tr.model.state_dict().items()}
    for k in sd:
        d = float((sd[k].double() - r0["sd"][k].double()).norm() / (sd[k].double().norm() + 1e-30))
        assert d < 2e-4, (k, d)


def test_two_rank_cli_writes_on_rank0_only(tmp_path):
    _write_pickle(tmp_path)
    _spawn(tmp_path, "main")
    root = tmp_path / "results" / "dp" / "trial_00000"
    res = [json.loads(l) for l in open(root / "result.json").read().strip().splitlines()]
    assert [r["training_iteration"] for r in res] == [1, 2]            # one line per iteration, not one per rank
    cks = [json.load(open(tmp_path / ("main_rank%d.json" % r)))["checkpoint"] for r in (0, 1)]
    assert cks[0] == cks[1] == str(root / "checkpoint_000002" / "model.pth")
    exported = torch.load(str(tmp_path / "exported.pt"))
    want = torch.load(cks[0])
    assert all(torch.equal(exported[k], want[k]) for k in want)
