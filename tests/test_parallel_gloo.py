"""world_size-2 data-parallel equivalence on CPU (gloo): per-rank shards + weighted coefficients + one averaging
all-reduce over the flat gradient buffers == the single-rank gradient of the global mini-batch.
The gradient producer here is the CPU oracle (the engine itself needs a GPU); the sharding / weighting / collective code
is the product's (physicsvae_b200.parallel)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pvae_oracle as orc
from physicsvae_b200 import parallel

CFG = dict(dsb=13, da=5, z=4)


def _model():
    torch.manual_seed(0)
    return orc.OracleModel(CFG["dsb"], CFG["da"], CFG["z"], orc.gen_layers(16, 2), orc.gen_layers(24, 3), orc.gen_layers(32, 2), orc.gen_layers(16, 2))


def _batch(B):
    import numpy as np
    X, Y = orc.build_transitions(orc.synthetic_episodes(2, 41, CFG["dsb"], CFG["da"], seed=3)["episodes"], num_samples=B)
    return torch.from_numpy(X).float()[:, 0, :], torch.from_numpy(Y)[:, 0, :], torch.randn(B, CFG["z"], generator=torch.Generator().manual_seed(4))


def _flat(grads, keys):
    return torch.cat([grads[k].flatten() for k in keys])


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        m = _model()
        x, y, eps = _batch(B)
        res = {}
        for world_phase in (True, False):
            s, e = parallel.shard_rows(0, B, rank, world)
            w = parallel.shard_weight(0, B, rank, world)
            loss, _, grads = orc.loss_and_grads(m, x[s:e], y[s:e], world_phase, eps=eps[s:e], cyc_coeff=0.05)
            keys = sorted(grads)
            flat = _flat(grads, keys) * w                 # == running the step with every loss coefficient scaled by w
            lossbuf = torch.tensor([loss * w])
            parallel.allreduce_avg_([flat, lossbuf])
            res[world_phase] = (keys, flat, float(lossbuf))
        if rank == 0:
            torch.save(res, out)
        assert parallel.world_size() == world and parallel.rank() == rank
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(B, tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    res = torch.load(out)
    m = _model()
    x, y, eps = _batch(B)
    for world_phase in (True, False):
        loss, _, grads = orc.loss_and_grads(m, x, y, world_phase, eps=eps, cyc_coeff=0.05)
        keys, flat, dloss = res[world_phase]
        assert keys == sorted(grads)
        ref = _flat(grads, keys)
        assert abs(dloss - loss) <= 1e-6 * abs(loss)
        assert float((flat - ref).norm() / ref.norm()) < 1e-5


def test_dp2_equals_single_rank_even_batch(tmp_path):
    _run(64, tmp_path)


def test_dp2_equals_single_rank_ragged_batch(tmp_path):
    _run(37, tmp_path)      # shards of 19 and 18 rows: the n_r * R / n weights make the average exact


def _replica_worker(rank, world, port, out):
    """`train_physics_vae.py --sweep_mode replicas` under torchrun: the CLI's own init (gloo here: no GPU) + point assignment."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    from physicsvae_b200 import train_physics_vae as tp
    try:
        tp.init_distributed("replicas")
        assert dist.is_initialized() and parallel.job_world_size() == world and parallel.job_rank() == rank
        assert parallel.world_size() == 1 and parallel.rank() == 0          # a replica trains alone: no sharding, no all-reduce
        t = torch.ones(3)
        parallel.allreduce_avg_([t])                                        # must be a no-op in replica mode
        assert torch.equal(t, torch.ones(3))
        mine = parallel.sweep_points(5, parallel.job_rank(), parallel.job_world_size())
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            torch.save(gathered, out)
        tp.init_distributed("dp")
        assert parallel.world_size() == world and parallel.rank() == rank
    finally:
        parallel.set_replica_mode(False)
        if dist.is_initialized():
            dist.destroy_process_group()


def test_sweep_replica_mode_under_two_ranks(tmp_path):
    out = str(tmp_path / "points.pt")
    mp.spawn(_replica_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    parts = torch.load(out)
    assert parts == [[0, 2, 4], [1, 3]]
