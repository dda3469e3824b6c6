"""Data-parallel plumbing: the only multi-GPU strategy the path has.  One process per GPU; every
global mini-batch [lo, hi) is cut into contiguous per-rank row ranges; each rank's loss coefficients are weighted by
n_rank * R / n so that the AVERAGE of the per-rank gradients (one all-reduce over the flat fp32 gradient buffers) equals
the gradient of the global mean loss exactly, also for ragged last batches.  torch.distributed (NCCL over NVLink on GPUs,
gloo in the CPU tests) is plumbing; there is no other collective on the path."""
import torch
import torch.distributed as dist


_replicas = False      # sweep mode: the ranks of the job are independent trainers, no gradient exchange


def set_replica_mode(flag):
    """Sweep over grid points with one independent trainer per GPU (the role of Ray Tune's parallel trials,
    train_physics_vae.py:264-285, 484-502): data-parallel sharding and the all-reduce are switched off."""
    global _replicas
    _replicas = bool(flag)


def job_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def job_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def world_size():
    """Ranks that share one trainer's mini-batches (1 in replica mode)."""
    return 1 if _replicas else job_world_size()


def rank():
    return 0 if _replicas else job_rank()


def sweep_points(n_points, r, world):
    """Grid points trainer r of `world` replicas runs: round-robin, so that every point runs exactly once."""
    return [i for i in range(n_points) if i % world == r]


def shard_rows(lo, hi, r, world):
    """Contiguous slice of rows [lo, hi) owned by rank r of `world`: sizes differ by at most one, earlier ranks larger."""
    n = hi - lo
    base, rem = divmod(n, world)
    start = lo + r * base + min(r, rem)
    return start, start + base + (1 if r < rem else 0)


def max_shard_rows(batch_size, world):
    return (batch_size + world - 1) // world


def shard_weight(lo, hi, r, world):
    """n_r * R / n: multiply the rank's loss coefficients by this and all-reduce-AVERAGE the gradients."""
    s, e = shard_rows(lo, hi, r, world)
    return (e - s) * world / float(hi - lo)


def allreduce_avg_(tensors, group=None):
    """In-place average of the given flat tensors over all ranks, one collective per tensor.  A training step passes ONE
    tensor: the contiguous [gradients | loss slots] range of the model's gradient pool."""
    w = world_size()
    if w == 1:
        return
    native_avg = dist.get_backend(group) == "nccl"      # NCCL averages inside the collective; gloo has no AVG
    for t in tensors:
        if native_avg:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t.div_(w)


def barrier():
    """All ranks that share a trainer wait for each other (no-op for a single rank and in replica mode)."""
    if world_size() > 1:
        dist.barrier()


def broadcast_(tensors, src=0):
    if world_size() == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src)
