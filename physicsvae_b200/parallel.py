"""Data-parallel plumbing: the only multi-GPU strategy the path has.  One process per GPU; every
global mini-batch [lo, hi) is cut into contiguous per-rank row ranges; each rank's loss coefficients are weighted by
n_rank * R / n so that the AVERAGE of the per-rank gradients (ONE all-reduce over the contiguous [gradients | loss slots]
range of the model's gradient pool) equals the gradient of the global mean loss exactly, also for ragged last batches.

The all-reduce itself: when the ranks share one NVLink / NVSwitch node the gradient pool is a SYMMETRIC allocation
(torch.distributed._symmetric_memory: the same buffer mapped into every peer) and the exchange is the library's own single
kernel over peer memory (pvae_symm_allreduce, include/pvae_sm100.h: device-side rank barrier, rank r reduces slice r -- pulled from
the peers by the bulk-copy engine, by plain peer loads, or by multimem.ld_reduce through the switch -- and stores it into every replica).  torch.distributed's NCCL all-reduce
(gloo in the CPU tests) is the fallback and the parity path (PVAE_SYMM_AR=0); torch.distributed is plumbing -- rendezvous,
broadcast of the initial parameters, barriers."""
import ctypes as C
import os
import sys

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


class SymmetricPool(object):
    """An fp32 buffer allocated symmetrically on every rank of the node and mapped into every peer, plus the flag block the
    library's all-reduce kernel synchronises through.  `view` is what the model uses as its gradient pool."""

    def __init__(self, n, device):
        import torch.distributed._symmetric_memory as symm
        from . import _abi
        self.lib = _abi.load()
        group = dist.group.WORLD
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.n = int(n)
        self.flags_off = (self.n + 3) // 4 * 4
        total = self.flags_off + int(self.lib.pvae_symm_flag_elems())
        try:
            symm.enable_symm_mem_for_group(group.group_name)          # needed by older releases, a deprecated no-op in newer ones
        except Exception:  # noqa
            pass
        self.buf = symm.empty(total, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.buf.data_ptr():
            raise RuntimeError("unexpected symmetric-memory pointer table")
        self.peer = (C.c_uint64 * self.world)(*ptrs)
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.mc = mc if os.environ.get("PVAE_SYMM_MULTIMEM", "0") == "1" else 0       # opt-in: reduce inside the switch (NVLS)
        self.has_multicast = bool(mc)
        self.view = self.buf[:self.n]
        torch.cuda.synchronize(device)
        dist.barrier()                                 # every rank's flag block is zeroed before anyone can signal into it

    def contains(self, t):
        b = self.buf.data_ptr()
        return t.dtype == torch.float32 and b <= t.data_ptr() and t.data_ptr() + t.numel() * 4 <= b + self.flags_off * 4

    def _span(self, t):
        off = (t.data_ptr() - self.buf.data_ptr()) // 4
        count = (t.numel() + 3) // 4 * 4               # pieces of the pool are padded to 16 bytes: the tail words are zeros
        if off % 4 or off + count > self.flags_off:
            raise ValueError("range is not 16-byte aligned inside the symmetric pool")
        return off, count

    def arm_overlap(self, engine, early):
        """Let the engine's training steps exchange `early` themselves, beside their last backward GEMMs (pvae_set_exchange);
        early None switches it off.  Returns True when armed."""
        from . import _abi
        ctas = int(os.environ.get("PVAE_OVERLAP_SMS", "8"))
        # Opt-in (PVAE_OVERLAP=1).  Measured with the bulk-copy exchange kernel (profiles/r02_scaling.md): nothing at N = 2 (0.5933 vs
        # 0.5915 ms world), -1.3 % on the world step at N = 8 (0.5895 vs 0.5973), a LOSS on the wide config's 24 MB of gradients
        # (0.388 vs 0.379 ms at N = 8): the GEMMs give up 8 SMs for longer than the exchange saves.
        if early is None or os.environ.get("PVAE_OVERLAP", "0") != "1" or not self.contains(early):
            _abi.check(self.lib.pvae_set_exchange(engine._h, None, 0, 0, 0, 0, 0, 0, 0))
            return False
        off, count = self._span(early)
        _abi.check(self.lib.pvae_set_exchange(engine._h, self.peer, self.mc, self.rank, self.world, off, count, self.flags_off, ctas))
        return True

    def allreduce_avg_(self, t):
        off, count = self._span(t)
        from . import _abi
        _abi.check(self.lib.pvae_symm_allreduce(self.peer, self.mc, self.rank, self.world, off, count, self.flags_off,
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))


_symm_pools = []
_symm_note = {"kind": "nccl", "why": "single rank"}


def symmetric_pool_factory():
    """A `grad_pool_factory` for PhysicsVAE (rllib_model_torch.py: _pool_alloc) that hands out symmetric memory, or None when the
    job cannot use it (one rank, replica mode, not NCCL, more than 8 ranks, PVAE_SYMM_AR=0).  Allocation is collective: every
    rank of the job builds its model / engine at the same points of the program."""
    if world_size() == 1 or not dist.is_initialized():
        return None
    if os.environ.get("PVAE_SYMM_AR", "1") == "0":
        _symm_note.update(kind="nccl", why="PVAE_SYMM_AR=0")
        return None
    if dist.get_backend() != "nccl" or dist.get_world_size() > 8:
        _symm_note.update(kind=str(dist.get_backend()), why="symmetric memory needs NCCL ranks on one node (<= 8)")
        return None

    def factory(n, device):
        try:
            pool = SymmetricPool(n, device)
        except Exception as e:  # noqa  (no peer access / no fabric handle support: every rank fails alike)
            _symm_note.update(kind="nccl", why="symmetric allocation failed: %s" % repr(e)[:160])
            if job_rank() == 0:
                print("[physicsvae_b200] symmetric gradient pool unavailable (%s); using NCCL all-reduce" % repr(e)[:160], file=sys.stderr)
            return torch.zeros(n, dtype=torch.float32, device=device)
        _symm_pools[:] = [p for p in _symm_pools if p.buf is not None][-3:] + [pool]      # keep the last few alive (engines get re-created)
        bulk = os.environ.get("PVAE_SYMM_BULK", "1") != "0"       # (read by the library: bulk-copy engine moves the slices, the default)
        _symm_note.update(kind="symm-multimem" if pool.mc else ("symm-p2p-bulk" if bulk else "symm-p2p"),
                          why="multicast available" if pool.has_multicast else "no multicast binding")
        return pool.view
    return factory


def pool_of(t):
    """The symmetric pool a tensor lives in, or None."""
    return next((p for p in reversed(_symm_pools) if p.contains(t)), None)


def allreduce_kind():
    """What the gradient exchange of this job runs on: "symm-p2p-bulk" / "symm-p2p" / "symm-multimem" (the library's own kernels over peer memory),
    "nccl" / "gloo" (torch.distributed), and why."""
    return dict(_symm_note)


def graph_capturable():
    """Can a training step with its collective be captured into a CUDA graph?  (gloo synchronises on the host.)"""
    return world_size() == 1 or dist.get_backend() == "nccl"


def allreduce_avg_(tensors, group=None):
    """In-place average of the given flat tensors over all ranks, one collective per tensor.  A training step passes ONE
    tensor: the contiguous [gradients | loss slots] range of the model's gradient pool -- exchanged by the library's own
    peer-memory kernel when that pool is a symmetric allocation, by torch.distributed otherwise."""
    w = world_size()
    if w == 1:
        return
    native_avg = dist.get_backend(group) == "nccl"      # NCCL averages inside the collective; gloo has no AVG
    for t in tensors:
        pool = next((p for p in reversed(_symm_pools) if p.contains(t)), None) if group is None else None
        if pool is not None:
            pool.allreduce_avg_(t)
        elif native_avg:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t.div_(w)


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index, sysfs="/sys", bdf=None):
    """Run this process on the CPUs of the NUMA node its GPU hangs off, so that the pinned staging buffers it allocates afterwards
    (first touch) are local to the GPU's PCIe root: with one loader process per GPU, uploads that cross the socket interconnect
    share its bandwidth between all ranks.  Best effort -- returns a record of what was done (or why nothing was)."""
    try:
        if bdf is None:
            pr = torch.cuda.get_device_properties(device_index)
            bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return {"bound": False, "why": "no NUMA node reported for %s" % bdf}
        with open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)) as f:
            local = _cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        use = allowed & local
        if not use:
            return {"bound": False, "node": node, "why": "none of the %d CPUs this process may use is on node %d" % (len(allowed), node)}
        if use != allowed:
            os.sched_setaffinity(0, use)
        return {"bound": True, "node": node, "cpus": len(use), "of": len(allowed), "pci": bdf}
    except Exception as e:                                           # (no sysfs in the container, no such attribute, ...)
        return {"bound": False, "why": "%s: %s" % (type(e).__name__, e)}


def barrier():
    """All ranks that share a trainer wait for each other (no-op for a single rank and in replica mode)."""
    if world_size() > 1:
        dist.barrier()


def broadcast_(tensors, src=0):
    if world_size() == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src)
