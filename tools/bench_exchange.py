"""The gradient-exchange kernel alone, under torchrun: correctness against NCCL's all-reduce, microseconds per call and the
algorithmic NVLink rate, for the variant the environment selects (PVAE_SYMM_BULK=0/1, PVAE_SYMM_CTAS=n, PVAE_SYMM_MULTIMEM=1).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_exchange.py [MB ...]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from physicsvae_b200 import parallel  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [float(a) for a in sys.argv[1:]] or [6.0, 24.0]
    out = []
    for mb in sizes:
        n = int(mb * 1e6 / 4) // 4 * 4
        pool = parallel.SymmetricPool(n, dev)
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        src = torch.randn(n, device=dev, generator=g)
        ref = src.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.AVG)
        pool.view.copy_(src)
        torch.cuda.synchronize(); dist.barrier()
        pool.allreduce_avg_(pool.view)
        torch.cuda.synchronize(); dist.barrier()
        err = float((pool.view - ref).norm() / ref.norm())
        # replicas bit-identical?
        chk = pool.view.double().sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same = all(float(c) == float(allc[0]) for c in allc)
        for _ in range(20):
            pool.allreduce_avg_(pool.view)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 200
        e0.record()
        for _ in range(iters):
            pool.allreduce_avg_(pool.view)
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        us = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
        dist.all_reduce(us, op=dist.ReduceOp.MAX)
        us = float(us)
        out.append({"MB": mb, "us_per_call": round(us, 2), "rel_l2_vs_nccl": err, "replicas_identical": same,
                    "nvlink_GBps_per_direction": round(n * 4 * (world - 1) / world / (us * 1e-6) / 1e9, 1)})
        del pool
    if rank == 0:
        print(json.dumps({"world": world, "bulk": os.environ.get("PVAE_SYMM_BULK", "0"), "ctas": os.environ.get("PVAE_SYMM_CTAS", "64"),
                          "multimem": os.environ.get("PVAE_SYMM_MULTIMEM", "0"), "results": out}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
