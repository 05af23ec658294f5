"""Measurement tool (not the bench, not the product): the gradient all-reduce alone, at the sizes of BASELINE C4 / C3
(500k / 1M Gaussians x 59 floats), plain vs ncclMemAlloc-registered operand (and torch's symmetric-memory multimem
all-reduce as a library reference point), under whatever NCCL_* environment the launcher sets.

    python -m torch.distributed.run --nproc-per-node N tools/allreduce_sweep.py [tag]
"""
import datetime
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist


def timed(fn, barrier, reps=20, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return statistics.median(ts), min(ts)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "default"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
    res = {"tag": tag, "world": world, "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}
    for name, n in (("c4_118MB", 500_000 * 59), ("c3_236MB", 1_000_000 * 59)):
        plain = torch.zeros(n, device=dev)
        med, best = timed(lambda: dist.all_reduce(plain), dist.barrier)
        res[name + "_plain_ms"] = (med, best)
        try:
            backend = dist.group.WORLD._get_backend(dev)
            pool = torch.cuda.MemPool(backend.mem_allocator)
            with torch.cuda.use_mem_pool(pool):
                reg = torch.zeros(n, device=dev)
            backend.register_mem_pool(pool)
            med, best = timed(lambda: dist.all_reduce(reg), dist.barrier)
            res[name + "_registered_ms"] = (med, best)
            del reg
        except Exception as ex:
            res[name + "_registered_ms"] = repr(ex)[:200]
        try:
            import torch.distributed._symmetric_memory as symm

            t = symm.empty(n, dtype=torch.float32, device=dev)
            symm.rendezvous(t, dist.group.WORLD.group_name)
            t.zero_()
            med, best = timed(lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name),
                              dist.barrier)
            res[name + "_symm_multimem_ms"] = (med, best)
        except Exception as ex:
            res[name + "_symm_multimem_ms"] = repr(ex)[:200]
        del plain
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
