#!/usr/bin/env python
"""Probe (torchrun): is symmetric memory + NVLS multicast available, and how do NCCL all_reduce and torch's multimem all-reduce
compare on the gradient-buffer size?  Usage: torchrun --nproc-per-node N tools/nvls_probe.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = 3_000_000 * 73
x = torch.randn(n, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


t_nccl = timeit(lambda: dist.all_reduce(x))
if rank == 0:
    print("NCCL all_reduce %d MB: %.3f ms" % (n * 4 // 2**20, t_nccl), flush=True)
try:
    group = dist.group.WORLD
    symm_mem.enable_symm_mem_for_group(group.group_name)
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, group.group_name)
    if rank == 0:
        print("symm_mem ok: multicast_ptr = %#x, world %d" % (hdl.multicast_ptr, hdl.world_size), flush=True)
    t.copy_(x)
    if hdl.multicast_ptr:
        t_mm = timeit(lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", group.group_name))
        if rank == 0:
            print("multimem_all_reduce_: %.3f ms" % t_mm, flush=True)
    t_two = timeit(lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", group.group_name))
    if rank == 0:
        print("two_shot_all_reduce_: %.3f ms" % t_two, flush=True)
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print("symm_mem unavailable:", repr(e)[:400], flush=True)
dist.destroy_process_group()
