"""Probe (torchrun, >= 2 GPUs): latency of the flat-gradient all-reduce (265k fp32) -- NCCL vs symmetric-memory one-shot /
multimem (NVLS) reductions, eager and inside a CUDA graph."""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
n = 264449 + 8
group = dist.group.WORLD
t = symm_mem.empty(n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, group)
plain = torch.empty(n, dtype=torch.float32, device=dev)
out = torch.empty(n, dtype=torch.float32, device=dev)
gname = group.group_name


def fill():
    t.fill_(float(rank + 1))
    plain.fill_(float(rank + 1))


variants = {
    "nccl_all_reduce": lambda: dist.all_reduce(plain),
    "symm_one_shot_out": lambda: torch.ops.symm_mem.one_shot_all_reduce_out(t, "sum", gname, out),
    "symm_two_shot_": lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", gname),
    "symm_multimem_": lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname),
}
expect = float(sum(range(1, world + 1)))
for name, fn in variants.items():
    try:
        fill()
        torch.cuda.synchronize()
        dist.barrier()
        fn()
        torch.cuda.synchronize()
        got = float((out if name == "symm_one_shot_out" else (plain if name.startswith("nccl") else t))[123])
        ok = abs(got - expect) < 1e-6
        # eager timing
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        eager_us = 1e3 * e0.elapsed_time(e1) / 50
        # graph timing
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for _ in range(10):
                fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_us = 1e3 * e0.elapsed_time(e1) / 100
        if rank == 0:
            print("%-20s correct=%s  eager %.1f us  in-graph %.1f us" % (name, ok, eager_us, graph_us), flush=True)
        del g
    except Exception as ex:
        if rank == 0:
            print("%-20s FAILED: %s: %s" % (name, type(ex).__name__, str(ex)[:200]), flush=True)
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
