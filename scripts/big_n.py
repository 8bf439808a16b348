"""Scale check of the fused step: 1 000 000 + 1 000 000 points (4x256) and 262 144 + 262 144 points on the 8x512 scene network,
plus tiny / ragged clouds.  python scripts/big_n.py"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digs_b200  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda:0")
for cfg_name, n in (("c2", 1000000), ("c4", 262144), ("c2", 1), ("c2", 3), ("c2", 23), ("c2", 25)):
    cfg = bench.CONFIGS[cfg_name]
    torch.manual_seed(1)
    net = digs_b200.DiGSNetwork(**cfg["net"], precision="bf16x2").to(dev)
    crit = digs_b200.DiGSLoss(**copy.deepcopy(cfg["loss"]))
    g = torch.Generator(device="cpu").manual_seed(n)
    v = torch.randn(1, n, 3, generator=g)
    v = v / v.norm(dim=-1, keepdim=True)
    m, nrm = (0.5 * v).to(dev), v.to(dev)
    nm = (torch.rand(1, n, 3, generator=g) * 2.2 - 1.1).to(dev)
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    torch.cuda.reset_peak_memory_stats()
    ev = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ld, out, _ = digs_b200.fused_step(net, crit, m, nm, nrm)
        e1.record()
        ev.append((e0, e1))
    torch.cuda.synchronize()
    ms = min(a.elapsed_time(b) for a, b in ev)
    fl = bench.flops_per_step(cfg, n, n)
    finite = all(bool(torch.isfinite(p.grad).all()) for p in net.parameters()) and bool(torch.isfinite(ld["loss"]))
    print("%s %8d + %8d points: %9.3f ms/step  %7.2f M points/s  %6.1f TFLOP/s algorithmic  peak memory %.1f GB  finite=%s" % (
        cfg_name, n, n, ms, 2 * n / ms / 1e3, fl / ms / 1e9, torch.cuda.max_memory_allocated() / 1e9, finite))
    del net, out, ld
    torch.cuda.empty_cache()
