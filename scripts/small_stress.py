"""Many small / ragged fused steps back to back (the overlapped weight-gradient GEMM with K ranges shorter than its ring):
python scripts/small_stress.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digs_b200  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = digs_b200.DiGSNetwork(0, 3, 256, "sine", "none", 4, "mfgi", [1.6, 0.1], precision="bf16x2").to(dev)
crit = digs_b200.DiGSLoss([3e3, 1e2, 1e2, 5e1, 1e2], "siren_wo_n_w_div", "linear", "l1", 50)
for p in net.parameters():
    p.grad = torch.zeros_like(p)
g = torch.Generator(device="cpu").manual_seed(1)
n_done = 0
for rep in range(6):
    for n in (1, 2, 23, 24, 25, 31, 33, 64, 100, 257, 700, 1500, 4099, 9000):
        nm_ = max(1, n - 2 * rep)
        m = torch.rand(1, n, 3, generator=g).to(dev) - 0.5
        nm = (torch.rand(1, nm_, 3, generator=g) * 2.2 - 1.1).to(dev)
        nrm = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1).to(dev)
        ld, out, _ = digs_b200.fused_step(net, crit, m, nm, nrm)
        n_done += 1
    torch.cuda.synchronize()
    print("rep", rep, "ok", n_done, float(ld["loss"]), flush=True)
print("finite grads:", all(bool(torch.isfinite(p.grad).all()) for p in net.parameters()))
