"""Smallest full hot-path invocation (GPU): ragged clouds, every kernel family, for compute-sanitizer runs."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import digs_b200

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x2"
torch.manual_seed(0)
net = digs_b200.DiGSNetwork(0, 3, 256, "sine", "none", 2, "mfgi", [1.6, 0.1], precision=mode).cuda()
crit = digs_b200.DiGSLoss([3e3, 1e2, 1e2, 5e1, 1e2], "siren_w_div", "linear", "l1", 50)
m = (torch.rand(1, 301, 3, device="cuda") - 0.5).requires_grad_()
nm = (torch.rand(1, 203, 3, device="cuda") * 2.2 - 1.1).requires_grad_()
nrm = torch.nn.functional.normalize(torch.randn(1, 301, 3, device="cuda"), dim=-1)
ld, _ = crit(net(nm, m), m, nm, nrm)
ld["loss"].backward()
axes, _, _ = digs_b200.grid_axes(20, np.array([[-1, 1], [-1, 1], [-1, 1]], dtype=float))
vol = digs_b200.grid_query(net, axes)
with torch.no_grad():
    v = net.decoder(torch.rand(77, 3, device="cuda"))
torch.cuda.synchronize()
print("ok", float(ld["loss"]), float(vol.mean()), float(v.mean()), all(torch.isfinite(p.grad).all() for p in net.parameters()))
