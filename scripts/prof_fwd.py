"""Profiling driver (GPU): forward-jet launches of the C2 off-surface cloud (saving / not saving) + reverse pass."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import digs_b200
from digs_b200 import jet as djet, _lib

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x2"
torch.manual_seed(3627473)
net = digs_b200.DiGSNetwork(0, 3, 256, "sine", "none", 4, "mfgi", [1.6, 0.1], precision=mode).cuda()
fc = net.decoder.fc_block
params = [p.detach() for p in fc.flat_params()]
xyz = (torch.rand(15000, 3, device="cuda") * 2.2 - 1.1)
for _ in range(2):
    djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, True, False, fc.mode())
for _ in range(2):
    f, g, l, saved = djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, True, True, fc.mode())
for _ in range(2):
    djet.jet_backward_raw(fc.spec(), params, xyz, None, 0, True, torch.ones_like(f), torch.ones_like(g), torch.ones_like(f), saved, fc.mode())
axes, _, _ = digs_b200.grid_axes(128, np.array([[-1, 1], [-1, 1], [-1, 1]], dtype=float))
for _ in range(2):
    digs_b200.grid_query(net, axes)
torch.cuda.synchronize()
print("ok")
