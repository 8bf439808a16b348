"""Timing helper (GPU): forward jet with / without saving, reverse pass, per precision mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import digs_b200
from digs_b200 import _lib, jet as djet

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

torch.manual_seed(0)
net = digs_b200.DiGSNetwork(0, 3, 256, "sine", "none", 4, "mfgi", [1.6, 0.1]).cuda()
fc = net.decoder.fc_block
params = [p.detach() for p in fc.flat_params()]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15000
xyz = (torch.rand(n, 3, device="cuda") * 2.2 - 1.1)
for name, mode in (("fp32", 0), ("bf16x2", 1), ("bf16", 2)):
    for lap in (True, False):
        C = 5 if lap else 4
        gf = n * C * 2 * 4 * 256 * 256 / 1e9
        t1 = timeit(lambda: djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, lap, True, mode))
        t0 = timeit(lambda: djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, lap, False, mode))
        f, g, l, saved = djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, lap, True, mode)
        fb, gb = torch.randn_like(f), torch.randn_like(g)
        lb = torch.randn_like(f) if lap else None
        tb = timeit(lambda: djet.jet_backward_raw(fc.spec(), params, xyz, None, 0, lap, fb, gb, lb, saved, mode))
        print("%-7s C=%d n=%d  fwd(save) %.3f ms (%.0f TF/s)  fwd(nosave) %.3f ms (%.0f TF/s)  bwd %.3f ms (%.0f TF/s)" %
              (name, C, n, t1, gf / t1, t0, gf / t0, tb, 2 * gf / tb), flush=True)
