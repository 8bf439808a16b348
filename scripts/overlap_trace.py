"""Timeline of the fused step's two grids (make DIAG=TRACE build): when do the tile-kernel CTAs exit, when do the
weight-gradient CTAs start / pass their wait / end, on which SM.  python scripts/overlap_trace.py [splits]"""
import copy
import os
import sys

import numpy as np
import torch

os.environ["DIGS_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digs_b200  # noqa: E402
from digs_b200 import synth, step as step_mod  # noqa: E402
import bench  # noqa: E402

cfg = bench.CONFIGS["c2"]
dev = torch.device("cuda:0")
torch.manual_seed(bench.MODEL_SEED)
net = digs_b200.DiGSNetwork(**cfg["net"], precision="bf16x2").to(dev)
crit = digs_b200.DiGSLoss(**copy.deepcopy(cfg["loss"]))
cloud, normals = synth.surface_cloud(100000, "torus", seed=0)
m, nrm, nm = (torch.tensor(x, device=dev) for x in synth.sample_batch(cloud, normals, cfg["n_points"], seed=1, box=cfg["box"]))
for p in net.parameters():
    p.grad = torch.zeros_like(p)
for r in range(4):
    digs_b200.fused_step(net, crit, m, nm, nrm)
torch.cuda.synchronize()
ws = step_mod._trace_ws
tr = ws[-8192 * 8:].view(torch.int64).cpu().numpy()
t_end, t_start = tr[:148], tr[256:256 + 148]
t0 = t_start.min()
print("tile CTAs: start spread %.1f us; exits at (us): min %.1f  p25 %.1f  median %.1f  p75 %.1f  max %.1f" % (
    (t_start.max() - t0) / 1e3, *[(np.percentile(t_end, q) - t0) / 1e3 for q in (0, 25, 50, 75, 100)]))
w = tr[512:].reshape(-1, 4)
w = w[w[:, 0] > 0]
print("wgrad CTAs traced:", len(w))
order = np.argsort(w[:, 0])
w = w[order]
print("wgrad start (us): first %.1f  p25 %.1f  median %.1f  p75 %.1f  last %.1f" % tuple((np.percentile(w[:, 0], q) - t0) / 1e3 for q in (0, 25, 50, 75, 100)))
print("wgrad wait  (us): mean %.1f  max %.1f" % ((w[:, 1] - w[:, 0]).mean() / 1e3, (w[:, 1] - w[:, 0]).max() / 1e3))
print("wgrad loads (us): mean %.1f  min %.1f  max %.1f" % ((w[:, 2] - w[:, 1]).mean() / 1e3, (w[:, 2] - w[:, 1]).min() / 1e3, (w[:, 2] - w[:, 1]).max() / 1e3))
print("wgrad last load issued at (us): %.1f" % ((w[:, 2].max() - t0) / 1e3))
per_sm = {}
for row in w:
    per_sm.setdefault(int(row[3]), []).append(row)
cnt = np.array([len(v) for v in per_sm.values()])
print("units per SM: min %d  median %d  max %d over %d SMs" % (cnt.min(), np.median(cnt), cnt.max(), len(cnt)))
for i in range(0, len(w), max(1, len(w) // 40)):
    print("  cta#%3d start %.1f waited %.1f ran %.1f sm %d" % (order[i], (w[i, 0] - t0) / 1e3, (w[i, 1] - w[i, 0]) / 1e3, (w[i, 2] - w[i, 1]) / 1e3, w[i, 3]))
