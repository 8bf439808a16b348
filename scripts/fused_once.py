"""Runs the fused step a few times at the benchmark size (for ncu captures): python scripts/fused_once.py [reps] [mode]"""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digs_b200  # noqa: E402
from digs_b200 import synth  # noqa: E402
import bench  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16x2"
cfg_name = sys.argv[3] if len(sys.argv) > 3 else "c2"
cfg = bench.CONFIGS[cfg_name]
dev = torch.device("cuda:0")
torch.manual_seed(bench.MODEL_SEED)
net = digs_b200.DiGSNetwork(**cfg["net"], precision=mode).to(dev)
crit = digs_b200.DiGSLoss(**copy.deepcopy(cfg["loss"]))
cloud, normals = synth.surface_cloud(100000, "torus", seed=0)
bs = cfg["batch"]
parts = [synth.sample_batch(cloud, normals, cfg["n_points"], seed=1 + s, box=cfg["box"]) for s in range(bs)]
m, nrm, nm = (torch.tensor(np.concatenate([p[i] for p in parts], 0), device=dev) for i in range(3))
lat = None
if cfg["net"]["encoder_type"] == "autodecoder":
    lat = 0.01 * torch.randn(bs, cfg["net"]["latent_size"], device=dev)
for p in net.parameters():
    p.grad = torch.zeros_like(p)
evs = []
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ld, out, _ = digs_b200.fused_step(net, crit, m, nm, nrm, latent=lat)
    e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
print("loss", float(ld["loss"]), "ms", [round(a.elapsed_time(b), 4) for a, b in evs])
