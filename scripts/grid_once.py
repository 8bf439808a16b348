"""Runs one grid query (for ncu captures): python scripts/grid_once.py [res] [mode]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digs_b200  # noqa: E402
import bench  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16x2"
torch.manual_seed(bench.MODEL_SEED)
net = digs_b200.DiGSNetwork(**bench.NET_KW, precision=mode).cuda()
axes, _, _ = digs_b200.grid_axes(res, np.array([[-1, 1], [-1, 1], [-1, 1]], dtype=float))
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vol = digs_b200.grid_query(net, axes)
    e1.record()
    torch.cuda.synchronize()
    print("grid %d^3: %.3f ms" % (res, e0.elapsed_time(e1)))
