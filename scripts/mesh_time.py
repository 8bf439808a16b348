"""Timing helper (GPU): 512^3 grid query -> marching tetrahedra -> surface sampling -> Chamfer, all on the device."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import digs_b200
from digs_b200 import mesh as dmesh

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(3627473)
net = digs_b200.DiGSNetwork(0, 3, 256, "sine", "none", 4, "mfgi", [1.6, 0.1]).cuda()
net.precision = "bf16x2"

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3, out

axes, _, _ = digs_b200.grid_axes(res, np.array([[-1, 1], [-1, 1], [-1, 1]], dtype=float))
ms_q, vol = timed(lambda: digs_b200.grid_query(net.decoder, axes))
n = len(axes[0])
ms_mt, (v, f, nrm) = timed(lambda: dmesh.marching_tets(vol.view(n, n, n), 0.0, (1, 1, 1), (0, 0, 0), (1, 0, 2)))
ms_all, out = timed(lambda: digs_b200.implicit2mesh(net.decoder, None, res, device="cuda"), reps=2)
ms_s, (pts, _) = timed(lambda: dmesh.sample_surface(v, f, 1000000))
gt = torch.randn(1000000, 3, device="cuda"); gt = 0.5 * gt / gt.norm(dim=1, keepdim=True)
ms_nn, d = timed(lambda: dmesh.nn_query(gt, pts), reps=2)
ms_cd, cd = timed(lambda: dmesh.compute_dists(pts, gt), reps=2)
print("res %d: grid query %.1f ms | marching tets %.1f ms (%d verts, %d faces) | implicit2mesh %.1f ms | sample 1M %.1f ms | "
      "nn 1Mx1M %.1f ms | compute_dists %.1f ms" % (res, ms_q, ms_mt, len(v), len(f), ms_all, ms_s, ms_nn, ms_cd))
