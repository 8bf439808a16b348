"""Runs the REFERENCE's own training script with the module swapped: `import models.DiGS as DiGS` resolves to digs_b200.

    python examples/reference_loop_swap.py [--iters 200] [--precision bf16x2] [--n_points 15000]

The script executed is the unmodified surface_reconstruction/train_surface_reconstruction.py of Chumbyte/DiGS (from
/root/reference, or from the copy `__graft_entry__.build()` installs under the git-ignored baseline/_ref/): its argparse, its
ReconDataset + DataLoader workers, its loop (`net(nonmnfld, mnfld)` -> `criterion(...)` -> `loss.backward()` ->
`clip_grad_norm_` -> `Adam.step()` -> `update_div_weight`, train_surface_reconstruction.py:93-158), its per-step logging and
its periodic `utils.implicit2mesh(DiGSNet.decoder, ...)` snapshots.  What is replaced:
  * `models.DiGS`  -> digs_b200 (DiGSNetwork / DiGSLoss drop-ins; precision from DIGS_B200_PRECISION),
  * packages that are not installed in this image and are not on the hot path get empty stand-ins: tensorboardX (a
    SummaryWriter that swallows scalars), plotly, matplotlib, plyfile, trimesh, skimage.measure (so the mesh snapshots stop
    AFTER the decoder has been queried on the 128^3 grid, inside the script's own try/except), open3d (its
    `io.read_point_cloud` hands back a synthetic 100 000-point torus cloud with analytic normals: there is no dataset here).
Prints one JSON line: iterations, seconds, loss of the first and last logged iteration.
"""
import argparse
import json
import os
import re
import runpy
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference_root():
    for cand in (os.environ.get("DIGS_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "surface_reconstruction", "train_surface_reconstruction.py")):
            return cand
    raise SystemExit("reference training script not found (run __graft_entry__.build() where /root/reference exists)")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Writer:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


class _Cloud:
    def __init__(self, points, normals):
        self.points, self.normals = points, normals


def install_stubs():
    from digs_b200 import synth
    _stub("tensorboardX", SummaryWriter=_Writer)
    _stub("plyfile", PlyData=object)
    _stub("trimesh")
    sk = _stub("skimage")
    sk.measure = _stub("skimage.measure")
    pl = _stub("plotly")
    for sub in ("graph_objs", "graph_objects", "figure_factory", "offline"):
        setattr(pl, sub, _stub("plotly." + sub))
    mp = _stub("matplotlib")
    mp.pyplot = _stub("matplotlib.pyplot")
    pts, nrm = synth.surface_cloud(100000, "torus", seed=0)
    o3d = _stub("open3d")
    o3d.io = _stub("open3d.io", read_point_cloud=lambda path: _Cloud(pts * 0.37 + 0.05, nrm))   # un-normalised on purpose


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--n_points", type=int, default=15000)
    ap.add_argument("--precision", default=os.environ.get("DIGS_B200_PRECISION", "bf16x2"))
    ap.add_argument("--workers_ok", type=int, default=1)
    args = ap.parse_args()
    ref = reference_root()
    os.environ["DIGS_B200_PRECISION"] = args.precision
    install_stubs()
    import digs_b200
    sys.path.insert(0, ref)
    import models                                   # the reference's package; its DiGS submodule is replaced
    sys.modules["models.DiGS"] = digs_b200
    models.DiGS = digs_b200
    logdir = tempfile.mkdtemp(prefix="digs_swap_")
    script_dir = os.path.join(ref, "surface_reconstruction")
    sys.path.insert(0, script_dir)
    sys.argv = ["train_surface_reconstruction.py", "--logdir", logdir, "--file_name", "synthetic_torus.ply",
                "--grid_res", "64", "--loss_type", "siren_wo_n_w_div", "--inter_loss_type", "exp", "--num_epochs", "0",
                "--gpu_idx", "0", "--n_samples", str(args.iters), "--n_points", str(args.n_points), "--batch_size", "1",
                "--lr", "5e-5", "--nonmnfld_sample_type", "grid", "--dataset_path", logdir,
                "--decoder_n_hidden_layers", "4", "--decoder_hidden_dim", "256", "--div_decay", "linear",
                "--div_decay_params", "1e2", "0.2", "1e2", "0.4", "0.0", "0.0", "--div_type", "l1", "--init_type", "mfgi",
                "--nl", "sine", "--track_weights", "0", "--sphere_init_params", "1.6", "0.1",
                "--loss_weights", "3e3", "1e2", "1e2", "5e1", "1e2", "--grad_clip_norm", "10.0"]
    cwd = os.getcwd()
    os.chdir(script_dir)
    t0 = time.time()
    try:
        runpy.run_path(os.path.join(script_dir, "train_surface_reconstruction.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
    seconds = time.time() - t0
    log = open(os.path.join(logdir, "synthetic_torus", "out.log")).read()
    losses = [float(x) for x in re.findall(r"\] Loss: ([-0-9.eE+]+) =", log)]
    print(json.dumps({"script": "surface_reconstruction/train_surface_reconstruction.py (unmodified)", "module": "digs_b200",
                      "precision": args.precision, "iterations": args.iters, "points_per_iteration": 2 * args.n_points,
                      "seconds_total": seconds, "loss_first": losses[0], "loss_last": losses[-1],
                      "logged_iterations": len(losses), "logdir": logdir}))


if __name__ == "__main__":
    main()
