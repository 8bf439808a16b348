"""The reference's single-shape training loop (surface_reconstruction/train_surface_reconstruction.py:85-172) on a
synthetic cloud, with digs_b200 dropped in for models.DiGS: same calls, same hyper-parameters (run_surf_recon_exp.sh).

    python examples/train_synthetic.py --iters 500 [--precision bf16x2] [--graph]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digs_b200 as DiGSNet                      # noqa: E402   (the reference script: import models.DiGS as DiGSNet)
from digs_b200 import synth                      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--n_points", type=int, default=15000)
    ap.add_argument("--precision", default="bf16x2")
    ap.add_argument("--graph", action="store_true", help="GraphedStep + FusedClipAdam instead of eager autograd + torch Adam")
    ap.add_argument("--device-sampler", action="store_true", help="with --graph: draw every batch on the device, inside the graph")
    ap.add_argument("--grid_res", type=int, default=128)
    ap.add_argument("--ply", default="", help="write the extracted mesh to this .ply file")
    args = ap.parse_args()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    np.random.seed(0)
    cloud, normals = synth.surface_cloud(100000, "torus", seed=0)
    torch.manual_seed(3627473)
    net = DiGSNet.DiGSNetwork(latent_size=0, in_dim=3, decoder_hidden_dim=256, nl="sine", encoder_type="none",
                              decoder_n_hidden_layers=4, init_type="mfgi", sphere_init_params=[1.6, 0.1],
                              precision=args.precision).to(dev)
    criterion = DiGSNet.DiGSLoss(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren_wo_n_w_div", div_decay="linear",
                                 div_type="l1", div_clamp=50)
    decay = (1e2, 0.2, 1e2, 0.4, 0.0, 0.0)
    if args.graph:
        # the whole iteration -- [sampling,] forward, loss, backward, clip + Adam -- is ONE CUDA graph replay
        opt = DiGSNet.FusedClipAdam(net.parameters(), lr=5e-5, max_norm=10.0, device_step=True)
        sampler = DiGSNet.DeviceSampler(cloud, normals, args.n_points, grid_range=1.1, seed=0) if args.device_sampler else None
        step = DiGSNet.GraphedStep(net, criterion, args.n_points, args.n_points, sampler=sampler, optimizer=opt)
    else:
        opt = torch.optim.Adam(net.parameters(), lr=5e-5, betas=(0.9, 0.999))
    t0 = time.time()
    for it in range(args.iters):
        if not (args.graph and args.device_sampler):
            m, nrm, nm = (torch.from_numpy(a).to(dev) for a in synth.sample_batch(cloud, normals, args.n_points, seed=it))
        if args.graph:
            loss_dict = step.step() if args.device_sampler else step.step(m, nm, nrm)
        else:
            m.requires_grad_()
            nm.requires_grad_()
            net.zero_grad()
            loss_dict, _ = criterion(net(nm, m), m, nm, nrm)
            loss_dict["loss"].backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), 10.)
            opt.step()
        criterion.update_div_weight(it, args.iters, decay)
        if it % 100 == 0 or it == args.iters - 1:
            print("iter %5d  loss %.4f  sdf %.5f  eikonal %.4f  div %.4f  (w_div %.1f)" % (
                it, float(loss_dict["loss"]), float(loss_dict["sdf_term"]), float(loss_dict["eikonal_term"]),
                float(loss_dict["div_loss"].reshape(-1)[0]), criterion.weights[4]))
    torch.cuda.synchronize()
    print("%d iterations in %.2f s" % (args.iters, time.time() - t0))
    axes, _, _ = DiGSNet.grid_axes(args.grid_res, np.array([[-1, 1], [-1, 1], [-1, 1]], dtype=float))
    vol = DiGSNet.grid_query(net, axes)
    inside = float((vol < 0).float().mean())
    print("grid %d^3: SDF in [%.3f, %.3f], %.1f %% of the box inside the surface" % (args.grid_res, float(vol.min()),
                                                                                   float(vol.max()), 100 * inside))
    # test-time path of the reference (test_surface_reconstruction.py:59-61, compute_metrics_srb.py:38-57,75), all on the device
    t1 = time.time()
    mesh = DiGSNet.implicit2mesh(net.decoder, None, args.grid_res, device=dev)["mesh_obj"]
    recon, _ = mesh.sample_surface(min(1000000, 20 * len(cloud)))
    cd, hd, cd_re2gt, cd_gt2re, _, _ = DiGSNet.compute_dists(recon, cloud)
    torch.cuda.synchronize()
    print("mesh: %d vertices, %d faces; Chamfer to the input cloud %.5f (recon->gt %.5f, gt->recon %.5f), Hausdorff %.4f  [%.2f s]"
          % (len(mesh.vertices), len(mesh.faces), cd, cd_gt2re, cd_re2gt, hd, time.time() - t1))
    if args.ply:
        mesh.export(args.ply, vertex_normal=True)


if __name__ == "__main__":
    main()
