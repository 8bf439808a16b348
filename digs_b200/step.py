"""One training iteration's hot path as TWO kernel launches (digs_step_fused).

Replaces, in one call, what the reference's loops do between fetching a batch and clipping the gradients
(train_surface_reconstruction.py:121-128, train_basic_shape.py:105-132, train_shapespace.py:130-145):

    output_pred = net(nonmnfld_points, mnfld_points)                              models/DiGS.py:43-66
    loss_dict, _ = criterion(output_pred, mnfld_points, nonmnfld_points, normals) models/losses.py:52-188
    loss_dict["loss"].backward()

Every DiGS loss term is a mean of per-point functions, so a persistent kernel can run forward jet -> output head -> loss
terms and adjoint seeds -> reverse jet per point tile without the per-layer state ever leaving the chip's L2
(csrc/step_tc.cu); a second launch (the weight-gradient GEMM over both clouds) finishes the parameter gradients.
The two-call API (DiGSNetwork.forward / DiGSLoss.forward + autograd) stays available and gives the same numbers; this
entry point exists because the two-call shape forces every activation through HBM between the calls.

In scope: the tensor-core precision modes (bf16x2, bf16) and the loss types of the fused loss kernel
(siren, siren_wo_n, siren_w_div, siren_wo_n_w_div with div_type l1 | l2).  Anything else raises -- callers use the
two-call API there (no silent fallback).
"""
import ctypes
import os

import torch

from . import _lib
from .jet import _ptr, _stream
from .losses import IN_SCOPE, _ADDS_NORMAL


def fused_step_supported(net, criterion):
    fc = net.decoder.fc_block
    return (fc.mode() != _lib.MODE_FP32 and criterion.loss_type in IN_SCOPE
            and (not criterion.use_div or criterion.div_type in ("l1", "l2"))
            and not net.compute_manifold_laplacian)


def fused_step(net, criterion, mnfld_points, nonmnfld_points, mnfld_n_gt, latent=None, grads=None, terms=None):
    """Forward jets of both clouds + DiGSLoss + the reverse pass, accumulated into the parameter gradients.

    mnfld_points (bs, Nm, d), nonmnfld_points (bs, Nn, d), mnfld_n_gt (bs, Nm, d): float32 CUDA tensors (coordinates only).
    latent      None, or (bs, latent_size): the auto-decoder latent of each cloud of the batch (train_shapespace.py:122-128
                concatenates it to every point; here it folds into a per-shape bias of layer 0).
    grads       optional list of tensors aligned with net.decoder.fc_block.flat_params() to ACCUMULATE into
                (default: each parameter's .grad, created as zeros when missing).
    terms       optional ZEROED float32[8] device tensor the kernel adds the loss scalars into (slots: loss, sdf, inter,
                eikonal, normal, div, -, -); GraphedStep passes the tail of its flat gradient buffer, so the scalars ride
                through the gradient all-reduce without any copy kernel.
    Returns (loss_dict, output_pred, latent_grad): the 8 loss tensors of models/losses.py:186-188, the dict of
    models/DiGS.py:63-66 (+ the gradient / divergence keys of the drop-in network), and d loss / d latent (or None).
    """
    if not fused_step_supported(net, criterion):
        raise RuntimeError("digs_b200.fused_step: needs a tensor-core precision mode and an in-scope loss type "
                           "(got precision=%r, loss_type=%r, div_type=%r)" % (net.precision, criterion.loss_type,
                                                                              criterion.div_type))
    if mnfld_n_gt is None:
        raise NameError("normal_term is undefined when mnfld_n_gt is None (models/losses.py:131-135,186)")
    fc = net.decoder.fc_block
    d = net.in_dim
    for name, t in (("mnfld_points", mnfld_points), ("nonmnfld_points", nonmnfld_points), ("mnfld_n_gt", mnfld_n_gt)):
        if not t.is_cuda:
            raise RuntimeError("digs_b200.fused_step: %s must be a CUDA tensor -- the hot path has no CPU fallback" % name)
        if t.shape[-1] != d:
            raise RuntimeError("digs_b200.fused_step: %s must have %d coordinate columns" % (name, d))
    dev = mnfld_points.device
    bs, n_m = mnfld_points.shape[0], mnfld_points.shape[1]
    n_n = nonmnfld_points.shape[1]
    xm = mnfld_points.detach().reshape(-1, d).float().contiguous()
    xn = nonmnfld_points.detach().reshape(-1, d).float().contiguous()
    nrm = mnfld_n_gt.detach().reshape(-1, d).float().contiguous()
    params = fc.flat_params()
    if grads is None:
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in params]
    spec = fc.spec()
    pdet = [p.detach() for p in params]
    cnet = spec.c_net(pdet)
    mode = fc.mode()
    want_lap = bool(criterion.use_div)
    lin0 = fc.linears()[0]
    sb = sbg_n = sbg_m = None
    if latent is not None:
        if latent.shape != (bs, lin0.in_features - d):
            raise RuntimeError("digs_b200.fused_step: latent must be (bs, %d)" % (lin0.in_features - d))
        lat = latent.detach().float()
        sb = torch.addmm(lin0.bias.detach(), lat, lin0.weight.detach()[:, d:].t()).contiguous()     # (bs, H)
        sbg_n, sbg_m = torch.zeros_like(sb), torch.zeros_like(sb)
    elif lin0.in_features != d:
        raise RuntimeError("digs_b200.fused_step: this network expects a latent of size %d" % (lin0.in_features - d))
    f_n = torch.empty(bs * n_n, dtype=torch.float32, device=dev)
    g_n = torch.empty(bs * n_n, d, dtype=torch.float32, device=dev)
    l_n = torch.empty(bs * n_n, dtype=torch.float32, device=dev) if want_lap else None
    f_m = torch.empty(bs * n_m, dtype=torch.float32, device=dev)
    g_m = torch.empty(bs * n_m, d, dtype=torch.float32, device=dev)
    if terms is None:
        terms = torch.zeros(8, dtype=torch.float32, device=dev)
    if criterion.loss_type == "siren_wo_n":
        criterion.weights[2] = 0                                      # models/losses.py:153
    cfg = _lib.DigsLossCfg()
    w = (list(criterion.weights) + [0.0] * 6)[:6]
    for i in range(6):
        cfg.weights[i] = float(w[i])
    wdev = getattr(criterion, "device_weights", None)
    cfg.weights_device = None if wdev is None else wdev.data_ptr()
    cfg.add_normal = int(criterion.loss_type in _ADDS_NORMAL)
    cfg.use_div = int(criterion.use_div)
    cfg.div_l2 = int(criterion.div_type == "l2")
    cfg.div_clamp = float(criterion.div_clamp)
    gc = criterion.global_counts
    cfg.Nm_total, cfg.Nn_total = (bs * n_m, bs * n_n) if gc is None else (int(gc[0]), int(gc[1]))
    gr = _lib.DigsGrads()
    for l in range(spec.Lh + 2):
        gr.dW[l] = grads[2 * l].data_ptr()
        gr.db[l] = grads[2 * l + 1].data_ptr()
    L = _lib.lib()
    wbytes = L.digs_step_workspace_bytes(ctypes.byref(cnet), bs * n_n, bs * n_m, int(want_lap), mode)
    ws = torch.empty(max(int(wbytes), 16), dtype=torch.uint8, device=dev)
    if os.environ.get("DIGS_B200_TRACE"):      # diagnostic builds (make DIAG=TRACE): keep the workspace for scripts/overlap_trace.py
        global _trace_ws
        _trace_ws = ws
    with torch.cuda.device(dev):
        rc = L.digs_step_fused(ctypes.byref(cnet), _ptr(xn), bs * n_n, _ptr(xm), bs * n_m, _ptr(nrm), d,
                               _ptr(sb), _ptr(sb), n_n, n_m, int(want_lap), ctypes.byref(cfg),
                               _ptr(f_n), _ptr(g_n), _ptr(l_n), _ptr(f_m), _ptr(g_m), _ptr(terms),
                               ctypes.byref(gr), _ptr(sbg_n), _ptr(sbg_m), _ptr(ws), wbytes, mode, _stream(dev))
    _lib.check(rc, "step_fused")
    loss = terms[0]
    zero = torch.zeros(1, device=dev)
    latent_reg_term, latent_grad, latent_rows = zero, None, None
    if latent is not None:
        # layer-0 latent columns, bias and the latent itself through the per-shape bias (SURVEY 8a "layer 0")
        sbg = sbg_n + sbg_m
        grads[0][:, d:].addmm_(sbg.t(), lat)
        grads[1].add_(sbg.sum(0))
        latent_grad = sbg @ lin0.weight.detach()[:, d:]
        if net.encoder_type == "autodecoder":
            # models/DiGS.py:48-49: mean over all points of ||latent|| = mean over the shapes of the GLOBAL batch (shapes
            # sharded over ranks contribute partial sums that the gradient all-reduce adds up)
            bs_glob = bs if gc is None else max(int(gc[1]) // max(n_n, 1), 1)
            nrm_l = lat.norm(dim=-1)
            latent_reg_term = nrm_l.sum() / bs_glob
            w5 = float(w[5])
            terms[0:1].add_(w5 * latent_reg_term)                      # models/losses.py:183-184 (in place: `terms` may be shared)
            terms[6:7].add_(latent_reg_term)
            loss = terms[0]
            latent_grad = latent_grad + (w5 / bs_glob) * torch.where(nrm_l[:, None] > 0, lat / nrm_l[:, None].clamp_min(1e-30),
                                                                     torch.zeros_like(lat))
            latent_rows = lat[:, None, :].expand(bs, n_n, lat.shape[1])
    div_loss = terms[5] if criterion.use_div else zero
    loss_dict = {"loss": loss, "sdf_term": terms[1], "inter_term": terms[2], "latent_reg_term": latent_reg_term,
                 "eikonal_term": terms[3], "normals_loss": terms[4], "div_loss": div_loss,
                 "curv_loss": torch.zeros((), device=dev)}
    output_pred = {"manifold_pnts_pred": f_m.reshape(bs, n_m), "nonmanifold_pnts_pred": f_n.reshape(bs, n_n),
                   "latent_reg": latent_reg_term if latent is not None and net.encoder_type == "autodecoder" else None,
                   "latent": latent_rows,
                   "manifold_pnts_grad": g_m.reshape(bs, n_m, d), "manifold_pnts_div": None,
                   "nonmanifold_pnts_grad": g_n.reshape(bs, n_n, d),
                   "nonmanifold_pnts_div": None if l_n is None else l_n.reshape(bs, n_n)}
    return loss_dict, output_pred, latent_grad
