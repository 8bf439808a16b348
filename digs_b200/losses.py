"""Drop-in for the reference's DiGSLoss (models/losses.py:40-228).

Same constructor, same forward signature and return value ((dict of 8 tensors, mnfld_grad)), same
shared mutable `weights` list and `update_div_weight` schedule.  Where the reference calls
utils.gradient five times (models/losses.py:72,76,100-103) this class consumes the gradient and
Laplacian that the drop-in DiGSNetwork.forward already produced, evaluates all loss terms and their
adjoint seeds in ONE kernel (digs_loss_fused) and hands the seeds back to the jet's reverse kernel
when the caller runs loss.backward().

Fused path (one kernel): loss_type in {siren, siren_wo_n, siren_w_div, siren_wo_n_w_div}, div_type in {l1, l2} -- every
recipe of the reference's run scripts.  The remaining variants (igr, igr_wo_n, *_w_curv; div_type gt_l1 / gt_l2) are
evaluated by `_composite_terms`: the same jets from the CUDA kernels, the loss algebra as a handful of torch
elementwise ops on them (models/losses.py:79-181), autograd carrying the seeds back into the jet's reverse kernel.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from .jet import _ptr, _stream

IN_SCOPE = ("siren", "siren_wo_n", "siren_w_div", "siren_wo_n_w_div")
COMPOSITE = ("igr", "igr_wo_n", "siren_w_curv", "siren_w_div_w_curv", "siren_wo_n_w_div_w_curv")
_ADDS_NORMAL = ("siren", "siren_w_div")


class _FusedLoss(torch.autograd.Function):
    """terms[6] = (loss, sdf, inter, eikonal, normal, div); only terms[0] carries gradient."""

    @staticmethod
    def forward(ctx, f_m, g_m, f_n, g_n, lap_n, normals, cfg):
        L = _lib.lib()
        dev = f_n.device
        d = g_n.shape[-1]
        nm, nn_ = f_m.numel(), f_n.numel()
        f_m, g_m, f_n, g_n = (t.detach().contiguous() for t in (f_m, g_m, f_n, g_n))
        lap = None if lap_n is None else lap_n.detach().contiguous()
        nrm = None if normals is None else normals.detach().to(torch.float32).contiguous()
        fb_m, gb_m = torch.empty_like(f_m), torch.empty_like(g_m)
        fb_n, gb_n = torch.empty_like(f_n), torch.empty_like(g_n)
        lb_n = torch.empty_like(f_n) if lap is not None else None
        terms = torch.zeros(8, dtype=torch.float32, device=dev)
        w = (ctypes.c_float * 6)(*[float(x) for x in (list(cfg["weights"]) + [0.0] * 6)[:6]])
        Nm_tot, Nn_tot = cfg["global_counts"] if cfg["global_counts"] is not None else (nm, nn_)
        wdev = cfg.get("device_weights")
        if wdev is not None:      # CUDA-graph mode: weights live on the device and are re-read at every replay
            rc = L.digs_loss_fused_dw(_ptr(f_m), _ptr(g_m), nm, _ptr(f_n), _ptr(g_n), _ptr(lap), nn_, _ptr(nrm), d,
                                      _ptr(wdev), int(cfg["add_normal"]), int(cfg["use_div"]), int(cfg["div_l2"]),
                                      float(cfg["div_clamp"]), int(Nm_tot), int(Nn_tot),
                                      _ptr(fb_m), _ptr(gb_m), _ptr(fb_n), _ptr(gb_n), _ptr(lb_n), _ptr(terms), _stream())
        else:
            rc = L.digs_loss_fused(_ptr(f_m), _ptr(g_m), nm, _ptr(f_n), _ptr(g_n), _ptr(lap), nn_, _ptr(nrm), d, w,
                                   int(cfg["add_normal"]), int(cfg["use_div"]), int(cfg["div_l2"]),
                                   float(cfg["div_clamp"]), int(Nm_tot), int(Nn_tot),
                                   _ptr(fb_m), _ptr(gb_m), _ptr(fb_n), _ptr(gb_n), _ptr(lb_n), _ptr(terms), _stream())
        _lib.check(rc, "loss_fused")
        ctx.save_for_backward(fb_m, gb_m, fb_n, gb_n, *([lb_n] if lb_n is not None else []))
        ctx.has_lap = lb_n is not None
        return terms

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, t_bar):
        sv = list(ctx.saved_tensors)
        scaled = torch._foreach_mul(sv, t_bar[0])          # one launch for all seeds
        fb_m, gb_m, fb_n, gb_n = scaled[:4]
        lb_n = scaled[4] if ctx.has_lap else None
        return fb_m, gb_m, fb_n, gb_n, lb_n, None, None


class DiGSLoss(nn.Module):
    def __init__(self, weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren", div_decay="none", div_type="l2",
                 div_clamp=50):
        super().__init__()
        self.weights = weights  # sdf, inter, normal, eikonal, div[, latent]; shared and mutated like the reference
        self.loss_type = loss_type
        self.div_decay = div_decay
        self.div_type = div_type
        self.div_clamp = div_clamp
        self.use_curvs = "curv" in self.loss_type
        self.use_div = "div" in self.loss_type
        self.global_counts = None  # (Nm_total, Nn_total) when the clouds are sharded over ranks

    def forward(self, output_pred, mnfld_points, nonmnfld_points, mnfld_n_gt=None, nonmnfld_dist=None,
                curvatures=None, network_weights=None, nonmnfld_div_gt=None, mnfld_div_gt=None):
        if self.loss_type not in IN_SCOPE + COMPOSITE:
            raise Warning("unrecognized loss type")
        if self.use_div and self.div_type not in ("l1", "l2", "gt_l1", "gt_l2"):
            raise Warning("unsupported divergence type. only suuports l1 and l2")
        if self.loss_type in COMPOSITE or (self.use_div and self.div_type in ("gt_l1", "gt_l2")):
            return self._composite_terms(output_pred, mnfld_points, mnfld_n_gt, curvatures, nonmnfld_div_gt, mnfld_div_gt)
        device = mnfld_points.device
        f_n = output_pred["nonmanifold_pnts_pred"]
        f_m = output_pred["manifold_pnts_pred"]
        latent_reg = output_pred["latent_reg"]
        latent = output_pred["latent"]
        g_n = output_pred.get("nonmanifold_pnts_grad")
        g_m = output_pred.get("manifold_pnts_grad")
        lap_n = output_pred.get("nonmanifold_pnts_div")
        if f_m is None:
            raise TypeError("DiGSLoss needs manifold predictions (models/losses.py:138)")
        if g_n is None or g_m is None:
            raise RuntimeError("digs_b200.DiGSLoss: output_pred carries no jet -- it must come from "
                               "digs_b200.DiGSNetwork.forward on points with requires_grad=True")
        if self.use_div and lap_n is None:
            raise RuntimeError("digs_b200.DiGSLoss: divergence loss needs DiGSNetwork.compute_laplacian=True")
        if mnfld_n_gt is None:
            raise NameError("normal_term is undefined when mnfld_n_gt is None (models/losses.py:131-135,186)")
        if self.loss_type == "siren_wo_n":
            self.weights[2] = 0                                      # models/losses.py:153
        cfg = dict(weights=self.weights, add_normal=self.loss_type in _ADDS_NORMAL, use_div=self.use_div,
                   div_l2=self.div_type == "l2", div_clamp=self.div_clamp, global_counts=self.global_counts,
                   device_weights=getattr(self, "device_weights", None))
        terms = _FusedLoss.apply(f_m, g_m, f_n, g_n, lap_n if self.use_div else None, mnfld_n_gt, cfg)
        loss = terms[0]
        det = terms.detach()
        div_loss = det[5] if self.use_div else torch.zeros(1, device=device)
        if latent_reg is not None:
            latent_reg_term = latent_reg.mean()                       # models/losses.py:31-38
        else:
            latent_reg_term = torch.zeros(1, device=device)
        if latent is not None and latent_reg is not None:
            loss = loss + self.weights[5] * latent_reg_term          # models/losses.py:183-184
        curv = torch.zeros((), device=device)
        return {"loss": loss, "sdf_term": det[1], "inter_term": det[2], "latent_reg_term": latent_reg_term,
                "eikonal_term": det[3], "normals_loss": det[4], "div_loss": div_loss, "curv_loss": curv}, g_m

    def _composite_terms(self, output_pred, mnfld_points, mnfld_n_gt, curvatures, nonmnfld_div_gt, mnfld_div_gt):
        """The loss variants outside the fused kernel, on the jets the CUDA path produced (models/losses.py:79-181)."""
        if self.global_counts is not None:
            raise NotImplementedError("digs_b200: point-sharded means are implemented by the fused loss kernel only")
        device = mnfld_points.device
        f_n, f_m = output_pred["nonmanifold_pnts_pred"], output_pred["manifold_pnts_pred"]
        g_n, g_m = output_pred.get("nonmanifold_pnts_grad"), output_pred.get("manifold_pnts_grad")
        lap_n, lap_m = output_pred.get("nonmanifold_pnts_div"), output_pred.get("manifold_pnts_div")
        latent_reg, latent = output_pred["latent_reg"], output_pred["latent"]
        if f_m is None:
            raise TypeError("DiGSLoss needs manifold predictions (models/losses.py:138)")
        if g_n is None or g_m is None:
            raise RuntimeError("digs_b200.DiGSLoss: output_pred carries no jet -- it must come from "
                               "digs_b200.DiGSNetwork.forward on points with requires_grad=True")
        zero = torch.zeros(1, device=device)
        lo, hi = 0.1, self.div_clamp
        curv_term, div_loss = zero, zero
        if self.use_curvs:
            if curvatures is None:
                raise Warning(" loss type requires curvatuers but none were provided.")
            if lap_m is None:
                raise RuntimeError("digs_b200.DiGSLoss: curvature losses need DiGSNetwork.compute_manifold_laplacian = True")
            mean_curv = curvatures.sum(dim=-1)
            if self.div_type == "l2":
                curv_term = (lap_m.square() - mean_curv.square()).clamp(lo, hi)
            elif self.div_type == "l1":
                curv_term = (lap_m.abs() - mean_curv.abs()).clamp(lo, hi)
        if self.use_div:
            if lap_n is None:
                raise RuntimeError("digs_b200.DiGSLoss: divergence loss needs DiGSNetwork.compute_laplacian=True")
            div_n = torch.where(lap_n.isnan(), torch.zeros_like(lap_n), lap_n)               # models/losses.py:107
            if self.div_type == "l2":
                div_term = div_n.square().clamp(lo, hi)
            elif self.div_type == "l1":
                div_term = div_n.abs().clamp(lo, hi)
            else:
                if lap_m is None:      # the reference only defines mnfld_divergence inside its curvature branch
                    raise NameError("mnfld_divergence is undefined without a curvature loss type (models/losses.py:117-121)")
                if self.div_type == "gt_l2":
                    div_term = (div_n - nonmnfld_div_gt).square() + (lap_m - mnfld_div_gt).square()
                else:
                    div_term = (div_n.abs() - nonmnfld_div_gt.abs()).abs() + (lap_m.abs() - mnfld_div_gt.abs()).abs()
            div_loss = div_term.mean()
        eikonal = (torch.cat([g_n, g_m], dim=-2).norm(2, dim=2) - 1).abs().mean()             # models/losses.py:12-29
        if mnfld_n_gt is None:
            raise NameError("normal_term is undefined when mnfld_n_gt is None (models/losses.py:131-135,186)")
        if "igr" in self.loss_type:
            normal = (g_m - mnfld_n_gt).abs().norm(2, dim=1).mean()                           # norm over the POINT axis, as written
        else:
            normal = (1 - torch.nn.functional.cosine_similarity(g_m, mnfld_n_gt, dim=-1).abs()).mean()
        sdf = f_m.abs().mean()
        inter = torch.exp(-1e2 * f_n.abs()).mean()
        w = self.weights
        if self.loss_type in ("siren_wo_n", "igr_wo_n"):
            w[2] = 0                                                                          # models/losses.py:153,161
        if self.loss_type in ("igr", "igr_wo_n"):
            w[1] = 0                                                                          # models/losses.py:156,160
        loss = w[0] * sdf + w[3] * eikonal
        if "igr" not in self.loss_type:
            loss = loss + w[1] * inter
        if self.loss_type in ("siren", "igr", "siren_w_div", "siren_w_curv", "siren_w_div_w_curv"):
            loss = loss + w[2] * normal
        if self.use_div:
            loss = loss + w[4] * div_loss
        if self.use_curvs:
            loss = loss + w[4] * curv_term.mean()
        latent_reg_term = latent_reg.mean() if latent_reg is not None else zero
        if latent is not None and latent_reg is not None:
            loss = loss + w[5] * latent_reg_term
        return {"loss": loss, "sdf_term": sdf, "inter_term": inter, "latent_reg_term": latent_reg_term,
                "eikonal_term": eikonal, "normals_loss": normal, "div_loss": div_loss,
                "curv_loss": curv_term.mean()}, g_m

    def update_div_weight(self, current_iteration, n_iterations, params=None):
        """Piece-wise schedule for weights[4] (models/losses.py:190-228).

        params = (w_start, [fraction, w]*, w_end): knots at fractions (0, ..., 1) of training."""
        if not hasattr(self, "decay_params_list"):
            assert len(params) >= 2, params
            middle = params[1:-1]
            assert len(middle) % 2 == 0
            knot_w = [params[0], *middle[1::2], params[-1]]
            knot_t = [0, *middle[::2], 1]
            self.decay_params_list = list(zip(knot_w, knot_t))
        frac = current_iteration / n_iterations
        w_end, t_end = min((kn for kn in self.decay_params_list if kn[1] >= frac), key=lambda kn: kn[1])
        w_beg, t_beg = max((kn for kn in self.decay_params_list if kn[1] <= frac), key=lambda kn: kn[1])
        before = current_iteration < t_beg * n_iterations
        inside = (not before) and current_iteration < t_end * n_iterations
        if self.div_decay == "linear":
            if before:
                self.weights[4] = w_beg
            elif inside:
                self.weights[4] = w_beg + (w_end - w_beg) * (frac - t_beg) / (t_end - t_beg)
            else:
                self.weights[4] = w_end
        elif self.div_decay == "quintic":
            if before:
                self.weights[4] = w_beg
            elif inside:
                self.weights[4] = w_beg + (w_end - w_beg) * (1 - (1 - (frac - t_beg) / (t_end - t_beg)) ** 5)
            else:
                self.weights[4] = w_end
        elif self.div_decay == "step":
            self.weights[4] = w_beg if before else w_end
        elif self.div_decay == "none":
            pass
        else:
            raise Warning("unsupported div decay value")
