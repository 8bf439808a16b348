"""TEST INFRASTRUCTURE ONLY.  The reference's *mechanism*, restated in torch (CPU).

Where jet_spec.py restates the mathematics in closed form, this file restates HOW the reference
computes it, so that timing it is a fair CPU baseline ("port") and so that the closed form is
cross-checked by an independent route:

* sine MLP forward as a chain of addmm + sin                models/DiGS.py:122-157
* sqrt-sphere head                                          models/DiGS.py:128-132
* gradient = autograd.grad(out, pts, ones, create_graph)    utils/utils.py:108-117
* Laplacian = one more autograd.grad per axis, trace        models/losses.py:100-106, NaN->0 :107
* loss terms and weighting                                  models/losses.py:109-184
* loss.backward() through the third-order graph             train_surface_reconstruction.py:128

It is a functional restatement (weights passed as a flat list), not the reference's module tree.
Pinned against the real reference via tests/golden (tests/test_oracle_cpu.py).
"""
import torch


def mlp(params, pts, head=None):
    """params: [W0,b0,W1,b1,...]; pts (n, in_dim).  Returns (n,) SDF prediction."""
    h = pts
    n_layers = len(params) // 2
    for i in range(n_layers):
        W, b = params[2 * i], params[2 * i + 1]
        h = torch.addmm(b, h, W.t())
        if i < n_layers - 1:
            h = torch.sin(30.0 * h)
    out = h[:, 0]
    if head is not None:
        radius, scale = head
        out = (torch.sign(out) * torch.sqrt(out.abs() + 1e-8) - radius) * scale
    return out


def spatial_grad(pts, out):
    return torch.autograd.grad(out, pts, torch.ones_like(out), create_graph=True, retain_graph=True)[0]


def laplacian(pts, g):
    d = pts.shape[-1]
    lap = 0
    for k in range(d):
        lap = lap + spatial_grad(pts, g[:, k])[:, k]
    return lap


def jet(params, pts, head=None, want_lap=True):
    """pts (n,d) leaf with requires_grad.  Returns f, grad, lap (graph attached)."""
    f = mlp(params, pts, head)
    g = spatial_grad(pts, f)
    lap = laplacian(pts, g) if want_lap else None
    return f, g, lap


def digs_loss(f_m, g_m, f_n, g_n, lap_n, normals, weights, loss_type, div_type, div_clamp):
    w = list(weights)
    div_loss = torch.zeros((), dtype=f_n.dtype)
    if "div" in loss_type:
        lap = torch.where(lap_n.isnan(), torch.zeros_like(lap_n), lap_n)
        t = lap.abs() if div_type == "l1" else lap.square()
        div_loss = torch.clamp(t, 0.1, div_clamp).mean()
    allg = torch.cat([g_n, g_m], dim=0)
    eik = (allg.norm(2, dim=1) - 1).abs().mean()
    normal_term = None
    if normals is not None:
        normal_term = (1 - torch.nn.functional.cosine_similarity(g_m, normals, dim=-1).abs()).mean()
    sdf = f_m.abs().mean()
    inter = torch.exp(-1e2 * f_n.abs()).mean()
    if loss_type == "siren_wo_n":
        w[2] = 0
    loss = w[0] * sdf + w[1] * inter + w[3] * eik
    if loss_type in ("siren", "siren_w_div"):
        loss = loss + w[2] * normal_term
    if loss_type in ("siren_w_div", "siren_wo_n_w_div"):
        loss = loss + w[4] * div_loss
    return dict(loss=loss, sdf_term=sdf, inter_term=inter, eikonal_term=eik,
                normals_loss=normal_term, div_loss=div_loss)


def step(params, mnfld, nonmnfld, normals, head, weights, loss_type="siren_wo_n_w_div",
         div_type="l1", div_clamp=50.0):
    """One training step's hot path on CPU tensors.  params must require grad.

    Returns (terms, grads list aligned with params, dict of jets)."""
    m = mnfld.detach().clone().requires_grad_(True)
    nm = nonmnfld.detach().clone().requires_grad_(True)
    use_div = "div" in loss_type
    f_m, g_m, _ = jet(params, m, head, want_lap=False)
    f_n, g_n, lap_n = jet(params, nm, head, want_lap=use_div)
    terms = digs_loss(f_m, g_m, f_n, g_n, lap_n, normals, weights, loss_type, div_type, div_clamp)
    grads = torch.autograd.grad(terms["loss"], params, allow_unused=True)
    jets = dict(f_m=f_m.detach(), g_m=g_m.detach(), f_n=f_n.detach(), g_n=g_n.detach(),
                lap_n=None if lap_n is None else lap_n.detach())
    return terms, grads, jets


def grid_values(params, pts, head=None, chunk=100000):
    """Value-only query in 100k chunks like utils/utils.py:343-348."""
    outs = []
    with torch.no_grad():
        for p in torch.split(pts, chunk, dim=0):
            outs.append(mlp(params, p.float(), head))
    return torch.cat(outs)
