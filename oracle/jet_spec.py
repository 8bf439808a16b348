"""TEST INFRASTRUCTURE ONLY.  Closed-form jet oracle (numpy).

Restates, as explicit recurrences, what the reference obtains from autograd:

* the sine MLP                               models/DiGS.py:122-134, 142-151, 154-157
* its sqrt-sphere output head                models/DiGS.py:128-132
* per-point gradient / Hessian trace         utils/utils.py:108-117, models/losses.py:72-76, 100-107
* the DiGS loss terms                        models/losses.py:109-184  (+ eikonal :13-29)
* the reverse pass of all of the above       caller's loss.backward(), train_surface_reconstruction.py:128

Channels per point: value h, Jacobian columns A_k (k<d), Laplacian P.
Everything runs in the dtype of the inputs; float64 is the ground truth the CUDA kernels and
the fp32 reference are both measured against.

Parity pinning: checked against outputs of the real reference (tests/golden/*.npz, produced
by oracle/make_golden.py in the build container) in tests/test_oracle_cpu.py.
"""
import numpy as np

OMEGA = 30.0  # models/DiGS.py:157


def split_params(params):
    """params: list [(W0,b0), ..., (W_last,b_last)] in nn.Linear layout (out,in)."""
    return params[0], params[1:-1], params[-1]


def _activate(z, Z, Q):
    s = np.sin(OMEGA * z)
    c = np.cos(OMEGA * z)
    h = s
    A = [OMEGA * c * Zk for Zk in Z]
    S2 = sum(Zk * Zk for Zk in Z)
    P = OMEGA * c * Q - (OMEGA * OMEGA) * s * S2
    return h, A, P


def forward(params, x, head=None, extra_bias0=None):
    """Forward jet.

    params      list of (W, b) numpy arrays, nn.Linear layout.
    x           (n, d) points (only the first d = x.shape[1] input columns of W0 are used).
    head        None for the identity head, or (radius, scale) for the sqrt-sphere head.
    extra_bias0 optional (n, H) per-point additive bias of layer 0 (latent folded in).
    Returns dict with f (n,), grad (n,d), lap (n,), and everything the adjoint needs.
    """
    (W0, b0), hidden, (wl, bl) = split_params(params)
    n, d = x.shape
    dt = x.dtype
    z = x @ W0[:, :d].T + b0
    if extra_bias0 is not None:
        z = z + extra_bias0
    Z = [np.broadcast_to(W0[:, k], z.shape).astype(dt) for k in range(d)]
    Q = np.zeros_like(z)
    pre = [(z, Z, Q)]
    h, A, P = _activate(z, Z, Q)
    for (W, b) in hidden:
        z = h @ W.T + b
        Z = [Ak @ W.T for Ak in A]
        Q = P @ W.T
        pre.append((z, Z, Q))
        h, A, P = _activate(z, Z, Q)
    w = wl[0]
    u = h @ w + bl[0]
    gu = np.stack([Ak @ w for Ak in A], axis=1)          # (n, d)
    lu = P @ w
    if head is None:
        f, grad, lap = u, gu, lu
    else:
        radius, scale = head
        a = np.abs(u) + 1e-8
        sg = np.sign(u)
        g1 = 0.5 / np.sqrt(a)
        g2 = -0.25 * sg * a ** -1.5
        f = scale * (sg * np.sqrt(a) - radius)
        grad = scale * g1[:, None] * gu
        lap = scale * (g1 * lu + g2 * (gu * gu).sum(1))
    return dict(f=f, grad=grad, lap=lap, u=u, gu=gu, lu=lu, pre=pre, x=x)


def backward(params, fwd, f_bar, g_bar, l_bar, head=None):
    """Adjoint of forward(): returns list of (dW, db) like params, plus d(extra_bias0) (n,H).

    f_bar (n,), g_bar (n,d), l_bar (n,) are dLoss/df, dLoss/dgrad, dLoss/dlap.
    """
    (W0, b0), hidden, (wl, bl) = split_params(params)
    x, u, gu, lu, pre = fwd["x"], fwd["u"], fwd["gu"], fwd["lu"], fwd["pre"]
    n, d = x.shape
    if head is None:
        u_bar, gu_bar, lu_bar = f_bar, g_bar, l_bar
    else:
        radius, scale = head
        a = np.abs(u) + 1e-8
        sg = np.sign(u)
        g1 = 0.5 / np.sqrt(a)
        g2 = -0.25 * sg * a ** -1.5
        g3 = 0.375 * a ** -2.5
        gg = (gu * gu).sum(1)
        u_bar = scale * (g1 * f_bar + g2 * (g_bar * gu).sum(1) + l_bar * (g2 * lu + g3 * gg))
        gu_bar = scale * (g1[:, None] * g_bar + 2.0 * (g2 * l_bar)[:, None] * gu)
        lu_bar = scale * g1 * l_bar
    # last layer
    z, Z, Q = pre[-1]
    h, A, P = _activate(z, Z, Q)
    w = wl[0]
    dw_last = u_bar @ h + sum(gu_bar[:, k] @ A[k] for k in range(d)) + lu_bar @ P
    db_last = np.array([u_bar.sum()], dtype=x.dtype)
    h_bar = u_bar[:, None] * w[None, :]
    A_bar = [gu_bar[:, k][:, None] * w[None, :] for k in range(d)]
    P_bar = lu_bar[:, None] * w[None, :]
    grads_hidden = []
    for li in range(len(hidden), -1, -1):          # sine layers Lh .. 0
        z, Z, Q = pre[li]
        s = np.sin(OMEGA * z)
        c = np.cos(OMEGA * z)
        S2 = sum(Zk * Zk for Zk in Z)
        w2, w3 = OMEGA * OMEGA, OMEGA ** 3
        Q_bar = OMEGA * c * P_bar
        Z_bar = [OMEGA * c * A_bar[k] - 2.0 * w2 * s * Z[k] * P_bar for k in range(d)]
        z_bar = (OMEGA * c * h_bar - w2 * s * sum(Z[k] * A_bar[k] for k in range(d))
                 - P_bar * (w2 * s * Q + w3 * c * S2))
        if li == 0:
            dW0 = np.zeros_like(W0)
            dW0[:, :d] = z_bar.T @ x + np.stack([Z_bar[k].sum(0) for k in range(d)], axis=1)
            db0 = z_bar.sum(0)
            ebias_bar = z_bar
            break
        W, b = hidden[li - 1]
        zp, Zp, Qp = pre[li - 1]
        hp, Ap, Pp = _activate(zp, Zp, Qp)
        dW = z_bar.T @ hp + sum(Z_bar[k].T @ Ap[k] for k in range(d)) + Q_bar.T @ Pp
        db = z_bar.sum(0)
        grads_hidden.append((dW, db))
        h_bar = z_bar @ W
        A_bar = [Z_bar[k] @ W for k in range(d)]
        P_bar = Q_bar @ W
    grads = [(dW0, db0)] + grads_hidden[::-1] + [(dw_last[None, :], db_last)]
    return grads, ebias_bar


# ----------------------------------------------------------------------------------------------
# DiGS loss from jet quantities  (models/losses.py:52-188)
# ----------------------------------------------------------------------------------------------
ADDS_NORMAL = ("siren", "siren_w_div")
USES_DIV = ("siren_w_div", "siren_wo_n_w_div")
IN_SCOPE_LOSS_TYPES = ("siren", "siren_wo_n", "siren_w_div", "siren_wo_n_w_div")


def loss_and_seeds(f_m, g_m, f_n, g_n, lap_n, normals, weights, loss_type="siren_wo_n_w_div",
                   div_type="l1", div_clamp=50.0, latent_reg=None):
    """Loss terms and their adjoint seeds.

    f_m (Nm,), g_m (Nm,d), f_n (Nn,), g_n (Nn,d), lap_n (Nn,) or None, normals (Nm,d) or None.
    Means are over all points (bs*N), exactly like .mean() on the (bs,N) tensors.
    Returns (terms dict, seeds dict(f_m, g_m, f_n, g_n, lap_n)).
    """
    assert loss_type in IN_SCOPE_LOSS_TYPES
    w = list(weights)
    if loss_type == "siren_wo_n":
        w[2] = 0.0                                            # losses.py:153
    Nm, Nn = f_m.shape[0], f_n.shape[0]
    dt = f_n.dtype
    use_div = loss_type in USES_DIV
    # divergence term, losses.py:99-122
    div_loss = dt.type(0.0)
    lap_seed = np.zeros_like(f_n)
    if use_div:
        lap = np.where(np.isnan(lap_n), 0.0, lap_n).astype(dt)           # :107
        nan_mask = np.isnan(lap_n)
        if div_type == "l1":
            t = np.abs(lap)
            dterm = np.sign(lap)
        elif div_type == "l2":
            t = lap * lap
            dterm = 2.0 * lap
        else:
            raise ValueError("div_type")
        div_loss = np.clip(t, 0.1, div_clamp).mean()
        inside = (t >= 0.1) & (t <= div_clamp) & (~nan_mask)             # clamp passes grad on the closed interval
        lap_seed = np.where(inside, dterm, 0.0).astype(dt) * (w[4] / Nn)
    # eikonal, losses.py:13-29,125  (cat order: nonmnfld then mnfld; mean over Nm+Nn)
    nrm_n = np.sqrt((g_n * g_n).sum(1))
    nrm_m = np.sqrt((g_m * g_m).sum(1))
    eik = (np.abs(nrm_n - 1).sum() + np.abs(nrm_m - 1).sum()) / (Nm + Nn)

    def eik_seed(g, nrm):
        safe = np.where(nrm > 0, nrm, 1.0)
        return (w[3] / (Nm + Nn)) * (np.sign(nrm - 1) / safe)[:, None] * g * (nrm > 0)[:, None]

    g_n_seed = eik_seed(g_n, nrm_n)
    g_m_seed = eik_seed(g_m, nrm_m)
    # normal term, losses.py:135  (cosine_similarity eps = 1e-8 clamps each norm)
    normal_term = None
    if normals is not None:
        nn_ = np.sqrt((normals * normals).sum(1))
        gm_c = np.maximum(nrm_m, 1e-8)
        nn_c = np.maximum(nn_, 1e-8)
        cos = (g_m * normals).sum(1) / (gm_c * nn_c)
        normal_term = (1 - np.abs(cos)).mean()
        if loss_type in ADDS_NORMAL:
            # d cos / d g = n/(|g||n|) - cos * g/|g|^2   (norm clamp inactive away from 0)
            dcos = normals / (gm_c * nn_c)[:, None] - (cos / (gm_c * gm_c))[:, None] * g_m
            g_m_seed = g_m_seed + (-w[2] / Nm) * np.sign(cos)[:, None] * dcos
    # sdf / inter, losses.py:138-141
    sdf = np.abs(f_m).mean()
    inter = np.exp(-100.0 * np.abs(f_n)).mean()
    f_m_seed = (w[0] / Nm) * np.sign(f_m)
    f_n_seed = (-100.0 * w[1] / Nn) * np.sign(f_n) * np.exp(-100.0 * np.abs(f_n))
    loss = w[0] * sdf + w[1] * inter + w[3] * eik
    if loss_type in ADDS_NORMAL:
        loss = loss + w[2] * normal_term
    if use_div:
        loss = loss + w[4] * div_loss
    if latent_reg is not None:
        loss = loss + w[5] * latent_reg
    terms = dict(loss=loss, sdf_term=sdf, inter_term=inter, eikonal_term=eik,
                 normals_loss=normal_term, div_loss=div_loss)
    seeds = dict(f_m=f_m_seed.astype(dt), g_m=g_m_seed.astype(dt), f_n=f_n_seed.astype(dt),
                 g_n=g_n_seed.astype(dt), lap_n=lap_seed)
    return terms, seeds


def step(params, mnfld, nonmnfld, normals, head, weights, loss_type="siren_wo_n_w_div",
         div_type="l1", div_clamp=50.0):
    """One full hot-path step: forward jets on both clouds, loss, parameter gradients."""
    fm = forward(params, mnfld, head)
    fn = forward(params, nonmnfld, head)
    terms, seeds = loss_and_seeds(fm["f"], fm["grad"], fn["f"], fn["grad"], fn["lap"], normals,
                                  weights, loss_type, div_type, div_clamp)
    zeros_m = np.zeros_like(fm["f"])
    gm, _ = backward(params, fm, seeds["f_m"], seeds["g_m"], zeros_m, head)
    gn, _ = backward(params, fn, seeds["f_n"], seeds["g_n"], seeds["lap_n"], head)
    grads = [(a[0] + b[0], a[1] + b[1]) for a, b in zip(gm, gn)]
    return dict(terms=terms, grads=grads, fm=fm, fn=fn, seeds=seeds)
