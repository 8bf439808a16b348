"""Drop-in for the reference's DiGSNetwork / Decoder / FCBlock (models/DiGS.py:20-134).

Same constructor signature, same module tree (decoder.fc_block.net[i][0]) and therefore the same
state_dict keys (decoder.fc_block.net.{i}.0.weight/bias, models/DiGS.py:88-99), same forward
signature and the same four keys in the returned dict (models/DiGS.py:63-66).  The arithmetic is
not PyTorch: forward launches the fused jet kernels of libdigs_b200.so and additionally returns
the spatial gradient and Laplacian that the reference's DiGSLoss would have obtained from
autograd (models/losses.py:72-76,100-106), under extra keys that the drop-in DiGSLoss consumes.

Initialisation schemes restate models/DiGS.py:161-253 with the same sequence of RNG draws, so a
seeded construction yields bit-identical parameters to the reference (tests/test_host_cpu.py).
"""
import math

import torch
import torch.nn as nn

from . import _lib
from . import jet as _jet
from .jet import JetFunction, JetPairFunction, NetSpec, value_forward_raw

OMEGA = 30.0
_MFGI_PERIODS = (1, 30)           # models/DiGS.py:218
_MFGI_PORTIONS = (0.25, 0.75)     # models/DiGS.py:220


class Sine(nn.Module):
    """sin(30 x) -- models/DiGS.py:154-157.  Only used when the module tree is run as plain PyTorch
    by a caller that bypasses the fused path (never by this package)."""

    def forward(self, x):
        return torch.sin(OMEGA * x)


# ---------------------------------------------------------------------------------------------
# initialisation (table driven; each entry draws random numbers in the reference's order)
# ---------------------------------------------------------------------------------------------
def _uniform_symmetric_(t, bound):
    t.uniform_(-bound, bound)


@torch.no_grad()
def _init_siren(layers):
    # models/DiGS.py:161-173: all weights U(+-sqrt(6/in)/30), then layer 0 U(+-1/in); biases untouched
    for lin in layers:
        _uniform_symmetric_(lin.weight, math.sqrt(6 / lin.in_features) / 30)
    _uniform_symmetric_(layers[0].weight, 1 / layers[0].in_features)


@torch.no_grad()
def _geometric_base_(lin):
    # models/DiGS.py:177-194
    out = lin.out_features
    _uniform_symmetric_(lin.weight, math.sqrt(3 / out))
    _uniform_symmetric_(lin.bias, 1 / (out * 1000))
    lin.weight.data /= 30
    lin.bias.data /= 30


@torch.no_grad()
def _geometric_tail_(layers):
    # models/DiGS.py:197-214: second-last ~ (pi/2) I, last ~ -1 with bias +H
    pen, last = layers[-2], layers[-1]
    n = pen.out_features
    if tuple(pen.weight.shape) != (n, n):
        raise AssertionError("second-last layer must be square")
    pen.weight.data = 0.5 * math.pi * torch.eye(n) + 0.001 * torch.randn(n, n)
    pen.bias.data = 0.5 * math.pi * torch.ones(n) + 0.001 * torch.randn(n)
    pen.weight.data /= 30
    pen.bias.data /= 30
    k = last.in_features
    if tuple(last.weight.shape) != (1, k) or tuple(last.bias.shape) != (1,):
        raise AssertionError("last layer must be (1, H)")
    last.weight.data = -1 * torch.ones(1, k) + 0.00001 * torch.randn(k)
    last.bias.data = torch.zeros(1) + k


@torch.no_grad()
def _init_geometric_sine(layers):
    for lin in layers:
        _geometric_base_(lin)
    _geometric_base_(layers[0])          # first_layer_geom_sine_init draws again, models/DiGS.py:186-194
    _geometric_tail_(layers)


@torch.no_grad()
def _init_mfgi(layers):
    for lin in layers:
        _geometric_base_(lin)
    # first layer: rows split 25 % / 75 % over periods (1, 30), models/DiGS.py:222-239
    first = layers[0]
    n_in, n_out = first.in_features, first.out_features
    counts = [int(p * n_out) for p in _MFGI_PORTIONS]
    if sum(counts) != n_out:
        raise AssertionError("MFGI split must cover the layer width")
    blocks = []
    for period, cnt in zip(_MFGI_PERIODS, counts):
        s = 30 / period
        blocks.append(torch.zeros(cnt, n_in).uniform_(-math.sqrt(3 / n_in) / s, math.sqrt(3 / n_in) / s))
    first.weight.data = torch.cat(blocks, dim=0)
    # second layer: tiny everywhere except the block fed by the low-frequency rows, models/DiGS.py:241-253
    second = layers[1]
    m = second.in_features
    if tuple(second.weight.shape) != (m, m):
        raise AssertionError("second layer must be square")
    k = int(_MFGI_PORTIONS[0] * m)
    w = torch.zeros(m, m).uniform_(-math.sqrt(3 / m), math.sqrt(3 / m) / 30) * 0.0005
    w[:k, :k] = torch.zeros(k, k).uniform_(-math.sqrt(3 / m) / 30, math.sqrt(3 / m) / 30)
    second.weight.data = w
    _geometric_tail_(layers)


_INITS = {"siren": _init_siren, "geometric_sine": _init_geometric_sine, "mfgi": _init_mfgi}
_SQRT_HEAD_INITS = ("mfgi", "geometric_sine")     # models/DiGS.py:128


class FCBlock(nn.Module):
    """in -> H, Lh x (H -> H), H -> 1 with sin(30.) after all but the last (models/DiGS.py:69-134)."""

    def __init__(self, in_features, out_features, num_hidden_layers, hidden_features, outermost_linear=True,
                 nonlinearity="sine", init_type="siren", sphere_init_params=(1.6, 1.0), spatial_dim=3):
        super().__init__()
        if nonlinearity != "sine":
            raise NotImplementedError("digs_b200 implements the sine network only (nl=%r); use the reference "
                                      "module for other nonlinearities" % (nonlinearity,))
        if init_type not in _INITS:
            raise NotImplementedError("digs_b200: init_type %r is outside the fused path" % (init_type,))
        if out_features != 1 or not outermost_linear:
            raise NotImplementedError("digs_b200: scalar output with a linear last layer only")
        self.init_type = init_type
        self.sphere_init_params = list(sphere_init_params)
        self.spatial_dim = spatial_dim
        act = Sine()
        blocks = [nn.Sequential(nn.Linear(in_features, hidden_features), act)]
        for _ in range(num_hidden_layers):
            blocks.append(nn.Sequential(nn.Linear(hidden_features, hidden_features), act))
        blocks.append(nn.Sequential(nn.Linear(hidden_features, out_features)))
        self.net = nn.Sequential(*blocks)
        _INITS[init_type]([blk[0] for blk in self.net])
        self.precision = "fp32"
        # When True and every parameter already has a contiguous fp32 .grad (e.g. digs_b200.dist.FlatGrads views), the
        # reverse kernels add into those buffers directly and autograd receives None for the parameters (the same
        # contract as fused weight-gradient accumulation in large-model trainers).  Off by default: plain autograd.
        self.accumulate_grads_in_place = False

    # -- description handed to the kernels -------------------------------------------------------
    def linears(self):
        return [blk[0] for blk in self.net]

    def flat_params(self):
        out = []
        for lin in self.linears():
            out += [lin.weight, lin.bias]
        return out

    def spec(self):
        lins = self.linears()
        head = _lib.HEAD_SQRT if self.init_type in _SQRT_HEAD_INITS else _lib.HEAD_IDENTITY
        radius, scale = self.sphere_init_params
        spec = NetSpec(self.spatial_dim, lins[0].out_features, len(lins) - 2, head, radius, scale,
                       lins[0].in_features)
        spec.accumulate_in_place = self.accumulate_grads_in_place
        spec.prepared = self._prepared_operands(spec)
        return spec

    def invalidate_prepared(self):
        """Forget the cached tensor-core operand copies (e.g. before a CUDA-graph capture, so that the preparation
        launch is recorded inside the graph and replayed with it)."""
        self._prep_key = None

    def _prepared_operands(self, spec):
        """bf16 hi/lo copies of 30 W and (30 W)^T for the tcgen05 kernels, rebuilt only when a parameter changed
        (tensor versions + jet.PARAM_EPOCH); one launch instead of one per forward / backward / query call."""
        mode = self.mode()
        if mode == _lib.MODE_FP32:
            return None
        params = self.flat_params()
        if not params[0].is_cuda:
            return None
        key = (mode, _jet.PARAM_EPOCH[0], tuple((p.data_ptr(), p._version) for p in params))
        if key != getattr(self, "_prep_key", None):
            self._prep_buf = _jet.prepare_weights(spec, [p.detach() for p in params], mode, getattr(self, "_prep_buf", None))
            self._prep_key = key
        return self._prep_buf

    def mode(self):
        return _lib.MODES[self.precision]

    def _split_latent(self, coords, pts_per_shape):
        """coords (n, d+l).  Returns (shape_bias (n_shapes,H) | None, pts_per_shape)."""
        d = self.spatial_dim
        lat_dim = coords.shape[1] - d
        if lat_dim == 0:
            return None, 0
        if lat_dim != self.linears()[0].in_features - d:
            raise RuntimeError("digs_b200: input width %d does not match the first layer" % coords.shape[1])
        if pts_per_shape is None:
            # decoder(points) called directly (utils/utils.py:345-348 repeats ONE latent over the chunk): rows may
            # in principle carry different latents -- BatchLinear would use each row's own (models/DiGS.py:142-151).
            # One device comparison decides between the per-shape fold and a per-row bias (never a wrong latent).
            full = coords[:, d:]
            if coords.shape[0] > 1 and not bool((full == full[:1]).all()):
                pts_per_shape = 1
            else:
                pts_per_shape = coords.shape[0]
        # the latent is constant over the points of a shape (train_shapespace.py:124-128,
        # utils/utils.py:345-346), so it folds into a per-shape bias of layer 0
        lat = coords[::pts_per_shape, d:]
        lin0 = self.linears()[0]
        return torch.addmm(lin0.bias, lat, lin0.weight[:, d:].t()).contiguous(), pts_per_shape

    def jet(self, coords, want_lap=True, pts_per_shape=None, attach_points=False):
        """coords (n, d+l) fp32 CUDA.  Returns f (n), grad (n,d), lap (n)|None with autograd to the params."""
        coords = coords.contiguous()
        sb, pps = self._split_latent(coords, pts_per_shape)
        pts = coords if attach_points else coords.detach()
        f, g, lap = JetFunction.apply(pts, sb, self.spec(), pps, bool(want_lap), self.mode(), *self.flat_params())
        return f, g, (lap if want_lap else None)

    def forward(self, coords, params=None, **kwargs):
        """decoder(points) -> (N, 1), models/DiGS.py:122-134."""
        if params is not None:
            raise NotImplementedError("digs_b200: external `params=` (hyper-network weights) are outside the fused path")
        lead = coords.shape[:-1]
        flat = coords.reshape(-1, coords.shape[-1])
        if not flat.is_cuda:
            raise RuntimeError("digs_b200: points must be CUDA tensors -- the hot path has no CPU fallback")
        flat = flat.float().contiguous()
        needs_graph = torch.is_grad_enabled() and (flat.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_graph:
            f, _, _ = self.jet(flat, want_lap=False, attach_points=flat.requires_grad)
        else:
            sb, pps = self._split_latent(flat, None)
            with torch.no_grad():
                f = value_forward_raw(self.spec(), [p.detach() for p in self.flat_params()], flat, sb, pps, self.mode())
        return f.reshape(*lead, 1)


class Decoder(nn.Module):
    def forward(self, *args, **kwargs):
        return self.fc_block(*args, **kwargs)


class DiGSNetwork(nn.Module):
    """Same signature as models/DiGS.py:25-66."""

    def __init__(self, latent_size, in_dim=3, decoder_hidden_dim=256, nl="sine", encoder_type=None,
                 decoder_n_hidden_layers=8, init_type="siren", sphere_init_params=[1.6, 1.0], precision=None):
        super().__init__()
        if precision is None:       # callers written for the reference cannot pass it: DIGS_B200_PRECISION selects the mode
            import os
            precision = os.environ.get("DIGS_B200_PRECISION", "fp32")
        if precision not in _lib.MODES:
            raise ValueError("precision must be one of %s" % (sorted(_lib.MODES),))
        self.encoder_type = encoder_type
        self.init_type = init_type
        if encoder_type == "autodecoder":
            pass
        elif encoder_type == "none":
            latent_size = 0
        else:
            raise ValueError("unsupported encoder type")
        self.in_dim = in_dim
        self.latent_size = latent_size
        self.decoder = Decoder()
        self.decoder.fc_block = FCBlock(in_dim + latent_size, 1, num_hidden_layers=decoder_n_hidden_layers,
                                        hidden_features=decoder_hidden_dim, outermost_linear=True, nonlinearity=nl,
                                        init_type=init_type, sphere_init_params=sphere_init_params,
                                        spatial_dim=in_dim)
        self.decoder.fc_block.precision = precision
        self.compute_laplacian = True     # the non-manifold cloud carries the Laplacian channel
        # the curvature losses (models/losses.py:79-96) also need the Laplacian on the surface cloud
        self.compute_manifold_laplacian = False

    @property
    def precision(self):
        return self.decoder.fc_block.precision

    @precision.setter
    def precision(self, value):
        if value not in _lib.MODES:
            raise ValueError("precision must be one of %s" % (sorted(_lib.MODES),))
        self.decoder.fc_block.precision = value

    def _cloud(self, pts, want_lap):
        bs, npts, width = pts.shape
        if not pts.is_cuda:
            raise RuntimeError("digs_b200: points must be CUDA tensors -- the hot path has no CPU fallback")
        flat = pts.reshape(-1, width).float()
        fc = self.decoder.fc_block
        with_jet = torch.is_grad_enabled() and pts.requires_grad
        if with_jet:
            f, g, lap = fc.jet(flat, want_lap=want_lap, pts_per_shape=npts)
            return (f.reshape(bs, npts), g.reshape(bs, npts, self.in_dim),
                    None if lap is None else lap.reshape(bs, npts))
        if torch.is_grad_enabled() and any(p.requires_grad for p in fc.parameters()):
            f, _, _ = fc.jet(flat, want_lap=False, pts_per_shape=npts)      # value with autograd to the parameters
            return f.reshape(bs, npts), None, None
        flat = flat.contiguous()
        sb, pps = fc._split_latent(flat, npts)                               # each cloud of the batch has its own latent
        with torch.no_grad():
            f = value_forward_raw(fc.spec(), [p.detach() for p in fc.flat_params()], flat, sb, pps, fc.mode())
        return f.reshape(bs, npts), None, None

    def forward(self, non_mnfld_pnts, mnfld_pnts=None):
        latent = latent_reg = None
        m_pred = m_grad = None
        if mnfld_pnts is not None and self.encoder_type == "autodecoder":
            latent = non_mnfld_pnts[:, :, 3:]                          # models/DiGS.py:48
            latent_reg = latent.norm(dim=-1).mean()
        m_lap = None
        paired = (mnfld_pnts is not None and torch.is_grad_enabled() and mnfld_pnts.requires_grad
                  and non_mnfld_pnts.requires_grad and mnfld_pnts.is_cuda and non_mnfld_pnts.is_cuda
                  and not self.compute_manifold_laplacian)
        if paired:
            # both clouds as one autograd node, their kernels overlapped on two streams (JetPairFunction)
            fc = self.decoder.fc_block
            bs, n_n, n_m = non_mnfld_pnts.shape[0], non_mnfld_pnts.shape[1], mnfld_pnts.shape[1]
            flat_n = non_mnfld_pnts.reshape(-1, non_mnfld_pnts.shape[-1]).float().contiguous()
            flat_m = mnfld_pnts.reshape(-1, mnfld_pnts.shape[-1]).float().contiguous()
            sb_n, pps_n = fc._split_latent(flat_n, n_n)
            sb_m, pps_m = fc._split_latent(flat_m, n_m)
            fn, gn, ln, fm, gm = JetPairFunction.apply(flat_n.detach(), flat_m.detach(), sb_n, sb_m, fc.spec(), pps_n, pps_m,
                                                       bool(self.compute_laplacian), fc.mode(), *fc.flat_params())
            n_pred, n_grad = fn.reshape(bs, n_n), gn.reshape(bs, n_n, self.in_dim)
            n_lap = ln.reshape(bs, n_n) if self.compute_laplacian else None
            m_pred, m_grad = fm.reshape(bs, n_m), gm.reshape(bs, n_m, self.in_dim)
        else:
            if mnfld_pnts is not None:
                m_pred, m_grad, m_lap = self._cloud(mnfld_pnts, want_lap=self.compute_manifold_laplacian)
            n_pred, n_grad, n_lap = self._cloud(non_mnfld_pnts, want_lap=self.compute_laplacian)
        return {"manifold_pnts_pred": m_pred,
                "nonmanifold_pnts_pred": n_pred,
                "latent_reg": latent_reg,
                "latent": latent,
                # extra keys (a superset is call-compatible; the reference's callers read only the four above)
                "manifold_pnts_grad": m_grad,
                "manifold_pnts_div": m_lap,
                "nonmanifold_pnts_grad": n_grad,
                "nonmanifold_pnts_div": n_lap}
