"""Python operator layer over the C-ABI: forward jet / reverse pass as one torch.autograd.Function.

Replaces, for one point cloud, what the reference builds out of
  FCBlock.forward (models/DiGS.py:122-134) + utils.gradient x(1+d) (utils/utils.py:108-117,
  models/losses.py:72-76,100-106) in the forward direction and the matching part of
  loss.backward() (train_surface_reconstruction.py:128) in the reverse direction.

PyTorch is used for device memory (caching allocator), the current stream and autograd plumbing
only; all arithmetic happens in libdigs_b200.so.  CPU tensors are rejected: there is no fallback.
"""
import ctypes
import torch

from . import _lib


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device=None):
    """Current stream of `device` (default: the current device).  Kernel launches are wrapped in
    torch.cuda.device(tensor.device) by the callers, so the stream and the launch device always agree."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("digs_b200: %s must be a CUDA tensor -- the hot path has no CPU fallback" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("digs_b200: %s must be float32 (got %s)" % (name, t.dtype))


class NetSpec:
    """Static description of the MLP (everything in DigsNet except the pointers)."""

    def __init__(self, d, H, Lh, head, radius, scale, w0_stride):
        self.d, self.H, self.Lh, self.head = int(d), int(H), int(Lh), int(head)
        self.radius, self.scale, self.w0_stride = float(radius), float(scale), int(w0_stride)
        self.prepared = None        # optional uint8 tensor filled by digs_prepare_weights for the current parameter values

    def c_net(self, params):
        """params: flat list [W0, b0, W1, b1, ...] of contiguous fp32 CUDA tensors."""
        n_layers = self.Lh + 2
        if len(params) != 2 * n_layers:
            raise RuntimeError("digs_b200: expected %d parameter tensors, got %d" % (2 * n_layers, len(params)))
        net = _lib.DigsNet()
        net.d, net.H, net.Lh, net.head = self.d, self.H, self.Lh, self.head
        net.radius, net.scale, net.w0_stride = self.radius, self.scale, self.w0_stride
        for l in range(n_layers):
            W, b = params[2 * l], params[2 * l + 1]
            _require_cuda(W, "weight %d" % l)
            _require_cuda(b, "bias %d" % l)
            if not (W.is_contiguous() and b.is_contiguous()):
                raise RuntimeError("digs_b200: parameters must be contiguous")
            net.W[l] = W.data_ptr()
            net.b[l] = b.data_ptr()
        net.prepared = None if self.prepared is None else self.prepared.data_ptr()
        return net


# Bumped by code that changes parameter VALUES behind autograd's back (raw kernels writing through data pointers, e.g.
# FusedClipAdam): tensor._version does not see those writes, and the prepared-operand cache keys on both.
PARAM_EPOCH = [0]


def bump_param_epoch():
    PARAM_EPOCH[0] += 1


def prepare_weights(spec, params, mode, buf=None):
    """Tensor-core operand copies of the weights (digs_prepare_weights); returns the buffer (None in fp32 mode)."""
    L = _lib.lib()
    spec_prepared, spec.prepared = spec.prepared, None
    net = spec.c_net(params)
    spec.prepared = spec_prepared
    nbytes = L.digs_prepared_bytes(ctypes.byref(net), mode)
    if nbytes == 0:
        return None
    dev = params[0].device
    if buf is None or buf.numel() < nbytes or buf.device != dev:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.digs_prepare_weights(ctypes.byref(net), _ptr(buf), buf.numel(), mode, _stream(dev))
    _lib.check(rc, "prepare_weights")
    return buf


def _scratch(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def jet_forward_raw(spec, params, xyz, shape_bias, pts_per_shape, want_lap, save, mode):
    """xyz (n, stride>=d) fp32 CUDA contiguous.  Returns f (n), grad (n,d), lap (n)|None, saved|None."""
    L = _lib.lib()
    _require_cuda(xyz, "points")
    n, stride = xyz.shape
    dev = xyz.device
    net = spec.c_net(params)
    f = torch.empty(n, dtype=torch.float32, device=dev)
    g = torch.empty(n, spec.d, dtype=torch.float32, device=dev)
    lap = torch.empty(n, dtype=torch.float32, device=dev) if want_lap else None
    saved = None
    sbytes = 0
    if save:
        sbytes = L.digs_saved_bytes(ctypes.byref(net), n, int(want_lap), mode)
        saved = _scratch(sbytes, dev)
    wbytes = L.digs_forward_workspace_bytes(ctypes.byref(net), n, int(want_lap), mode)
    if n > 0 and wbytes == 0:
        _lib.check(-1, "forward_workspace_bytes")
    ws = _scratch(wbytes, dev)
    with torch.cuda.device(dev):
        rc = L.digs_jet_forward(ctypes.byref(net), _ptr(xyz), n, stride, _ptr(shape_bias), int(pts_per_shape),
                                int(want_lap), _ptr(f), _ptr(g), _ptr(lap), _ptr(saved), sbytes, _ptr(ws), wbytes,
                                mode, _stream(dev))
    _lib.check(rc, "jet_forward")
    return f, g, lap, saved


def jet_backward_raw(spec, params, xyz, shape_bias, pts_per_shape, want_lap, f_bar, g_bar, l_bar, saved, mode,
                     grad_out=None):
    """Returns (list of parameter gradients aligned with params, shape_bias_grad|None).

    grad_out: optional list of preallocated fp32 tensors to ACCUMULATE into (flat-buffer views)."""
    L = _lib.lib()
    n, stride = xyz.shape
    dev = xyz.device
    net = spec.c_net(params)
    if grad_out is None:
        grad_out = [torch.zeros_like(p) for p in params]
    gr = _lib.DigsGrads()
    for l in range(spec.Lh + 2):
        gr.dW[l] = grad_out[2 * l].data_ptr()
        gr.db[l] = grad_out[2 * l + 1].data_ptr()
    sb_grad = torch.zeros_like(shape_bias) if shape_bias is not None else None
    wbytes = L.digs_backward_workspace_bytes(ctypes.byref(net), n, int(want_lap), mode)
    ws = _scratch(wbytes, dev)
    for t in (f_bar, g_bar, l_bar):
        if t is not None:
            _require_cuda(t, "adjoint seed")
    with torch.cuda.device(dev):
        rc = L.digs_jet_backward(ctypes.byref(net), _ptr(xyz), n, stride, _ptr(shape_bias), int(pts_per_shape),
                                 int(want_lap), _ptr(f_bar), _ptr(g_bar), _ptr(l_bar), _ptr(saved), saved.numel(),
                                 _ptr(ws), wbytes, ctypes.byref(gr), _ptr(sb_grad), mode, _stream(dev))
    _lib.check(rc, "jet_backward")
    return grad_out, sb_grad


def value_forward_raw(spec, params, xyz, shape_bias, pts_per_shape, mode):
    L = _lib.lib()
    _require_cuda(xyz, "points")
    n, stride = xyz.shape
    net = spec.c_net(params)
    f = torch.empty(n, dtype=torch.float32, device=xyz.device)
    wbytes = L.digs_value_workspace_bytes(ctypes.byref(net), n, mode)
    ws = _scratch(wbytes, xyz.device)
    with torch.cuda.device(xyz.device):
        rc = L.digs_value_forward(ctypes.byref(net), _ptr(xyz), n, stride, _ptr(shape_bias), int(pts_per_shape),
                                  _ptr(f), _ptr(ws), wbytes, mode, _stream(xyz.device))
    _lib.check(rc, "value_forward")
    return f


class JetFunction(torch.autograd.Function):
    """(f, grad f, lap f) = jet(points; params); differentiable once w.r.t. params and shape_bias.

    The gradient w.r.t. the points themselves is provided at first order only (f_bar * grad f); asking
    for it while grad/lap outputs also carry adjoints would need Hessian-vector products and raises.
    """

    @staticmethod
    def forward(ctx, xyz, shape_bias, spec, pts_per_shape, want_lap, mode, *params):
        ctx.grad_targets = None
        if getattr(spec, "accumulate_in_place", False) and all(
                p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32 for p in params):
            ctx.grad_targets = [p.grad for p in params]
        params = [p.detach() for p in params]
        xyz_d = xyz.detach()
        sb = None if shape_bias is None else shape_bias.detach().contiguous()
        f, g, lap, saved = jet_forward_raw(spec, params, xyz_d, sb, pts_per_shape, want_lap, True, mode)
        ctx.spec, ctx.pps, ctx.want_lap, ctx.mode = spec, pts_per_shape, want_lap, mode
        ctx.has_sb = sb is not None
        ctx.save_for_backward(xyz_d, g, saved, *( [sb] if sb is not None else [] ), *params)
        ctx.set_materialize_grads(False)
        if lap is None:
            lap = torch.empty(0, dtype=torch.float32, device=xyz.device)
            ctx.mark_non_differentiable(lap)
        return f, g, lap

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, f_bar, g_bar, l_bar):
        tensors = ctx.saved_tensors
        xyz, g, saved = tensors[0], tensors[1], tensors[2]
        k = 3
        sb = None
        if ctx.has_sb:
            sb = tensors[3]
            k = 4
        params = list(tensors[k:])
        if not ctx.want_lap:
            l_bar = None
        fb = None if f_bar is None else f_bar.contiguous()
        gb = None if g_bar is None else g_bar.contiguous()
        lb = None if l_bar is None else l_bar.contiguous()
        # Accumulate straight into already-materialised .grad buffers (e.g. the views of dist.FlatGrads) when the caller
        # opted in: the kernels add with atomics anyway, so this removes a zero-fill and an add per parameter.
        direct = ctx.grad_targets
        grads, sb_grad = jet_backward_raw(ctx.spec, params, xyz, sb, ctx.pps, ctx.want_lap, fb, gb, lb, saved,
                                          ctx.mode, grad_out=direct)
        if direct is not None:
            grads = [None] * len(params)
        xyz_grad = None
        if ctx.needs_input_grad[0]:
            if gb is not None or lb is not None:
                raise RuntimeError(
                    "digs_b200: gradient w.r.t. the points through grad/lap outputs needs second-order "
                    "point derivatives, which the fused jet does not provide (SURVEY 8b limit); use the "
                    "reference autograd path for that caller")
            xyz_grad = torch.zeros_like(xyz)
            if fb is not None:
                xyz_grad[:, :ctx.spec.d] = fb[:, None] * g
        return (xyz_grad, sb_grad, None, None, None, None, *grads)


_SIDE_STREAMS = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class JetPairFunction(torch.autograd.Function):
    """Jets of the off-surface cloud (with Laplacian) and the surface cloud (without) as ONE autograd node.

    The two clouds are independent, so their kernels are issued on two streams and joined before returning: the persistent
    kernels of one cloud fill the SMs that the other cloud's last, partially filled wave of tiles leaves idle.  Keeping the
    fork / join inside a single node means autograd (and CUDA-graph capture) only ever sees the caller's stream.
    Parameter gradients of both clouds are summed (or accumulated in place, see JetFunction)."""

    @staticmethod
    def forward(ctx, xyz_n, xyz_m, sb_n, sb_m, spec, pps_n, pps_m, lap_n, mode, *params):
        ctx.grad_targets = None
        if getattr(spec, "accumulate_in_place", False) and all(
                p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32 for p in params):
            ctx.grad_targets = [p.grad for p in params]
        params = [p.detach() for p in params]
        xn, xm = xyz_n.detach(), xyz_m.detach()
        sbn = None if sb_n is None else sb_n.detach().contiguous()
        sbm = None if sb_m is None else sb_m.detach().contiguous()
        main = torch.cuda.current_stream(xn.device)
        side = _side_stream(xn.device)
        side.wait_stream(main)
        fn, gn, ln, saved_n = jet_forward_raw(spec, params, xn, sbn, pps_n, lap_n, True, mode)
        with torch.cuda.stream(side):
            fm, gm, _, saved_m = jet_forward_raw(spec, params, xm, sbm, pps_m, False, True, mode)
        main.wait_stream(side)
        ctx.spec, ctx.pps_n, ctx.pps_m, ctx.lap_n, ctx.mode = spec, pps_n, pps_m, lap_n, mode
        ctx.has_sb = sbn is not None
        ctx.save_for_backward(xn, xm, saved_n, saved_m, *([sbn, sbm] if sbn is not None else []), *params)
        ctx.set_materialize_grads(False)
        if ln is None:
            ln = torch.empty(0, dtype=torch.float32, device=xn.device)
            ctx.mark_non_differentiable(ln)
        return fn, gn, ln, fm, gm

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, fn_bar, gn_bar, ln_bar, fm_bar, gm_bar):
        t = ctx.saved_tensors
        xn, xm, saved_n, saved_m = t[0], t[1], t[2], t[3]
        k = 4
        sbn = sbm = None
        if ctx.has_sb:
            sbn, sbm = t[4], t[5]
            k = 6
        params = list(t[k:])
        c = lambda v: None if v is None else v.contiguous()
        if not ctx.lap_n:
            ln_bar = None
        direct = ctx.grad_targets
        main = torch.cuda.current_stream(xn.device)
        side = _side_stream(xn.device)
        side.wait_stream(main)
        gr_n, sbg_n = jet_backward_raw(ctx.spec, params, xn, sbn, ctx.pps_n, ctx.lap_n, c(fn_bar), c(gn_bar), c(ln_bar),
                                       saved_n, ctx.mode, grad_out=direct)
        with torch.cuda.stream(side):
            gr_m, sbg_m = jet_backward_raw(ctx.spec, params, xm, sbm, ctx.pps_m, False, c(fm_bar), c(gm_bar), None,
                                           saved_m, ctx.mode, grad_out=direct)
        main.wait_stream(side)
        if direct is not None:
            grads = [None] * len(params)
        else:
            grads = [a + b for a, b in zip(gr_n, gr_m)]
        return (None, None, sbg_n, sbg_m, None, None, None, None, None, *grads)
