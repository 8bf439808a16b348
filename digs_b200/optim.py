"""Fused optimizer tail: clip_grad_norm_(10) + Adam as two launches over flat buffers.

Replaces `torch.nn.utils.clip_grad_norm_(net.parameters(), 10.)` + `optimizer.step()` of the reference loops
(train_surface_reconstruction.py:130-132, train_shapespace.py:147-149).  Parameters and gradients are re-pointed into
two contiguous buffers (every `p.data` / `p.grad` becomes a view), so the update is one pass over 264 449 floats instead
of ~40 foreach kernels, and composes with `dist.FlatGrads.allreduce()` and `GraphedStep`.
"""
import ctypes

import torch

from . import _lib
from .dist import FlatGrads
from .jet import _ptr, _stream, bump_param_epoch


class FusedClipAdam:
    def __init__(self, params, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, max_norm=10.0, flat_grads=None, device_step=False):
        self.params = list(params)
        if not self.params or not self.params[0].is_cuda:
            raise RuntimeError("digs_b200.FusedClipAdam needs CUDA parameters -- no CPU fallback")
        self.lr, self.betas, self.eps, self.max_norm = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(max_norm)
        self.flat_grads = flat_grads if flat_grads is not None else FlatGrads(self.params, extra=8)
        self.n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat_params = torch.empty(self.n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                view = self.flat_params[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view                       # same Parameter object, storage now inside the flat buffer
                off += p.numel()
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self._scratch = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        # device_step: the Adam step count lives in device memory and is advanced by the kernel itself, so step() can
        # be recorded into a CUDA graph (digs_b200.GraphedStep(optimizer=...)) and replayed
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev) if device_step else None
        bump_param_epoch()          # parameter storage moved into the flat buffer

    def zero_grad(self):
        self.flat_grads.zero_()

    @torch.no_grad()
    def step(self):
        self.step_count += 1
        dev = self.flat_params.device
        with torch.cuda.device(dev):
            if self._step_dev is not None:
                rc = _lib.lib().digs_clip_adam_dev(_ptr(self.flat_params), _ptr(self.flat_grads.flat), _ptr(self.exp_avg),
                                                   _ptr(self.exp_avg_sq), self.n, self.lr, self.betas[0], self.betas[1],
                                                   self.eps, self.max_norm, _ptr(self._step_dev), _ptr(self._scratch),
                                                   _stream(dev))
            else:
                rc = _lib.lib().digs_clip_adam(_ptr(self.flat_params), _ptr(self.flat_grads.flat), _ptr(self.exp_avg),
                                               _ptr(self.exp_avg_sq), self.n, self.lr, self.betas[0], self.betas[1],
                                               self.eps, self.max_norm, self.step_count, _ptr(self._scratch), _stream(dev))
        _lib.check(rc, "clip_adam")
        bump_param_epoch()          # the kernel wrote the parameters through raw pointers: tensor versions did not move
