"""The hot-path step captured once into a CUDA graph.

A reference training iteration issues ~900 kernel launches plus nine `.item()` syncs for logging
(train_surface_reconstruction.py:121-132, utils/utils.py:30-36).  The fused path needs about a dozen launches, at
which point Python / launch latency is comparable to the GPU time of a 30 000-point step.  `GraphedStep` records

    flat-gradient zero-fill -> DiGSNetwork.forward (both clouds) -> DiGSLoss.forward -> loss.backward()

into ONE CUDA graph over static input buffers and replays it per iteration.  Everything that changes between
iterations stays outside the graph or in device memory:
  * the point batches are copied into the static buffers (host or device source, asynchronously);
  * the loss weights -- `update_div_weight` rewrites weights[4] every iteration (models/losses.py:208-224) -- are
    re-read from a small device tensor at every replay (digs_loss_fused_dw);
  * the gradient all-reduce (when sharded over ranks), clipping and the optimizer run after the replay as usual.
Gradients land in a `digs_b200.dist.FlatGrads` buffer (every parameter's .grad is a view of it).
"""
import torch

from .dist import FlatGrads


class GraphedStep:
    def __init__(self, net, criterion, n_mnfld, n_nonmnfld, batch=1, flat=None, warmup=3, sampler=None):
        params = list(net.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("digs_b200.GraphedStep needs the network on a CUDA device -- no CPU fallback")
        if net.encoder_type == "autodecoder":
            raise NotImplementedError("GraphedStep covers the single-shape recipes (encoder_type='none')")
        d = net.in_dim
        self.net, self.criterion = net, criterion
        self.sampler = sampler          # optional digs_b200.DeviceSampler: the batch is then drawn INSIDE the graph
        if sampler is not None and (batch != 1 or sampler.n_points != n_mnfld or n_mnfld != n_nonmnfld):
            raise ValueError("GraphedStep: the sampler draws one cloud of n_points surface + n_points off-surface samples")
        self.flat = flat if flat is not None else FlatGrads(params, extra=8)
        net.decoder.fc_block.accumulate_grads_in_place = True
        self.mnfld = torch.zeros(batch, n_mnfld, d, device=dev).requires_grad_()
        self.nonmnfld = torch.zeros(batch, n_nonmnfld, d, device=dev).requires_grad_()
        self.normals = torch.zeros(batch, n_mnfld, d, device=dev)
        self._w_host = torch.zeros(8, 6, dtype=torch.float32).pin_memory()   # ring: an in-flight async copy is never overwritten
        self._w_slot = 0
        self._w_dev = torch.zeros(6, dtype=torch.float32, device=dev)
        criterion.device_weights = self._w_dev
        self._push_weights()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        before = _lib.lib().digs_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss_dict = self._body()
        self.launches_per_step = int(_lib.lib().digs_launch_count() - before)   # library kernels recorded in the graph
        self.loss_dict = {k: v.detach() for k, v in self.loss_dict.items()}

    def _push_weights(self):
        w = list(self.criterion.weights) + [0.0] * 6
        if self.criterion.loss_type == "siren_wo_n":
            w[2] = 0.0                                               # models/losses.py:153
        slot = self._w_host[self._w_slot]
        self._w_slot = (self._w_slot + 1) % self._w_host.shape[0]
        for i in range(6):
            slot[i] = float(w[i])
        self._w_dev.copy_(slot, non_blocking=True)

    def _body(self):
        if self.sampler is not None:
            self.sampler.sample_into(self.mnfld, self.normals, self.nonmnfld)
        self.flat.zero_()
        out = self.net(self.nonmnfld, self.mnfld)
        loss_dict, _ = self.criterion(out, self.mnfld, self.nonmnfld, self.normals)
        loss_dict["loss"].backward()
        return loss_dict

    @torch.no_grad()
    def step(self, mnfld_points=None, nonmnfld_points=None, mnfld_normals=None):
        """Copies one batch into the static buffers (or, with a sampler, lets the graph draw it), replays the graph.
        Returns the dict of 8 loss tensors (static device tensors, overwritten by the next call); parameter gradients
        are in `self.flat` / every `.grad`."""
        if self.sampler is None:
            self.mnfld.copy_(mnfld_points, non_blocking=True)
            self.nonmnfld.copy_(nonmnfld_points, non_blocking=True)
            self.normals.copy_(mnfld_normals, non_blocking=True)
        self._push_weights()
        self.graph.replay()
        return self.loss_dict
