"""The training iteration captured once into a CUDA graph.

A reference training iteration issues ~900 kernel launches plus nine `.item()` syncs for logging
(train_surface_reconstruction.py:114-158, utils/utils.py:30-36).  `GraphedStep` records

    [batch sampling] -> gradient zero-fill -> weight-operand preparation -> fused step (forward jets of both clouds,
    DiGSLoss, reverse pass: digs_step_fused, two launches) -> [gradient all-reduce over ranks] -> [clip + Adam]

into ONE CUDA graph over static buffers and replays it per iteration with zero host synchronisations.  The parts in
brackets are optional (`sampler=`, a process group with more than one rank, `optimizer=`).  Everything that changes
between iterations lives in device memory:
  * the point batch is copied into the static buffers asynchronously (or drawn inside the graph by a DeviceSampler);
  * the loss weights -- `update_div_weight` rewrites weights[4] every iteration (models/losses.py:208-224) -- are
    re-read from a small device tensor at every replay;
  * the Adam step count is advanced on the device (FusedClipAdam(device_step=True)).
Precision mode fp32 and the loss variants outside the fused kernel take the two-call path
(DiGSNetwork.forward -> DiGSLoss.forward -> backward) inside the same graph.
Gradients land in a `digs_b200.dist.FlatGrads` buffer (every parameter's .grad is a view of it); the eight loss scalars
ride in its trailing slots through the all-reduce, so sharded runs report the global loss without an extra collective.
"""
import torch
import torch.distributed as dist

from .dist import FlatGrads, gather_latent_grads
from .jet import bump_param_epoch
from .step import fused_step, fused_step_supported


class GraphedStep:
    def __init__(self, net, criterion, n_mnfld, n_nonmnfld, batch=1, flat=None, warmup=3, sampler=None, optimizer=None,
                 latent_table=None, group=None, fused=None):
        params = list(net.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("digs_b200.GraphedStep needs the network on a CUDA device -- no CPU fallback")
        d = net.in_dim
        self.net, self.criterion, self.optimizer, self.group = net, criterion, optimizer, group
        self.autodecoder = net.encoder_type == "autodecoder"
        if self.autodecoder and latent_table is None:
            raise ValueError("GraphedStep: an auto-decoder network needs latent_table=(num_shapes, latent_size)")
        self.fused = fused_step_supported(net, criterion) if fused is None else bool(fused)
        if self.fused and not fused_step_supported(net, criterion):
            raise ValueError("GraphedStep(fused=True): precision mode / loss type outside digs_step_fused")
        if self.autodecoder and not self.fused:
            raise NotImplementedError("GraphedStep: the auto-decoder recipe is captured through the fused step only")
        self.sampler = sampler          # optional digs_b200.DeviceSampler: the batch is then drawn INSIDE the graph
        if sampler is not None and (batch != 1 or sampler.n_points != n_mnfld or n_mnfld != n_nonmnfld):
            raise ValueError("GraphedStep: the sampler draws one cloud of n_points surface + n_points off-surface samples")
        if optimizer is not None:
            if getattr(optimizer, "_step_dev", None) is None:
                raise ValueError("GraphedStep(optimizer=...): needs digs_b200.FusedClipAdam(..., device_step=True)")
            flat = optimizer.flat_grads if flat is None else flat
            if flat is not optimizer.flat_grads:
                raise ValueError("GraphedStep: `flat` must be the optimizer's FlatGrads buffer")
        self.flat = flat if flat is not None else FlatGrads(params, extra=8)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        fc = net.decoder.fc_block
        # the three static input buffers are slices of ONE allocation (mnfld | normals | nonmnfld): a caller that keeps its host
        # batch packed the same way (pack_host_batch) pays one host-to-device copy per iteration instead of three
        n_in = batch * (2 * n_mnfld + n_nonmnfld) * d
        self._inputs = torch.zeros(n_in, device=dev)
        o1, o2 = batch * n_mnfld * d, 2 * batch * n_mnfld * d
        self.mnfld = self._inputs[:o1].view(batch, n_mnfld, d)
        self.normals = self._inputs[o1:o2].view(batch, n_mnfld, d)
        self.nonmnfld = self._inputs[o2:].view(batch, n_nonmnfld, d)
        self.latent_table = latent_table
        self.latent_grad = None
        if self.autodecoder:
            if latent_table.device != dev or latent_table.dim() != 2:
                raise ValueError("GraphedStep: latent_table must be a (num_shapes, latent_size) tensor on the network's device")
            self.indices = torch.zeros(batch, dtype=torch.long, device=dev)
            self.latent_grad = torch.zeros_like(latent_table)         # dense, like Adam over the whole table
            if latent_table.requires_grad:
                latent_table.grad = self.latent_grad                  # train_shapespace.py:92-103: Adam over {net, table}
        self._w_host = torch.zeros(8, 6, dtype=torch.float32).pin_memory()   # ring: an in-flight async copy is never overwritten
        self._w_slot = 0
        self._w_dev = torch.zeros(6, dtype=torch.float32, device=dev)
        self._push_weights()
        self._param_ptrs = [p.data_ptr() for p in params]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._body(warm=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        before = _lib.lib().digs_launch_count()
        self.graph, self.loss_dict = self._capture()
        self.launches_per_step = int(_lib.lib().digs_launch_count() - before)   # library kernels recorded in the graph
        self.loss_dict = {k: v.detach() for k, v in self.loss_dict.items()}

    def _capture(self, host_in=None, host_loss=None, pool=None):
        """Records one iteration.  With host_in / host_loss (pinned tensors) the graph starts with the host-to-device copy of
        the batch and ends with the device-to-host copy of the eight loss scalars -- memcpy nodes INSIDE the graph (measured:
        a separate copy between two graph launches costs 15-45 us of stream time each, scripts/e2e_probe.py)."""
        fc = self.net.decoder.fc_block
        graph = torch.cuda.CUDAGraph()
        fc.invalidate_prepared()        # the operand preparation must be recorded inside the graph
        mode = "thread_local" if self.world > 1 else "global"   # NCCL's watchdog thread touches CUDA while we capture
        kw = {"pool": pool} if pool is not None else {}
        with torch.cuda.graph(graph, capture_error_mode=mode, **kw):
            if host_in is not None:
                self._inputs.copy_(host_in, non_blocking=True)
            loss_dict = self._body(warm=False)
            if host_loss is not None:
                host_loss.copy_(self.flat.extra[:8], non_blocking=True)
        return graph, loss_dict

    def release(self):
        """Drop the captured graph.  Call this (or delete the object) BEFORE torch.distributed.destroy_process_group():
        tearing the NCCL communicator down while a graph that captured its kernels is alive blocks forever."""
        self.graph = None
        self._io_graphs = [None, None]

    def pack_host_batch(self, mnfld_points, nonmnfld_points, mnfld_normals):
        """One pinned host tensor laid out like the static device inputs; returns (packed, mnfld, nonmnfld, normals) where the
        last three are views of `packed` -- pass them to step() and the batch travels in a single copy."""
        packed = torch.empty(self._inputs.numel(), dtype=torch.float32).pin_memory()
        o1, o2 = self.mnfld.numel(), self.mnfld.numel() + self.normals.numel()
        views = (packed[:o1].view_as(self.mnfld), packed[o2:].view_as(self.nonmnfld), packed[o1:o2].view_as(self.normals))
        for v, src in zip(views, (mnfld_points, nonmnfld_points, mnfld_normals)):
            v.copy_(src)
        return (packed,) + views

    def _push_weights(self):
        w = list(self.criterion.weights) + [0.0] * 6
        if self.criterion.loss_type == "siren_wo_n":
            w[2] = 0.0                                               # models/losses.py:153
        w = [float(x) for x in w[:6]]
        if w == getattr(self, "_w_last", None):
            return                                                   # unchanged since the last push: nothing to copy
        self._w_last = w
        slot = self._w_host[self._w_slot]
        self._w_slot = (self._w_slot + 1) % self._w_host.shape[0]
        for i in range(6):
            slot[i] = float(w[i])
        self._w_dev.copy_(slot, non_blocking=True)

    def _body(self, warm):
        """One iteration.  The device-weights / in-place-accumulation switches are scoped to this body: eager calls of
        the same criterion / network outside the graph keep their normal behaviour."""
        crit, fc = self.criterion, self.net.decoder.fc_block
        had_dw = getattr(crit, "device_weights", None)
        had_acc = fc.accumulate_grads_in_place
        crit.device_weights = self._w_dev
        fc.accumulate_grads_in_place = True
        try:
            if self.sampler is not None:
                self.sampler.sample_into(self.mnfld, self.normals, self.nonmnfld)
            self.flat.zero_()
            if self.fused:
                latent = None
                if self.autodecoder:
                    latent = self.latent_table.detach().index_select(0, self.indices)      # train_shapespace.py:122
                # the kernel adds the loss scalars straight into the trailing slots of the flat gradient buffer (zeroed
                # with it above): one collective carries gradients and loss terms, and no copy kernels are recorded
                loss_dict, _, lat_grad = fused_step(self.net, crit, self.mnfld, self.nonmnfld, self.normals, latent=latent,
                                                    terms=self.flat.extra)
                if self.autodecoder:
                    # rows of all ranks' shapes -> the dense table gradient every rank updates identically
                    gather_latent_grads(self.latent_grad, self.indices, lat_grad, group=self.group)
                ex = self.flat.extra
                views = {"loss": ex[0], "sdf_term": ex[1], "inter_term": ex[2], "eikonal_term": ex[3], "normals_loss": ex[4],
                         "div_loss": ex[5] if crit.use_div else torch.zeros((), device=ex.device),
                         "latent_reg_term": ex[6], "curv_loss": ex[7]}
            else:
                m = self.mnfld.detach().requires_grad_()
                nm = self.nonmnfld.detach().requires_grad_()
                out = self.net(nm, m)
                loss_dict, _ = crit(out, m, nm, self.normals)
                loss_dict["loss"].backward()
                keys = ("loss", "sdf_term", "inter_term", "eikonal_term", "normals_loss", "div_loss", "latent_reg_term", "curv_loss")
                ex = self.flat.extra
                for i, k in enumerate(keys):
                    ex[i:i + 1].copy_(loss_dict[k].detach().reshape(-1)[:1])
                views = {k: ex[i] for i, k in enumerate(keys)}
            if self.world > 1:
                dist.all_reduce(self.flat.flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.optimizer is not None:
                self.optimizer.step()
            return views
        finally:
            if had_dw is None:
                if hasattr(crit, "device_weights"):
                    del crit.device_weights
            else:
                crit.device_weights = had_dw
            fc.accumulate_grads_in_place = had_acc

    @torch.no_grad()
    def step(self, mnfld_points=None, nonmnfld_points=None, mnfld_normals=None, indices=None):
        """Copies one batch into the static buffers (or, with a sampler, lets the graph draw it), replays the graph.
        Returns the dict of 8 loss tensors (static device tensors, overwritten by the next call; summed over ranks when
        sharded); parameter gradients are in `self.flat` / every `.grad`, latent-table gradients in `self.latent_grad`."""
        if [p.data_ptr() for p in self.net.parameters()] != self._param_ptrs:
            raise RuntimeError("digs_b200.GraphedStep: parameter storage moved since capture (construct FusedClipAdam / "
                               "re-point .data BEFORE GraphedStep) -- the graph would read stale memory")
        if self.sampler is None:
            m, nm, nr = mnfld_points, nonmnfld_points, mnfld_normals
            packed = (not m.is_cuda and m.is_contiguous() and nm.is_contiguous() and nr.is_contiguous()
                      and m.dtype == nm.dtype == nr.dtype == torch.float32
                      and m.untyped_storage().data_ptr() == nm.untyped_storage().data_ptr() == nr.untyped_storage().data_ptr()
                      and m.storage_offset() == 0 and nr.storage_offset() == m.numel()
                      and nm.storage_offset() == m.numel() + nr.numel() and m.numel() == self.mnfld.numel()
                      and nr.numel() == self.normals.numel() and nm.numel() == self.nonmnfld.numel())
            if packed:      # one transfer for the whole batch (pack_host_batch layout)
                host = torch.empty(0, dtype=torch.float32).set_(m.untyped_storage(), 0, (self._inputs.numel(),))
                self._inputs.copy_(host, non_blocking=True)
            else:
                self.mnfld.copy_(m, non_blocking=True)
                self.nonmnfld.copy_(nm, non_blocking=True)
                self.normals.copy_(nr, non_blocking=True)
        if self.autodecoder:
            self.indices.copy_(indices, non_blocking=True)
        self._push_weights()
        self.graph.replay()
        if self.optimizer is not None:
            bump_param_epoch()      # the replayed optimizer wrote the parameters: cached operand copies are stale for eager calls
        return self.loss_dict

    # ---- pipelined host feed ---------------------------------------------------------------------------------------------
    # The reference feeds its loop from DataLoader workers (train_surface_reconstruction.py:44-46, 114-119: the batch of
    # iteration i+1 is prepared while iteration i runs) and reads the loss for logging.  submit() / result() are that loop
    # shape without a host synchronisation per iteration.  Two extra graphs are recorded, one per pinned staging buffer:
    # each begins with the host-to-device copy of ITS staging buffer into the static inputs and ends with the device-to-host
    # copy of the eight loss scalars into ITS pinned result slot, so an iteration is a single graph launch and the host may
    # fill the other staging buffer while it runs.  Every iteration still pays its own copy in and its own read out.
    def _pipe_init(self):
        n_in = self._inputs.numel()
        self._host_in = [torch.zeros(n_in, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._host_loss = [torch.zeros(8, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._io_graphs = [self._capture(self._host_in[k], self._host_loss[k], pool=self.graph.pool())[0] for k in range(2)]
        self._io_done = [torch.cuda.Event(), torch.cuda.Event()]
        self._io_pending = [False, False]
        self._ticket = 0
        torch.cuda.synchronize(self._inputs.device)

    def staging(self, ticket=None):
        """(mnfld, nonmnfld, normals) views of the pinned staging buffer the NEXT submit() will read (or the one of a given
        future ticket): a data loader that writes its batch straight into them saves submit() the host-side copy."""
        if not hasattr(self, "_io_graphs"):
            self._pipe_init()
        k = (self._ticket if ticket is None else ticket) & 1
        if self._io_pending[k]:
            self._io_done[k].synchronize()          # the iteration that last read this buffer must have finished
        buf = self._host_in[k]
        o1, o2 = self.mnfld.numel(), self.mnfld.numel() + self.normals.numel()
        return buf[:o1].view_as(self.mnfld), buf[o2:].view_as(self.nonmnfld), buf[o1:o2].view_as(self.normals)

    @torch.no_grad()
    def submit(self, packed_host=None, indices=None):
        """Queues one iteration and returns a ticket for result().  `packed_host`: a host float32 tensor in pack_host_batch
        layout (copied into the staging buffer), or None when the caller has filled staging() itself.  Does not wait for the
        device unless two iterations are already in flight (then it waits for the older one: its staging buffer is reused)."""
        if self.sampler is not None:
            raise RuntimeError("GraphedStep.submit: this graph draws its batch on the device (sampler=); call step()")
        if not hasattr(self, "_io_graphs"):
            self._pipe_init()
        if packed_host is not None and (packed_host.is_cuda or packed_host.dtype != torch.float32
                                        or packed_host.numel() != self._inputs.numel()):
            raise ValueError("GraphedStep.submit: needs a host float32 tensor in pack_host_batch() layout")
        if [p.data_ptr() for p in self.net.parameters()] != self._param_ptrs:
            raise RuntimeError("digs_b200.GraphedStep: parameter storage moved since capture")
        t = self._ticket
        k = t & 1
        if self._io_pending[k]:
            self._io_done[k].synchronize()          # iteration t - 2 still owns this staging buffer / result slot
        if packed_host is not None:
            self._host_in[k].copy_(packed_host.view(-1))
        if self.autodecoder:
            self.indices.copy_(indices, non_blocking=True)
        self._push_weights()
        self._io_graphs[k].replay()
        self._io_done[k].record(torch.cuda.current_stream(self._inputs.device))
        self._io_pending[k] = True
        if self.optimizer is not None:
            bump_param_epoch()
        self._ticket = t + 1
        return t

    def result(self, ticket):
        """Blocks until iteration `ticket` has finished on the device and returns its loss scalars (loss, sdf, inter,
        eikonal, normals, div, latent_reg, curv) as a list of 8 floats.  Only the two most recent tickets are kept."""
        if not hasattr(self, "_io_graphs") or not (self._ticket - 2 <= ticket < self._ticket):
            raise ValueError("GraphedStep.result: ticket %d is no longer (or not yet) in flight" % ticket)
        k = ticket & 1
        self._io_done[k].synchronize()
        return self._host_loss[k].tolist()
