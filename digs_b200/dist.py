"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

The hot path shards only where it is embarrassingly parallel:
  * training points: both clouds are split across ranks, adjoint seeds are normalised by the GLOBAL point
    counts (DiGSLoss.global_counts) so that ONE sum-allreduce of the flat parameter-gradient buffer gives
    exactly the single-GPU gradient; loss scalars ride in the same buffer;
  * grid queries: contiguous iy slabs per rank (digs_b200.grid.slab_range), no collective;
  * shapes (auto-decoder, train_shapespace.py:118-149): whole shapes per rank (shard_shapes).  The decoder gradients are
    summed like above; the latent rows a rank touched travel in one all-gather of (indices, rows), and every rank then
    holds the full dense latent-table gradient -- the reference runs dense Adam over the whole table
    (train_shapespace.py:84-103: rows with zero gradient still move by momentum), so the replicated tables stay
    identical without ever broadcasting them.
The reference has no working multi-GPU path (nn.DataParallel over a batch of one cloud,
train_surface_reconstruction.py:55-57), so single-GPU equivalence is the parity test.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [begin, end) share of n items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_points(points, rank, world):
    """points (bs, N, w) -> this rank's (bs, N_r, w) slice along the point axis."""
    b, e = shard_range(points.shape[1], rank, world)
    return points[:, b:e]


def shard_shapes(indices, rank, world):
    """The contiguous share of a batch of shape indices (tensor or sequence) that `rank` trains on."""
    b, e = shard_range(len(indices), rank, world)
    return indices[b:e]


def gather_latent_grads(latent_grad_dense, local_indices, local_rows, group=None):
    """latent_grad_dense (num_shapes, l) <- rows of ALL ranks: one all-gather of the (index, gradient row) pairs each rank
    produced for its shapes, scattered (added) into the zeroed dense table gradient.  Equal shard sizes per rank."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    latent_grad_dense.zero_()
    if world == 1:
        latent_grad_dense.index_add_(0, local_indices, local_rows)
        return latent_grad_dense
    n, l = local_rows.shape
    rows = torch.empty(world * n, l, dtype=local_rows.dtype, device=local_rows.device)
    idx = torch.empty(world * n, dtype=local_indices.dtype, device=local_indices.device)
    dist.all_gather_into_tensor(rows, local_rows.contiguous(), group=group)
    dist.all_gather_into_tensor(idx, local_indices.contiguous(), group=group)
    latent_grad_dense.index_add_(0, idx, rows)
    return latent_grad_dense


class FlatGrads:
    """Views every parameter's .grad into one contiguous fp32 buffer (+ `extra` trailing scalars) so the
    data-parallel reduction is a single collective (1.06 MB for the 4x256 network)."""

    def __init__(self, params, extra=8):
        self.params = [p for p in params]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total + extra, dtype=torch.float32, device=dev)
        self.extra = self.flat[total:]
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def allreduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat
