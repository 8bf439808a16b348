"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

The hot path shards only where it is embarrassingly parallel:
  * training points: both clouds are split across ranks, adjoint seeds are normalised by the GLOBAL point
    counts (DiGSLoss.global_counts) so that ONE sum-allreduce of the flat parameter-gradient buffer gives
    exactly the single-GPU gradient; loss scalars ride in the same buffer;
  * grid queries: contiguous iy slabs per rank (digs_b200.grid.slab_range), no collective;
  * shapes (auto-decoder): whole shapes per rank.
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
