"""Device-resident point cloud + on-device batch sampling.

Replaces the DataLoader hop of the reference (ReconDataset.__getitem__, recon_dataset.py:143-174: a host permutation of
the whole cloud, np.random.uniform for the off-surface points, and a pinned H2D copy every iteration) with one kernel
launch (digs_sample_batch) that writes the batch straight into device buffers -- e.g. the static inputs of GraphedStep.
"""
import ctypes

import torch

from . import _lib
from .jet import _ptr, _stream


class DeviceSampler:
    def __init__(self, points, normals, n_points, grid_range=1.1, seed=0, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("digs_b200.DeviceSampler needs a CUDA device -- no CPU fallback")
        self.points = torch.as_tensor(points, dtype=torch.float32).contiguous().to(dev)
        self.normals = None if normals is None else torch.as_tensor(normals, dtype=torch.float32).contiguous().to(dev)
        self.n_cloud, self.d = self.points.shape
        self.n_points, self.grid_range, self.seed = int(n_points), float(grid_range), int(seed)
        self.counter = torch.zeros(2, dtype=torch.int64, device=dev)       # (iteration, ticket)

    def sample_into(self, mnfld, normals, nonmnfld):
        """Fills (bs=1, n_points, d) fp32 CUDA buffers in place (no_grad; safe on leaf tensors that require grad)."""
        with torch.no_grad():
            rc = _lib.lib().digs_sample_batch(_ptr(self.points), _ptr(self.normals), self.n_cloud, self.d, self.n_points,
                                              self.grid_range, ctypes.c_uint64(self.seed), _ptr(self.counter),
                                              _ptr(mnfld), _ptr(normals), _ptr(nonmnfld), _stream())
        _lib.check(rc, "sample_batch")

    def sample(self):
        dev = self.points.device
        m = torch.empty(1, self.n_points, self.d, device=dev)
        nm = torch.empty(1, self.n_points, self.d, device=dev)
        nrm = torch.empty(1, self.n_points, self.d, device=dev) if self.normals is not None else None
        self.sample_into(m, nrm, nm)
        return m, nrm, nm
