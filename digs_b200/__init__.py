"""digs_b200 -- B200-native (sm_100a) hot path for DiGS: fused value/gradient/Laplacian jets of the sine MLP,
the DiGSLoss reverse pass through them, and the dense-grid SDF query.  Drop-in names mirror the reference:

    from digs_b200 import DiGSNetwork, DiGSLoss, implicit2mesh

All arithmetic runs in libdigs_b200.so (C-ABI in include/digs_b200.h); there is no CPU fallback.
"""
from .network import DiGSNetwork, Decoder, FCBlock, Sine
from .losses import DiGSLoss
from .grid import (implicit2mesh, get_3d_grid, grid_axes, grid_query, slab_range, compute_deriv_props,
                   get_2d_grid_uniform)
from .graph import GraphedStep
from .step import fused_step, fused_step_supported
from .optim import FusedClipAdam
from .sampler import DeviceSampler
from .mesh import marching_tets, nn_query, compute_dists, sample_surface, occupancy_iou, TriMesh

__all__ = ["DiGSNetwork", "Decoder", "FCBlock", "Sine", "DiGSLoss", "implicit2mesh", "get_3d_grid", "grid_axes",
           "grid_query", "slab_range", "GraphedStep", "fused_step", "fused_step_supported", "FusedClipAdam", "DeviceSampler", "compute_deriv_props", "get_2d_grid_uniform",
           "marching_tets", "nn_query", "compute_dists", "sample_surface", "occupancy_iou", "TriMesh"]
