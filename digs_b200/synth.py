"""Synthetic inputs of the shapes the reference recipes use.

There is no network access for SRB / ShapeNet / DFAUST, so benchmarks and tests draw point clouds from
analytic surfaces, normalised like recon_dataset.py:67-78 (centre, scale by max |x|) and sampled per
iteration like recon_dataset.py:143-174 (random subset of the cloud + uniform off-surface points).
"""
import numpy as np


def surface_cloud(n_points=100000, kind="torus", seed=0, dim=3):
    """Unoriented cloud + analytic normals on a closed surface, centred and scaled to max |x| = 1."""
    rng = np.random.RandomState(seed)
    if dim == 2:
        t = rng.uniform(0, 2 * np.pi, n_points)
        pts = 0.5 * np.stack([np.cos(t), np.sin(t)], 1)
        nrm = np.stack([np.cos(t), np.sin(t)], 1)
        return pts.astype(np.float32), nrm.astype(np.float32)
    if kind == "sphere":
        v = rng.normal(size=(n_points, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pts, nrm = 0.5 * v, v
    elif kind == "torus":
        R, r = 0.6, 0.25
        a = rng.uniform(0, 2 * np.pi, n_points)
        b = rng.uniform(0, 2 * np.pi, n_points)
        pts = np.stack([(R + r * np.cos(b)) * np.cos(a), (R + r * np.cos(b)) * np.sin(a), r * np.sin(b)], 1)
        nrm = np.stack([np.cos(b) * np.cos(a), np.cos(b) * np.sin(a), np.sin(b)], 1)
    else:
        raise ValueError(kind)
    pts = pts - pts.mean(0, keepdims=True)
    pts = pts / np.abs(pts).max()
    return pts.astype(np.float32), nrm.astype(np.float32)


def sample_batch(cloud, normals, n_points=15000, seed=0, box=1.1):
    """One iteration's batch: (mnfld (1,n,d), normals (1,n,d), nonmnfld (1,n,d)) float32 numpy."""
    rng = np.random.RandomState(seed)
    idx = rng.permutation(cloud.shape[0])[:n_points]
    d = cloud.shape[1]
    nonm = rng.uniform(-box, box, size=(n_points, d)).astype(np.float32)
    return cloud[idx][None], normals[idx][None], nonm[None]
