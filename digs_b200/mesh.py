"""Mesh extraction and reconstruction metrics on the device.

What the reference does on the host with third-party packages, per evaluated shape:
  * utils/utils.py:354-358      skimage.measure.marching_cubes(volume, level=0, spacing) + origin / scale / translate;
  * utils/utils.py:360-366      trimesh.Trimesh(vertices, faces, vertex_normals=...) and `.export(path, vertex_normal=True)`
                                 (train_surface_reconstruction.py:105, visualizations.py:1024);
  * compute_metrics_srb.py:75   trimesh.sample.sample_surface(mesh, 1_000_000);
  * compute_metrics_srb.py:38-57 two scipy KDTree queries -> Chamfer / Hausdorff distances.
Here the SDF volume never leaves HBM: the iso-surface is extracted by the marching-tetrahedra kernels of
csrc/mesh.cu, surface samples are drawn with torch on the device, and nearest neighbours come from the brute-force
kernel `digs_nn_query`.  skimage / trimesh are not needed.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .jet import _ptr, _stream, _require_cuda

_POPCOUNT = None


def _popcount_table(device):
    global _POPCOUNT
    if _POPCOUNT is None or _POPCOUNT.device != device:
        _POPCOUNT = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int32, device=device)
    return _POPCOUNT


def marching_tets(volume, level=0.0, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), axis_of=(0, 1, 2), normals=True):
    """Iso-surface of a dense fp32 CUDA volume (n0, n1, n2).

    Memory axis k is logical axis axis_of[k]; a grid index i along it sits at origin[k] + spacing[k] * i.
    Returns (verts (V,3) float32, faces (F,3) int32, vertex normals (V,3) float32 or None), all on the device.
    Vertices lie on grid edges (cell edges, face diagonals, body diagonals of the Kuhn split) at the linear
    interpolation point of `level`; faces are oriented with their normal towards larger values."""
    _require_cuda(volume, "volume")
    if volume.dim() != 3 or volume.dtype != torch.float32:
        raise ValueError("marching_tets: volume must be a 3-D float32 tensor")
    vol = volume.contiguous()
    n0, n1, n2 = vol.shape
    dev = vol.device
    L = _lib.lib()
    flags = torch.empty(vol.numel(), dtype=torch.uint8, device=dev)
    ntri = torch.empty(vol.numel(), dtype=torch.uint8, device=dev)
    _lib.check(L.digs_mt_classify(_ptr(vol), n0, n1, n2, float(level), _ptr(flags), _ptr(ntri), _stream()), "mt_classify")
    vcnt = _popcount_table(dev)[flags.long()]
    vinc = torch.cumsum(vcnt, 0, dtype=torch.int64)
    tinc = torch.cumsum(ntri, 0, dtype=torch.int64)
    n_verts, n_tris = int(vinc[-1]), int(tinc[-1])          # the one host sync: output sizes
    if n_verts >= 2 ** 31 or 3 * n_tris >= 2 ** 31:
        raise RuntimeError("marching_tets: mesh exceeds int32 indexing")
    voff = (vinc - vcnt).to(torch.int32)
    toff = (tinc - ntri).to(torch.int32)
    del vinc, tinc, vcnt
    verts = torch.empty(n_verts, 3, dtype=torch.float32, device=dev)
    nrm = torch.empty(n_verts, 3, dtype=torch.float32, device=dev) if normals else None
    faces = torch.empty(n_tris, 3, dtype=torch.int32, device=dev)
    if n_verts:
        ax = (ctypes.c_int * 3)(*[int(a) for a in axis_of])
        org = (ctypes.c_float * 3)(*[float(o) for o in origin])
        sp = (ctypes.c_float * 3)(*[float(s) for s in spacing])
        _lib.check(L.digs_mt_emit(_ptr(vol), n0, n1, n2, float(level), ax, org, sp, _ptr(flags), _ptr(ntri), _ptr(voff),
                                  _ptr(toff), _ptr(verts), _ptr(nrm), _ptr(faces) if n_tris else None, _stream()), "mt_emit")
    return verts, faces, nrm


def nn_query(ref, query, return_index=False):
    """Distance from every query point to its nearest neighbour in `ref` (exact, fp32); KDTree(ref).query(query)."""
    _require_cuda(ref, "ref")
    ref = ref.detach().to(torch.float32).contiguous()
    query = query.detach().to(device=ref.device, dtype=torch.float32).contiguous()
    if ref.dim() != 2 or query.dim() != 2 or ref.shape[1] != query.shape[1]:
        raise ValueError("nn_query: ref (Nr,d) and query (Nq,d) with the same d")
    dist = torch.empty(query.shape[0], dtype=torch.float32, device=ref.device)
    idx = torch.empty(query.shape[0], dtype=torch.int32, device=ref.device) if return_index else None
    _lib.check(_lib.lib().digs_nn_query(_ptr(ref), ref.shape[0], _ptr(query), query.shape[0], ref.shape[1], _ptr(dist),
                                        _ptr(idx), _stream()), "nn_query")
    return (dist, idx) if return_index else dist


def _as_cuda(points, device=None):
    t = torch.as_tensor(np.asarray(points), dtype=torch.float32) if not torch.is_tensor(points) else points.to(torch.float32)
    if t.device.type != "cuda":
        t = t.to(device if device is not None else "cuda")
    return t


def compute_dists(recon_points, gt_points, eval_type="Default"):
    """compute_metrics_srb.py:38-57 with both KDTree queries on the device.  Same return tuple:
    (chamfer, hausdorff, cd_re2gt, cd_gt2re, hd_re2gt, hd_gt2re) as Python floats."""
    re, gt = _as_cuda(recon_points), None
    gt = _as_cuda(gt_points, re.device)
    re2gt = nn_query(re, gt)       # recon_kd_tree.query(gt_points)
    gt2re = nn_query(gt, re)       # gt_kd_tree.query(recon_points)
    if eval_type == "DeepSDF":
        stats = torch.stack([(re2gt ** 2).mean(), (gt2re ** 2).mean(), re2gt.max(), gt2re.max()]).tolist()
        cd_re2gt, cd_gt2re, hd_re2gt, hd_gt2re = stats
        chamfer = cd_re2gt + cd_gt2re
    else:
        stats = torch.stack([re2gt.mean(), gt2re.mean(), re2gt.max(), gt2re.max()]).tolist()
        cd_re2gt, cd_gt2re, hd_re2gt, hd_gt2re = stats
        chamfer = 0.5 * (cd_re2gt + cd_gt2re)
    return chamfer, max(hd_re2gt, hd_gt2re), cd_re2gt, cd_gt2re, hd_re2gt, hd_gt2re


def sample_surface(verts, faces, count, generator=None):
    """Area-weighted uniform samples on a triangle mesh: trimesh.sample.sample_surface(mesh, count)
    (compute_metrics_srb.py:75, eval_shapespace.py:162).  Returns (points (count,3), face_index (count,))."""
    f = faces.long()
    a, b, c = verts[f[:, 0]], verts[f[:, 1]], verts[f[:, 2]]
    area = 0.5 * torch.linalg.cross(b - a, c - a).norm(dim=1)
    cum = torch.cumsum(area.double(), 0)
    r = torch.rand(count, device=verts.device, dtype=torch.float64, generator=generator) * cum[-1]
    fi = torch.searchsorted(cum, r).clamp_(max=f.shape[0] - 1)
    uv = torch.rand(count, 2, device=verts.device, generator=generator)
    flip = uv.sum(1) > 1
    uv[flip] = 1 - uv[flip]
    pts = a[fi] + uv[:, :1] * (b[fi] - a[fi]) + uv[:, 1:] * (c[fi] - a[fi])
    return pts, fi


def occupancy_iou(decoder, points, occupancies):
    """IoU of the predicted interior (SDF < 0) against ground-truth occupancies at query points: the ShapeNet metric of
    compute_metrics_shapenet.py:131-136, evaluated with the value-only kernel and reduced on the device."""
    pts = _as_cuda(points)
    with torch.no_grad():
        sdf = decoder(pts.reshape(1, -1, pts.shape[-1])).reshape(-1)
    pred = sdf < 0
    occ = torch.as_tensor(np.asarray(occupancies) if not torch.is_tensor(occupancies) else occupancies).to(pts.device).reshape(-1) != 0
    return float((occ & pred).sum()) / float((occ | pred).sum())


class TriMesh:
    """The slice of trimesh.Trimesh the reference touches: vertices / faces / vertex_normals arrays and
    `.export(path, vertex_normal=True)` (binary PLY)."""

    def __init__(self, vertices, faces, vertex_normals=None):
        self.vertices_t, self.faces_t, self.normals_t = vertices, faces, vertex_normals

    @property
    def vertices(self):
        return self.vertices_t.cpu().numpy()

    @property
    def faces(self):
        return self.faces_t.cpu().numpy()

    @property
    def vertex_normals(self):
        return None if self.normals_t is None else self.normals_t.cpu().numpy()

    def sample_surface(self, count, generator=None):
        return sample_surface(self.vertices_t, self.faces_t, count, generator)

    def export(self, file_obj, vertex_normal=True):
        v, f = self.vertices.astype("<f4"), self.faces.astype("<i4")
        n = self.vertex_normals if vertex_normal else None
        props = "property float x\nproperty float y\nproperty float z\n"
        cols = [v]
        if n is not None:
            props += "property float nx\nproperty float ny\nproperty float nz\n"
            cols.append(n.astype("<f4"))
        header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\n%selement face %d\n"
                  "property list uchar int vertex_indices\nend_header\n" % (v.shape[0], props, f.shape[0]))
        rec = np.empty(f.shape[0], dtype=[("n", "u1"), ("i", "<i4", (3,))])
        rec["n"], rec["i"] = 3, f
        with open(file_obj, "wb") as fh:
            fh.write(header.encode("ascii"))
            fh.write(np.ascontiguousarray(np.concatenate(cols, axis=1)).tobytes())
            fh.write(rec.tobytes())
