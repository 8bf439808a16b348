"""GPU parity of the mesh-extraction / metric kernels through the C-ABI, against oracle/marching_tets.py,
the marching-cubes vertex rule, scipy's KDTree and analytic surfaces."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _smooth_volume(shape, seed):
    rng = np.random.RandomState(seed)
    g = [np.linspace(-1, 1, n, dtype=np.float32) for n in shape]
    X, Y, Z = np.meshgrid(*g, indexing="ij")
    vol = np.sqrt(X ** 2 + 1.3 * Y ** 2 + 0.8 * Z ** 2) - 0.55
    for _ in range(4):
        k = rng.uniform(1, 4, size=3)
        vol = vol + 0.08 * np.sin(k[0] * X + rng.uniform(0, 6)) * np.cos(k[1] * Y) * np.sin(k[2] * Z + 1.0)
    return vol.astype(np.float32)


def _sphere_volume(n, r=0.6):
    ax = np.linspace(-1, 1, n, dtype=np.float32)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    return (np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - r).astype(np.float32), float(ax[1] - ax[0])


@pytest.mark.parametrize("axis_of", [(0, 1, 2), (1, 0, 2), (2, 0, 1)])
def test_marching_tets_matches_oracle_array_for_array(axis_of):
    from digs_b200.mesh import marching_tets
    from oracle import marching_tets as omt
    vol = _smooth_volume((21, 26, 30), seed=3)
    spacing, origin = (0.11, 0.07, 0.05), (-1.0, 0.25, 2.0)
    v, f, n = marching_tets(torch.tensor(vol, device=DEV), 0.0, spacing, origin, axis_of)
    ov, of = omt.marching_tets(vol, 0.0, spacing, origin, axis_of)
    assert v.shape == ov.shape and f.shape == of.shape and len(of) > 2000
    np.testing.assert_allclose(v.cpu().numpy(), ov, rtol=0, atol=2e-6)
    np.testing.assert_array_equal(f.cpu().numpy(), of)
    nn = n.cpu().numpy()
    np.testing.assert_allclose(np.linalg.norm(nn, axis=1), 1.0, atol=1e-5)


def test_marching_tets_contains_exactly_the_marching_cubes_vertex_set():
    """Lewiner / Lorensen-Cline marching cubes (skimage, utils/utils.py:354) places one vertex on every CELL edge whose ends
    straddle the level, by linear interpolation.  Those are the device vertices with an axis-aligned edge mask."""
    from digs_b200.mesh import marching_tets
    from digs_b200 import _lib
    from digs_b200.jet import _ptr, _stream
    from oracle import marching_tets as omt
    vol = _smooth_volume((24, 24, 24), seed=5)
    tv = torch.tensor(vol, device=DEV)
    v, f, _ = marching_tets(tv, 0.0)
    flags = torch.empty(vol.size, dtype=torch.uint8, device=DEV)
    ntri = torch.empty(vol.size, dtype=torch.uint8, device=DEV)
    _lib.check(_lib.lib().digs_mt_classify(_ptr(tv), 24, 24, 24, 0.0, _ptr(flags), _ptr(ntri), _stream()), "mt_classify")
    fl = flags.cpu().numpy()
    bits = np.unpackbits(fl[:, None], axis=1, bitorder="little")[:, :7].astype(bool)      # (points, mask-1)
    np.testing.assert_array_equal(bits.reshape(24, 24, 24, 7), omt.edge_flags(vol))
    is_axis = np.zeros(7, bool)
    is_axis[[0, 1, 3]] = True                                                              # masks 1, 2, 4
    axis_vertex = np.broadcast_to(is_axis, bits.shape)[bits]                               # per vertex, in emission order
    mc = omt.mc_edge_vertices(vol)
    got = v.cpu().numpy()[axis_vertex]
    assert got.shape == mc.shape and len(mc) > 500
    np.testing.assert_allclose(got, mc, rtol=0, atol=2e-6)
    # independent restatement of the vertex rule (sign-change product test, per axis)
    cnt = sum(int(((np.moveaxis(vol, a, 0)[:-1] < 0) != (np.moveaxis(vol, a, 0)[1:] < 0)).sum()) for a in range(3))
    assert cnt == len(mc)


@pytest.mark.parametrize("n", [64, 192])
def test_sphere_mesh_is_closed_oriented_and_has_the_analytic_area(n):
    from digs_b200.mesh import marching_tets
    from oracle.marching_tets import mesh_stats
    vol, h = _sphere_volume(n)
    v, f, nrm = marching_tets(torch.tensor(vol, device=DEV), 0.0, (h, h, h), (-1.0, -1.0, -1.0))
    vv, ff = v.cpu().numpy(), f.cpu().numpy()
    closed, oriented, euler, area, volume = mesh_stats(vv, ff)
    assert closed and oriented and euler == 2
    assert abs(area - 4 * np.pi * 0.36) < 0.02 * 4 * np.pi * 0.36
    assert abs(volume - 4 / 3 * np.pi * 0.216) < 0.02 * 4 / 3 * np.pi * 0.216            # positive: normals point outwards
    assert np.abs(np.linalg.norm(vv, axis=1) - 0.6).max() < h * h                         # linear interpolation error
    cosang = (nrm.cpu().numpy() * vv / 0.6).sum(1)
    assert cosang.min() > 0.99


def test_empty_and_degenerate_volumes():
    from digs_b200.mesh import marching_tets
    vol = torch.ones(8, 9, 10, device=DEV)
    v, f, n = marching_tets(vol, 0.0)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    with pytest.raises(RuntimeError):
        marching_tets(torch.ones(1, 4, 4, device=DEV), 0.0)
    with pytest.raises(RuntimeError):
        marching_tets(torch.ones(4, 4, 4), 0.0)                                           # CPU tensor: no fallback


@pytest.mark.parametrize("d,nr,nq", [(3, 30011, 20003), (2, 5000, 7001), (3, 1, 100), (3, 4097, 1)])
def test_nn_query_matches_kdtree(d, nr, nq):
    from scipy.spatial import cKDTree
    from digs_b200.mesh import nn_query
    rng = np.random.RandomState(d * 1000 + nr)
    ref = rng.uniform(-1, 1, size=(nr, d)).astype(np.float32)
    qry = rng.uniform(-1, 1, size=(nq, d)).astype(np.float32)
    dist, idx = nn_query(torch.tensor(ref, device=DEV), torch.tensor(qry, device=DEV), return_index=True)
    kd, ki = cKDTree(ref.astype(np.float64)).query(qry.astype(np.float64))
    np.testing.assert_allclose(dist.cpu().numpy(), kd, rtol=2e-6, atol=1e-7)
    same = idx.cpu().numpy() == ki
    assert same.mean() > 0.999                                                            # ties / fp32 rounding only
    bad = ~same
    if bad.any():
        alt = np.linalg.norm(ref[idx.cpu().numpy()[bad]].astype(np.float64) - qry[bad], axis=1)
        np.testing.assert_allclose(alt, kd[bad], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("eval_type", ["Default", "DeepSDF"])
def test_compute_dists_matches_reference_formula(eval_type):
    """compute_metrics_srb.py:38-57 restated with scipy's KDTree (the reference's own backend) as the checker."""
    from scipy.spatial import cKDTree
    from digs_b200.mesh import compute_dists
    rng = np.random.RandomState(1)
    gt = rng.normal(size=(40000, 3))
    gt = 0.5 * gt / np.linalg.norm(gt, axis=1, keepdims=True)
    recon = rng.normal(size=(50000, 3))
    recon = 0.51 * recon / np.linalg.norm(recon, axis=1, keepdims=True) + rng.normal(scale=2e-3, size=(50000, 3))
    re2gt, _ = cKDTree(recon).query(gt)
    gt2re, _ = cKDTree(gt).query(recon)
    if eval_type == "DeepSDF":
        want = ((re2gt ** 2).mean() + (gt2re ** 2).mean(), max(re2gt.max(), gt2re.max()), (re2gt ** 2).mean(), (gt2re ** 2).mean(),
                re2gt.max(), gt2re.max())
    else:
        want = (0.5 * (re2gt.mean() + gt2re.mean()), max(re2gt.max(), gt2re.max()), re2gt.mean(), gt2re.mean(), re2gt.max(),
                gt2re.max())
    got = compute_dists(recon, gt, eval_type)
    np.testing.assert_allclose(got, want, rtol=2e-5)


def test_sample_surface_is_uniform_on_the_mesh():
    from digs_b200.mesh import marching_tets, sample_surface
    vol, h = _sphere_volume(96)
    v, f, _ = marching_tets(torch.tensor(vol, device=DEV), 0.0, (h, h, h), (-1.0, -1.0, -1.0))
    g = torch.Generator(device=DEV)
    g.manual_seed(0)
    pts, fi = sample_surface(v, f, 200000, generator=g)
    p = pts.cpu().numpy()
    assert np.abs(np.linalg.norm(p, axis=1) - 0.6).max() < 2 * h * h
    assert np.abs(p.mean(0)).max() < 5e-3                                                 # uniform on the sphere: zero mean
    assert abs((p[:, 2] > 0.3).mean() - 0.25) < 5e-3                                      # cap z > r/2 holds 1/4 of the area
    assert int(fi.max()) < f.shape[0]


def test_implicit2mesh_end_to_end(tmp_path):
    """utils.implicit2mesh (utils/utils.py:329-370) through the drop-in: volume, iso-surface, origin / scale / translate,
    PLY export (train_surface_reconstruction.py:105) -- the mesh must be the network's own zero level set."""
    import digs_b200
    from oracle.marching_tets import mesh_stats
    torch.manual_seed(3627473)
    net = digs_b200.DiGSNetwork(0, 3, 256, "sine", "none", 4, "mfgi", [1.6, 0.1]).to(DEV)
    out = digs_b200.implicit2mesh(net.decoder, None, 64, translate=[0.1, -0.2, 0.05], scale=2.0, get_mesh=True, device=DEV)
    assert out["mesh_trace"] is None
    mesh = out["mesh_obj"]
    v, f = mesh.vertices, mesh.faces
    assert len(v) > 1000 and f.min() >= 0 and f.max() < len(v)
    closed, oriented, euler, area, volume = mesh_stats(v, f)
    assert closed and oriented and volume > 0
    # undo scale / translate (utils/utils.py:358) and evaluate the network there: the vertices are on its zero level set
    pts = torch.tensor((v + np.array([0.1, -0.2, 0.05], dtype=np.float32)) * 2.0, device=DEV)
    with torch.no_grad():
        sdf = net.decoder(pts.unsqueeze(0)).reshape(-1)
    cell = 2.2 / 63
    assert float(sdf.abs().max()) < 0.5 * cell                                            # within the interpolation error of a cell
    assert float(sdf.abs().mean()) < 0.05 * cell
    # vertex normals agree with the network's own gradient direction
    p = pts.clone().requires_grad_()
    o = net(p.unsqueeze(0), None)
    g = o["nonmanifold_pnts_grad"].reshape(-1, 3)
    cosang = torch.nn.functional.cosine_similarity(g, torch.tensor(mesh.vertex_normals, device=DEV), dim=1)
    assert float(cosang.mean()) > 0.99
    path = tmp_path / "mesh.ply"
    mesh.export(str(path), vertex_normal=True)
    raw = path.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    assert b"element vertex %d" % len(v) in head and b"element face %d" % len(f) in head and b"property float nx" in head
    assert len(body) == len(v) * 24 + len(f) * 13
    np.testing.assert_array_equal(np.frombuffer(body[:12], "<f4"), v[0])


def test_occupancy_iou_matches_reference_formula():
    """compute_metrics_shapenet.py:131-136: occupancy = (decoder(points) < 0); IoU = |gt & pred| / |gt | pred|."""
    import digs_b200
    from oracle import ref_autograd
    from _util import head_of
    torch.manual_seed(3627473)
    net = digs_b200.DiGSNetwork(0, 3, 256, "sine", "none", 4, "mfgi", [1.6, 0.1]).to(DEV)
    rng = np.random.RandomState(0)
    pts = rng.uniform(-0.55, 0.55, size=(100000, 3)).astype(np.float32)
    occ = (np.linalg.norm(pts, axis=1) < 0.5).astype(np.uint8)           # "ground truth": a ball a little larger than the initial sphere
    got = digs_b200.occupancy_iou(net.decoder, pts, occ)
    params = [p.detach().cpu() for p in net.decoder.fc_block.flat_params()]
    ref = ref_autograd.grid_values(params, torch.tensor(pts), head_of(net)).numpy()
    pred = (ref < 0).astype(np.uint8)
    want = (occ & pred).sum() / (occ | pred).sum()
    assert 0.05 < want < 1.0
    assert abs(got - want) < 2e-4          # a handful of points within fp32 rounding of the level set may flip
