"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the iso-surface extraction in digs_b200/csrc/mesh.cu.

What it is checked against.  The reference extracts meshes with skimage.measure.marching_cubes
(utils/utils.py:354-355; scikit_image==0.18.1, requirements.txt:4), which is third-party code that is neither under
/root/reference nor installed here, and no reference test pins a mesh: PARITY UNPINNED for the triangulation.  What
IS pinned is marching cubes' vertex rule -- one vertex per cell edge whose endpoint values straddle the level, at the
linear interpolation point (`mc_edge_vertices` below; Lorensen & Cline 1987, Lewiner et al. 2003 section 2) -- and the
Chamfer metric built on the vertices (compute_metrics_srb.py:38-57, checked against scipy's KDTree in the tests).

`marching_tets` restates the device algorithm step by step (Kuhn split of every cell into six tetrahedra around the
main diagonal, unique vertices owned by the low end of their edge, faces oriented towards larger values), in the
same output order, so the device result can be compared array against array.
"""
import numpy as np

TETS = ((1, 2), (1, 4), (2, 1), (2, 4), (4, 1), (4, 2))     # corners 0, a, a|b, 7


def _off(m):
    return ((m >> 2) & 1, (m >> 1) & 1, m & 1)


def _shifted(vol, m):
    """View of vol at index + offset(m), cropped to the points where that neighbour exists."""
    o = _off(m)
    n0, n1, n2 = vol.shape
    return vol[o[0]:, o[1]:, o[2]:], (n0 - o[0], n1 - o[1], n2 - o[2])


def edge_flags(vol, level=0.0):
    """(n0,n1,n2,7) bool: edge with offset mask m (1..7) starting at the point straddles the level."""
    inside = vol < level
    fl = np.zeros(vol.shape + (7,), dtype=bool)
    for m in range(1, 8):
        nb, (s0, s1, s2) = _shifted(inside, m)
        fl[:s0, :s1, :s2, m - 1] = inside[:s0, :s1, :s2] != nb
    return fl


def mc_edge_vertices(vol, level=0.0):
    """Marching cubes' vertex set in index space: crossings on the axis-aligned cell edges only (masks 1, 2, 4)."""
    v, _ = _vertices(vol, level, masks=(1, 2, 4))
    return v


def _vertices(vol, level, masks=tuple(range(1, 8))):
    vol = np.asarray(vol, dtype=np.float32)
    fl = edge_flags(vol, level)
    keep = np.zeros(7, dtype=bool)
    keep[[m - 1 for m in masks]] = True
    p0, p1, p2, mi = np.nonzero(fl & keep)                    # C order: by point, then by mask -- the device's vertex order
    m = mi + 1
    o = np.stack([(m >> 2) & 1, (m >> 1) & 1, m & 1], 1)
    v0 = vol[p0, p1, p2]
    v1 = vol[p0 + o[:, 0], p1 + o[:, 1], p2 + o[:, 2]]
    t = (np.float32(level) - v0) / (v1 - v0)
    idx = np.stack([p0, p1, p2], 1).astype(np.float32) + t[:, None] * o.astype(np.float32)
    return idx, fl


def marching_tets(vol, level=0.0, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), axis_of=(0, 1, 2)):
    """Returns (verts (V,3) float32 in logical coordinates, faces (F,3) int32)."""
    vol = np.asarray(vol, dtype=np.float32)
    n0, n1, n2 = vol.shape
    idx, fl = _vertices(vol, level)
    sp, org = np.asarray(spacing, np.float32), np.asarray(origin, np.float32)

    def to_logical(ix):
        out = np.empty_like(ix)
        for k in range(3):
            out[:, axis_of[k]] = sp[k] * ix[:, k] + org[k]
        return out

    verts = to_logical(idx)
    cnt = fl.sum(-1).reshape(-1)
    voff = np.cumsum(cnt) - cnt
    flat_fl = fl.reshape(-1, 7)
    # bits below each mask, for the rank of an edge among its point's flagged edges
    below = np.cumsum(flat_fl, 1) - flat_fl

    inside = vol < level
    c0, c1, c2 = np.meshgrid(np.arange(n0 - 1), np.arange(n1 - 1), np.arange(n2 - 1), indexing="ij")
    c0, c1, c2 = c0.ravel(), c1.ravel(), c2.ravel()
    cell = (c0 * n1 + c1) * n2 + c2

    def corner(mask):
        o = _off(mask)
        return c0 + o[0], c1 + o[1], c2 + o[2]

    def edge(sel, ca, cb):
        """vertex id and logical position of the crossing between corners ca, cb (arrays of masks) of cells `sel`."""
        lo, hi = np.minimum(ca, cb), np.maximum(ca, cb)
        m = hi ^ lo
        w = [c0[sel] + ((lo >> 2) & 1), c1[sel] + ((lo >> 1) & 1), c2[sel] + (lo & 1)]
        wl = (w[0] * n1 + w[1]) * n2 + w[2]
        vid = voff[wl] + below[wl, m - 1]
        va = vol[w[0], w[1], w[2]]
        vb = vol[c0[sel] + ((hi >> 2) & 1), c1[sel] + ((hi >> 1) & 1), c2[sel] + (hi & 1)]
        t = (np.float32(level) - va) / (vb - va)
        o = np.stack([(m >> 2) & 1, (m >> 1) & 1, m & 1], 1).astype(np.float32)
        ix = np.stack(w, 1).astype(np.float32) + t[:, None] * o
        return vid, to_logical(ix)

    def corner_pos(sel, masks):
        ix = np.stack([c0[sel] + ((masks >> 2) & 1), c1[sel] + ((masks >> 1) & 1), c2[sel] + (masks & 1)], 1).astype(np.float32)
        return to_logical(ix)

    tris, keys = [], []
    for q, (a, b) in enumerate(TETS):
        cr = np.array([0, a, a | b, 7])
        ins = np.stack([inside[corner(int(mk))] for mk in cr], 1)          # (cells, 4)
        ni = ins.sum(1)
        for case in (1, 2, 3):
            sel = np.nonzero(ni == case)[0]
            if sel.size == 0:
                continue
            order_in = np.argsort(~ins[sel], 1, kind="stable")             # inside corners first, original order kept
            crs = cr[order_in]                                              # (k, 4): [inside..., outside...]
            in_c, out_c = crs[:, :case], crs[:, case:]
            dirv = -sum(corner_pos(sel, in_c[:, i]) for i in range(case)) / np.float32(case) \
                + sum(corner_pos(sel, out_c[:, i]) for i in range(4 - case)) / np.float32(4 - case)
            if case in (1, 3):
                apex = in_c[:, 0] if case == 1 else out_c[:, 0]
                base = out_c if case == 1 else in_c
                quads = [[edge(sel, apex, base[:, i]) for i in range(3)]]
            else:
                e = [edge(sel, in_c[:, 0], out_c[:, 0]), edge(sel, in_c[:, 0], out_c[:, 1]),
                     edge(sel, in_c[:, 1], out_c[:, 1]), edge(sel, in_c[:, 1], out_c[:, 0])]
                quads = [[e[0], e[1], e[2]], [e[0], e[2], e[3]]]
            for k, tri in enumerate(quads):
                (i0, p0), (i1, p1), (i2, p2) = tri
                nrm = np.cross(p1 - p0, p2 - p0)
                flip = (nrm * dirv).sum(1) < 0
                f = np.stack([i0, np.where(flip, i2, i1), np.where(flip, i1, i2)], 1)
                tris.append(f)
                keys.append(cell[sel] * 12 + q * 2 + k)
    if not tris:
        return verts, np.zeros((0, 3), np.int32)
    faces = np.concatenate(tris)[np.argsort(np.concatenate(keys), kind="stable")]
    return verts, faces.astype(np.int32)


def mesh_stats(verts, faces):
    """(closed, consistently oriented, Euler characteristic, area, signed volume) of a triangle mesh."""
    f = faces.astype(np.int64)
    nv = int(verts.shape[0])
    he = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    und = np.sort(he, 1)
    key = und[:, 0] * nv + und[:, 1]
    uniq, counts = np.unique(key, return_counts=True)
    closed = bool(np.all(counts == 2))
    dkey = he[:, 0] * nv + he[:, 1]
    oriented = closed and np.unique(dkey).size == dkey.size     # every directed edge once => the two uses are opposite
    euler = int(np.unique(f).size) - int(uniq.size) + int(f.shape[0])
    a, b, c = (verts[f[:, i]].astype(np.float64) for i in range(3))
    cr = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(cr, axis=1).sum()
    volume = (a * cr).sum() / 6.0
    return closed, oriented, euler, area, volume
