"""CPU checks of the iso-surface oracle (oracle/marching_tets.py): it is what the GPU mesh tests compare against, so its own
anchors are analytic surfaces and an independent loop restatement of the marching-cubes vertex rule."""
import numpy as np
import pytest

from oracle import marching_tets as omt


def _grid(n):
    ax = np.linspace(-1, 1, n, dtype=np.float32)
    return np.meshgrid(ax, ax, ax, indexing="ij"), float(ax[1] - ax[0])


def test_oracle_sphere_and_torus_topology_area_volume():
    (X, Y, Z), h = _grid(40)
    v, f = omt.marching_tets(np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.6, 0.0, (h, h, h), (-1, -1, -1))
    closed, oriented, euler, area, vol = omt.mesh_stats(v, f)
    assert closed and oriented and euler == 2
    assert abs(area / (4 * np.pi * 0.36) - 1) < 0.01 and abs(vol / (4 / 3 * np.pi * 0.216) - 1) < 0.01
    v, f = omt.marching_tets(np.sqrt((np.sqrt(X ** 2 + Y ** 2) - 0.6) ** 2 + Z ** 2) - 0.2, 0.0, (h, h, h), (-1, -1, -1))
    closed, oriented, euler, area, vol = omt.mesh_stats(v, f)
    assert closed and oriented and euler == 0
    assert abs(area / (4 * np.pi ** 2 * 0.12) - 1) < 0.02 and abs(vol / (2 * np.pi ** 2 * 0.6 * 0.04) - 1) < 0.05


def test_oracle_axis_permutation_gives_the_same_surface():
    (X, Y, Z), h = _grid(18)
    vol = (np.sqrt(X ** 2 + 2 * Y ** 2 + Z ** 2) - 0.5).astype(np.float32)
    v0, f0 = omt.marching_tets(vol, 0.0, (h, h, h), (-1, -1, -1))
    v1, f1 = omt.marching_tets(np.ascontiguousarray(vol.transpose(1, 0, 2)), 0.0, (h, h, h), (-1, -1, -1), axis_of=(1, 0, 2))
    s0, s1 = omt.mesh_stats(v0, f0), omt.mesh_stats(v1, f1)
    assert s0[:3] == s1[:3] == (True, True, 2)
    assert abs(s0[3] - s1[3]) < 2e-3 * s0[3] and abs(s0[4] - s1[4]) < 2e-3 * s0[4] and s1[4] > 0


def test_marching_cubes_vertex_rule_against_a_plain_loop():
    rng = np.random.RandomState(0)
    vol = rng.normal(size=(6, 7, 5)).astype(np.float32)
    want = []
    for i in range(6):
        for j in range(7):
            for k in range(5):
                for m, (di, dj, dk) in ((4, (1, 0, 0)), (2, (0, 1, 0)), (1, (0, 0, 1))):
                    ii, jj, kk = i + di, j + dj, k + dk
                    if ii < 6 and jj < 7 and kk < 5 and (vol[i, j, k] < 0) != (vol[ii, jj, kk] < 0):
                        t = (0 - vol[i, j, k]) / (vol[ii, jj, kk] - vol[i, j, k])
                        want.append((i + t * di, j + t * dj, k + t * dk))
    got = omt.mc_edge_vertices(vol)
    want = np.array(sorted(want), dtype=np.float32)
    got = got[np.lexsort((got[:, 2], got[:, 1], got[:, 0]))]
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, atol=1e-6)


def test_oracle_empty_volume():
    v, f = omt.marching_tets(np.ones((4, 4, 4), np.float32))
    assert v.shape == (0, 3) and f.shape == (0, 3)
