"""Marching-tetrahedra fallback of MeshExtractor (CPU): closed, consistently oriented, on the level set."""
import numpy as np
import pytest

from dsp_slam_b200.mesh import marching_tetrahedra


def _sphere(n, r, c=(0.03, -0.02, 0.05)):
    ax = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    return np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) - r, 2.0 / (n - 1), np.array(c)


@pytest.mark.parametrize("n,r", [(16, 0.6), (33, 0.45)])
def test_sphere_is_closed_oriented_and_on_the_surface(n, r):
    vol, h, c = _sphere(n, r)
    v, f = marching_tetrahedra(vol, 0.0, (h, h, h))
    v = v.astype(np.float64) + np.array([-1.0, -1.0, -1.0])
    assert v.shape[0] > 100 and f.dtype == np.int32 and f.min() >= 0 and f.max() < v.shape[0]
    # vertices lie on the sphere (linear interpolation of an exact distance field: error O(h^2 / r))
    d = np.linalg.norm(v - c, axis=1)
    assert np.abs(d - r).max() < 0.6 * h * h / r + 1e-6
    # every undirected edge is shared by exactly two triangles, with opposite directions (closed + oriented)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).astype(np.int64)
    key = e[:, 0] * v.shape[0] + e[:, 1]
    rkey = e[:, 1] * v.shape[0] + e[:, 0]
    assert np.unique(key).size == key.size                      # no directed edge twice
    assert np.array_equal(np.sort(key), np.sort(rkey))          # each has its reverse
    # Euler characteristic of a sphere
    E = key.size // 2
    assert v.shape[0] - E + f.shape[0] == 2
    # outward orientation: positive signed volume ~ 4/3 pi r^3
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    vol_signed = np.einsum("ij,ij->i", p0 - c, np.cross(p1 - c, p2 - c)).sum() / 6.0
    assert abs(vol_signed - 4.0 / 3.0 * np.pi * r ** 3) < 0.03 * 4.0 / 3.0 * np.pi * r ** 3


def test_no_crossing_gives_empty_mesh_and_level_shift():
    vol, h, _ = _sphere(12, 0.5)
    v, f = marching_tetrahedra(vol + 10.0, 0.0, (h, h, h))
    assert v.shape == (0, 3) and f.shape == (0, 3)
    v1, f1 = marching_tetrahedra(vol, 0.1, (h, h, h))           # larger level set -> larger sphere
    v0, f0 = marching_tetrahedra(vol, 0.0, (h, h, h))
    assert np.linalg.norm(v1 + [-1, -1, -1], axis=1).mean() > np.linalg.norm(v0 + [-1, -1, -1], axis=1).mean()
