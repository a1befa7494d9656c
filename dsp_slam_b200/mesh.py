"""Iso-surface extraction fallback for MeshExtractor when scikit-image is not installed.

The reference calls skimage.measure.marching_cubes_lewiner on the (n,n,n) SDF grid
(reconstruct/utils.py:120-140).  DSP-SLAM's own environment has scikit-image, and the drop-in uses it
when present; this module provides a dependency-free alternative so `extract_mesh_from_code` also works
without it: marching TETRAHEDRA on the Kuhn (6 tetrahedra per cube, shared main diagonal) subdivision,
which is translation-consistent (no cracks between cubes) and needs no 256-entry case tables.  The mesh is a
valid, consistently oriented (normals towards growing SDF, i.e. outwards) triangulation of the same level
set; vertex count / order differ from skimage's Lewiner marching cubes.
"""
import numpy as np

# cube corner offsets (x, y, z) and the six tetrahedra around the 0-6 diagonal
_CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.int64)
_TETS = np.array([[0, 5, 1, 6], [0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6]], dtype=np.int64)


def marching_tetrahedra(volume, level=0.0, spacing=(1.0, 1.0, 1.0)):
    """volume: (nx, ny, nz) float array sampled on a regular lattice with the given spacing.
    Returns (vertices (V,3) float32, faces (F,3) int32).  Empty arrays if the level set is not crossed."""
    vol = np.asarray(volume, dtype=np.float64)
    nx, ny, nz = vol.shape
    sp = np.asarray(spacing, dtype=np.float64)
    if min(nx, ny, nz) < 2:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)
    # global ids of the 8 corners of every cube: (C, 8)
    ix, iy, iz = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), np.arange(nz - 1), indexing="ij")
    base = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)                  # (C,3)
    corner = base[:, None, :] + _CORNERS[None, :, :]                                # (C,8,3)
    gid = (corner[..., 0] * ny + corner[..., 1]) * nz + corner[..., 2]              # (C,8)
    flat = vol.ravel()
    # keep only cubes that the level set crosses
    cv = flat[gid]
    crossed = (cv.min(axis=1) < level) & (cv.max(axis=1) >= level)
    gid = gid[crossed]
    if gid.shape[0] == 0:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)
    tg = gid[:, _TETS].reshape(-1, 4)                                               # (T,4) global vertex ids
    tv = flat[tg]
    inside = tv < level
    cnt = inside.sum(axis=1)
    keep = (cnt > 0) & (cnt < 4)
    tg, tv, inside, cnt = tg[keep], tv[keep], inside[keep], cnt[keep]
    order = np.argsort(~inside, axis=1, kind="stable")                              # inside vertices first
    tg = np.take_along_axis(tg, order, axis=1)
    tv = np.take_along_axis(tv, order, axis=1)

    def pos(g):
        z = g % nz
        y = (g // nz) % ny
        x = g // (ny * nz)
        return np.stack([x, y, z], axis=-1).astype(np.float64) * sp

    tp = pos(tg)                                                                    # (T,4,3)

    tri_a, tri_b = [], []        # per triangle corner: the two lattice vertices of the edge it lies on
    ref_dir = []                 # direction from the inside to the outside of the tetrahedron

    def emit(sel, edges):
        """edges: list of 3 (i, j) column pairs -> one triangle per selected tetrahedron."""
        if not sel.any():
            return
        a = np.stack([tg[sel, i] for i, _ in edges], axis=1)
        b = np.stack([tg[sel, j] for _, j in edges], axis=1)
        tri_a.append(a)
        tri_b.append(b)
        k = cnt[sel][0]
        ins = tp[sel][:, :k].mean(axis=1)
        outs = tp[sel][:, k:].mean(axis=1)
        ref_dir.append(outs - ins)

    s1, s2, s3 = cnt == 1, cnt == 2, cnt == 3
    emit(s1, [(0, 1), (0, 2), (0, 3)])                                              # one vertex inside
    emit(s3, [(3, 0), (3, 1), (3, 2)])                                              # one vertex outside
    emit(s2, [(0, 2), (0, 3), (1, 3)])                                              # two inside: a quad = two triangles
    emit(s2, [(0, 2), (1, 3), (1, 2)])
    if not tri_a:
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)
    A = np.concatenate(tri_a)                                                       # (F,3)
    B = np.concatenate(tri_b)
    R = np.concatenate(ref_dir)
    # one mesh vertex per crossed lattice edge
    lo, hi = np.minimum(A, B), np.maximum(A, B)
    key = lo * np.int64(nx * ny * nz) + hi
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    faces = inv.reshape(-1, 3)
    ua, ub = uniq // (nx * ny * nz), uniq % (nx * ny * nz)
    va, vb = flat[ua], flat[ub]
    t = np.where(vb != va, (level - va) / np.where(vb != va, vb - va, 1.0), 0.5)
    verts = pos(ua) + t[:, None] * (pos(ub) - pos(ua))
    # orient every triangle so that its normal points from inside (sdf < level) to outside
    p0, p1, p2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    nrm = np.cross(p1 - p0, p2 - p0)
    flip = np.einsum("ij,ij->i", nrm, R) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    # drop triangles that collapsed (level set passing exactly through lattice vertices)
    area2 = np.einsum("ij,ij->i", nrm, nrm)
    faces = faces[area2 > 1e-30]
    return verts.astype(np.float32), faces.astype(np.int32)
