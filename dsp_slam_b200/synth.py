"""Seeded synthetic detections with the shapes DSP-SLAM's sequence loaders hand to the optimiser.

Mirrors what reconstruct/kitti_sequence.py:114-216 produces per detection (T_cam_obj initial guess,
surface points in the camera frame, foreground rays + depths, background rays) for an analytic
latent-conditioned ellipsoid -- the shape family the fixture decoders in tests/golden/ are fitted
to (tools/fit_fixture_decoder.py).  numpy only; no CUDA, no torch.  SURVEY.md section 8(d).
"""
import numpy as np

F32 = np.float32

CLASS_RADII = {"cars": (0.30, 0.25, 0.60), "chairs": (0.35, 0.45, 0.35)}


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def make_object(seed, n_pts, n_fg_rays=None, n_bg_rays=0, cls="cars", code_scale=0.1,
                noise=0.01, trans_jitter=0.15, yaw_jitter=0.1, init_code_frac=None):
    """One synthetic detection.

    Returns dict with (all float32, Fortran-ordered like pybind11's Eigen casters deliver them,
    src/LocalMapping_util.cc:179-180):
      t_cam_obj_init (4,4)  initial Sim(3) object->camera guess
      t_cam_obj_gt   (4,4)  ground truth
      pts   (n_pts,3)       surface points, camera frame
      rays  (n_fg+n_bg,3)   ray directions (z = 1), foreground first
      depth (n_fg,)         observed depth of the foreground rays
      code_gt (64,), code_init (64,) or None
    """
    rng = np.random.default_rng(10_000 + seed)
    radii0 = np.array(CLASS_RADII[cls])
    z_gt = code_scale * rng.standard_normal(64)
    radii = radii0 * (1.0 + z_gt[:3])
    s = rng.uniform(1.5, 2.5)
    yaw = rng.uniform(-np.pi, np.pi)
    t = np.array([rng.uniform(-5, 5), rng.uniform(0.5, 1.5), rng.uniform(6, 25)])
    flip = np.diag([1.0, -1.0, -1.0])          # camera y down / object y up
    R = _rot_y(yaw) @ flip
    T_gt = np.eye(4)
    T_gt[:3, :3] = s * R
    T_gt[:3, 3] = t

    def surface(n):
        d = rng.standard_normal((n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        p_obj = d * radii
        # keep the camera-facing one of each antipodal pair (proxy for visibility)
        pa = (p_obj @ (s * R).T) + t
        pb = ((-p_obj) @ (s * R).T) + t
        p = np.where((np.linalg.norm(pa, axis=1) <= np.linalg.norm(pb, axis=1))[:, None], pa, pb)
        return p + noise * rng.standard_normal((n, 3))

    pts = surface(n_pts)
    out = dict(t_cam_obj_gt=T_gt, pts=pts, code_gt=z_gt)

    if n_fg_rays is None:
        n_fg_rays = n_pts
    if n_fg_rays or n_bg_rays:
        fg_src = pts[:n_fg_rays] if n_fg_rays <= n_pts else surface(n_fg_rays)
        fg = fg_src / fg_src[:, 2:3]
        depth = fg_src[:, 2].copy()
        bg = np.zeros((0, 3))
        if n_bg_rays:
            centre = t / t[2]
            half = 1.0 * s / t[2]
            Rinv = R.T / s
            got = []
            while sum(len(g) for g in got) < n_bg_rays:
                uv = centre[:2] + rng.uniform(-half, half, size=(4 * n_bg_rays, 2))
                r = np.concatenate([uv, np.ones((len(uv), 1))], axis=1)
                # ray/ellipsoid test in the radii-normalised object frame (inflated 15 %)
                o = (Rinv @ (-t)) / (1.15 * radii)
                dirs = (r @ Rinv.T) / (1.15 * radii)
                a = (dirs * dirs).sum(1)
                b = (dirs * o).sum(1)
                c = (o * o).sum() - 1.0
                miss = (b * b - a * c) < 0
                got.append(r[miss])
            bg = np.concatenate(got)[:n_bg_rays]
        out["rays"] = np.concatenate([fg, bg], axis=0)
        out["depth"] = depth

    # initial guess: translation + yaw jitter
    dyaw = yaw_jitter * rng.standard_normal()
    dt = trans_jitter * rng.standard_normal(3)
    T0 = np.eye(4)
    T0[:3, :3] = s * (_rot_y(yaw + dyaw) @ flip)
    T0[:3, 3] = t + dt
    out["t_cam_obj_init"] = T0
    out["code_init"] = None if init_code_frac is None else init_code_frac * z_gt

    for k, v in list(out.items()):
        if isinstance(v, np.ndarray):
            out[k] = np.asfortranarray(v.astype(F32))
    return out


def make_batch(n_obj, n_pts, n_fg_rays=None, n_bg_rays=0, cls="cars", seed0=0, **kw):
    clss = cls if isinstance(cls, (list, tuple)) else [cls] * n_obj
    return [make_object(seed0 + i, n_pts, n_fg_rays, n_bg_rays, cls=clss[i], **kw)
            for i in range(n_obj)]
