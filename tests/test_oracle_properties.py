"""Size-independent properties of the oracle's closed forms (CPU): cheap guards on the restated maths."""
import numpy as np


def test_exp_maps_group_properties(oracle):
    rng = np.random.default_rng(1)
    for _ in range(50):
        x = (rng.standard_normal(7) * np.array([0.3, 0.3, 0.3, 0.5, 0.5, 0.5, 0.1])).astype(np.float32)
        T = oracle.exp_sim3(x).astype(np.float64)
        s = np.exp(np.float64(x[6]))
        R = T[:3, :3] / s
        assert np.abs(R @ R.T - np.eye(3)).max() < 5e-6 and abs(np.linalg.det(R) - 1) < 5e-6
        assert np.array_equal(T[3], [0, 0, 0, 1])
        A = oracle.exp_se3(x[:6]).astype(np.float64)
        B = oracle.exp_se3(-x[:6]).astype(np.float64)
        assert np.abs(A @ B - np.eye(4)).max() < 5e-6                 # exp(-x) = exp(x)^-1
    # pure translation
    T = oracle.exp_se3(np.array([1, 2, 3, 0, 0, 0], np.float32))
    assert np.array_equal(T[:3, 3], [1, 2, 3]) and np.array_equal(T[:3, :3], np.eye(3, dtype=np.float32))


def test_huber_and_occupancy_shapes(oracle):
    b = 0.025
    a = np.array([0.0, 1e-6, b * (1 - 1e-6), b, b * (1 + 1e-6), 1.0], np.float32)
    w = oracle.huber_weights(a.copy(), b)
    assert w[0] == 0.0 and np.all(w[1:4] == 1.0) and abs(w[4] - 1.0) < 1e-5 and 0 < w[5] < 1       # continuous at b
    rho_far = (w[5] * a[5]) ** 2
    assert abs(rho_far - (2 * b * 1.0 - b * b)) < 1e-6
    s = np.linspace(-0.05, 0.05, 101).astype(np.float32)
    o = oracle.sdf_to_occupancy(s, 0.01)
    assert o.min() == 0.0 and o.max() == 1.0 and np.all(np.diff(o) <= 0)
    assert abs(float(oracle.sdf_to_occupancy(np.float32(0.0), 0.01)) - 0.5) < 1e-7


def test_pose_jacobian_is_the_derivative_of_the_left_perturbation(oracle):
    """[g, x cross g, g.x] must equal d/d(xi) of f(exp_sim3(xi) x) at xi = 0 for f with gradient g."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal(3)
    g = rng.standard_normal(3)
    J = oracle.pose_jacobian_rows(g[None].astype(np.float32), x[None].astype(np.float32))[0]
    eps = 1e-3
    for k in range(7):
        xi = np.zeros(7, np.float32); xi[k] = eps
        Tp = oracle.exp_sim3(xi).astype(np.float64); Tm = oracle.exp_sim3(-xi).astype(np.float64)
        dp = (Tp[:3, :3] @ x + Tp[:3, 3]) - (Tm[:3, :3] @ x + Tm[:3, 3])
        fd = g @ dp / (2 * eps)
        if k == 6:
            # quirk loss_utils.py:223: exp_sim3 with s <= 1e-8 uses c = 0 but R scaling e^s still applies
            assert abs(fd - J[k]) < 2e-3 * max(1, abs(J[k]))
        else:
            assert abs(fd - J[k]) < 2e-3 * max(1, abs(J[k])), (k, fd, J[k])


def test_render_term_invariants(oracle, oracle_decoders, stages):
    r = oracle.render_term(oracle_decoders["cars"], stages["rnd_rays"], stages["rnd_depth_obs"], stages["sdf_t_obj_cam"],
                           stages["rnd_depths"], stages["sdf_z"], 0.01)
    J, res, ctr = r
    d = stages["rnd_depths"]
    assert np.all(ctr["d_u"] >= d[0] - 1e-4) and np.all(ctr["d_u"] <= 1.1 * d[-1] + 1e-4)    # rendered depth in range
    assert np.all(np.abs(res) <= 0.3 + 1e-7) and ctr["m"] <= ctr["band"] <= ctr["V"]
    assert np.all(np.diff(ctr["ii"]) >= 0)                                               # rows ordered ray-major
