"""Pins the CPU oracle against analytic known-answer tests.

The reference ships no tests or golden vectors and cannot be built here (PARITY UNPINNED), so
these closed-form identities -- each derived from the reference's own formulas -- are what anchors the oracle.
"""
import numpy as np
import pytest

import orc
from eslam_b200 import synth

K = np.array([[535.4, 0, 320.1], [0, 539.2, 247.6], [0, 0, 1.0]])
I7 = np.array([0, 0, 0, 0, 0, 0, 1.0])


def quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


@pytest.mark.parametrize("theta", [0.0, 1e-6, 1e-3, 1.0, 3.0])
def test_exp_log_roundtrip(theta):  # C.1
    rng = np.random.default_rng(1)
    axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
    xi = np.concatenate([axis * theta, rng.normal(size=3)])
    T = orc.se3_exp(xi)
    assert abs(np.linalg.norm(T[3:]) - 1) < 1e-12 and T[6] >= 0
    back = orc.se3_log(T)
    # exp() uses R = I + W + W^2 for theta < 1e-5 (se3quat.h:290-301): first-order accurate only there
    tol = 1e-9 if theta < 1e-5 else 2e-6 if theta > 2.9 else 1e-8
    assert np.allclose(back, xi, atol=tol)


def test_exp_matches_rodrigues():
    xi = np.array([0.3, -0.2, 0.5, 1.0, 2.0, 3.0])
    T = orc.se3_exp(xi)
    th = np.linalg.norm(xi[:3]); W = np.array([[0, -xi[2], xi[1]], [xi[2], 0, -xi[0]], [-xi[1], xi[0], 0]])
    R = np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * W @ W
    V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * W + (th - np.sin(th)) / th ** 3 * W @ W
    assert np.allclose(quat_R(T[3:]), R, atol=1e-13)
    assert np.allclose(T[:3], V @ xi[3:], atol=1e-13)


def test_se3_mul_inverse():
    a = orc.se3_exp([0.1, 0.2, 0.3, 1, 2, 3]); b = orc.se3_exp([-0.4, 0.1, 0.2, 0.5, -1, 0.3])
    ab = orc.se3_mul(a, b)
    assert np.allclose(quat_R(ab[3:]), quat_R(a[3:]) @ quat_R(b[3:]), atol=1e-13)
    assert np.allclose(ab[:3], quat_R(a[3:]) @ b[:3] + a[:3], atol=1e-13)
    e = orc.se3_mul(a, orc.se3_inverse(a))
    assert np.allclose(e, I7, atol=1e-13)


def test_min_log_error_yaw_ambiguity():  # C.2
    e = np.concatenate([orc.se3_exp([0.1, -0.3, 0.7, 1.0, -2.0, 0.5]), [0.3, 0.5, 0.2]])
    assert np.allclose(orc.min_log_error_9dof(e, e), 0, atol=1e-9)
    for yaw in (np.pi / 2, np.pi):
        r = orc.ellipsoid_rotate(e, yaw)
        if abs(abs(yaw) - np.pi / 2) < 1e-9:
            assert np.allclose(r[7:9], e[7:9][::-1])
        assert np.allclose(orc.min_log_error_9dof(e, r), 0, atol=2e-8), yaw
    # quirk kept from the reference: a relative yaw of EXACTLY 180 deg makes SE3Quat::log divide by sqrt(1-d^2)=0
    # (se3quat.h:252); when that candidate is the first of the four (-90 deg applied to a -90 deg copy) its NaN norm
    # poisons minCoeff (Ellipsoid.cpp:112) and the whole residual is non-finite.
    r = orc.ellipsoid_rotate(e, -np.pi / 2)
    assert not np.isfinite(orc.min_log_error_9dof(e, r)[:6]).all()
    r = orc.ellipsoid_rotate(e, -np.pi / 2 + 1e-3)
    assert np.isfinite(orc.min_log_error_9dof(e, r)).all()


def test_sphere_bbox():  # C.3
    r, z = 0.4, 3.0
    e = np.array([0, 0, z, 0, 0, 0, 1, r, r, r])
    bb = orc.ellipsoid_bbox(e, I7, K)
    k = r / np.sqrt(z * z - r * r)
    exp = np.array([K[0, 2] - K[0, 0] * k, K[1, 2] - K[1, 1] * k, K[0, 2] + K[0, 0] * k, K[1, 2] + K[1, 1] * k])
    assert np.allclose(bb, exp, rtol=1e-10)


def test_bbox_matches_sampled_projection():
    e = np.concatenate([orc.se3_exp([0.2, -0.1, 0.9, 0.3, -0.2, 3.5]), [0.5, 0.3, 0.2]])
    bb = orc.ellipsoid_bbox(e, I7, K)
    R = quat_R(e[3:7])
    u, v = np.meshgrid(np.linspace(0, np.pi, 400), np.linspace(0, 2 * np.pi, 800))
    P = np.stack([e[7] * np.sin(u) * np.cos(v), e[8] * np.sin(u) * np.sin(v), e[9] * np.cos(u)], -1).reshape(-1, 3) @ R.T + e[:3]
    px = K[0, 0] * P[:, 0] / P[:, 2] + K[0, 2]; py = K[1, 1] * P[:, 1] / P[:, 2] + K[1, 2]
    assert np.allclose(bb, [px.min(), py.min(), px.max(), py.max()], atol=0.02)


def _views(n, radius=3.0):
    Twc = []
    for f in range(n):
        a = 2 * np.pi * f / n
        pos = np.array([radius * np.cos(a), radius * np.sin(a), 1.0 + 0.3 * np.sin(3 * a)])
        fwd = -pos / np.linalg.norm(pos); right = np.cross(fwd, [0, 0, 1.0]); right /= np.linalg.norm(right); down = np.cross(fwd, right)
        Twc.append(np.concatenate([pos, synth._rot_to_quat(np.stack([right, down, fwd], 1))]))
    return np.array(Twc)


def test_initializer_inverts_projection():  # C.4
    # NOTE (reference behaviour, Initializer.cpp:219-227): the half-axes are taken from the eigenvalues of the 4x4 Q, which
    # equal those of Q33 only for an ellipsoid centred at the world origin; off-origin the axes are a few % off by design.
    Twc = _views(8)
    for centre, axes_tol in (([0.0, 0.0, 0.0], 1e-6), ([0.1, -0.05, 0.2], 0.05)):
        e = np.array(centre + [0, 0, np.sin(0.35), np.cos(0.35), 0.45, 0.25, 0.35])
        det = np.array([np.concatenate([orc.ellipsoid_bbox(e, orc.se3_inverse(T), K), [1.0]]) for T in Twc])
        assert det[:, :4].min() > 1 and det[:, 2].max() < 638 and det[:, 3].max() < 478
        out, ok = orc.init_quadric(Twc, det, K)
        assert ok
        assert np.allclose(out[:3], e[:3], atol=1e-6)
        assert np.allclose(sorted(out[7:]), sorted(e[7:]), rtol=axes_tol, atol=1e-6)
    # too few planes -> failure (Initializer.cpp:40-45)
    _, ok2 = orc.init_quadric(Twc[:2], det[:2], K)
    assert not ok2


def test_gravity_edge():  # C.5
    for alpha in (0.0, 0.1, 0.7):
        e = np.concatenate([orc.se3_exp([alpha, 0, 0, 0.3, 0.2, 0.1]), [0.3, 0.2, 0.1]])
        r = orc.edge_error(2, I7, e, [0, 0, 1, 0], K)
        assert abs(r[0] - alpha) < 1e-7


def test_numeric_jacobian_vs_fd_highres():  # C.6 : g2o delta=1e-9 central differences ~ smooth derivative
    rng = np.random.default_rng(7)
    e = np.concatenate([orc.se3_exp([0.1, 0.05, 0.6, 0.3, -0.2, 0.4]), [0.5, 0.3, 0.4]])
    Twc = _views(3, 4.0)
    Tcw = np.array([orc.se3_inverse(T) for T in Twc])
    meas = np.zeros((3, 10)); meas[0, :4] = orc.ellipsoid_bbox(e, Tcw[0], K) + 1.5
    meas[1] = orc.ellipsoid_transform_from(orc.ellipsoid_exp_update(e, rng.normal(0, 0.02, 9)), Tcw[1]); meas[2, :4] = [0, 0, 1, 0]
    g = orc.Graph([0, 1, 2], [0, 1, 0], [0, 0, 0], meas, [1.0, 2.0, 3.0], Tcw, 1, K)
    r, J, chi = orc.edges_eval(g, e[None])
    h = 1e-5
    for k, D in enumerate((4, 9, 1)):
        for d in range(9):
            u = np.zeros(9); u[d] = h
            rp, _, _ = orc.edges_eval(g, orc.ellipsoid_exp_update(e, u)[None], jac=False)
            rm, _, _ = orc.edges_eval(g, orc.ellipsoid_exp_update(e, -u)[None], jac=False)
            ref = (rp[k, :D] - rm[k, :D]) / (2 * h)
            # delta=1e-9 central differences sit on the fp64 noise floor: |r|*eps*cond/2e-9 (pixels-sized residuals of the
            # 2-D edge go through a 3x3 inverse -> ~1e-2 absolute; metre/radian-sized residuals -> ~1e-6)
            assert np.allclose(J[k, :D, d], ref, rtol=1e-4, atol=5e-2 if k == 0 else 2e-5), (k, d)
        assert np.allclose(chi[k], g.info[k] * (r[k, :D] ** 2).sum())


def test_lm_recovers_ellipsoid():
    rng = np.random.default_rng(3)
    e = np.array([0.2, -0.1, 0.3, 0, 0, np.sin(0.2), np.cos(0.2), 0.4, 0.25, 0.3])
    Twc = _views(12, 3.5); Tcw = np.array([orc.se3_inverse(T) for T in Twc])
    et, fr, ob, meas, info = [2], [0], [0], [np.r_[0, 0, 1, 0, np.zeros(6)]], [100.0 ** 2]
    for f in range(12):
        et.append(0); fr.append(f); ob.append(0); meas.append(np.r_[orc.ellipsoid_bbox(e, Tcw[f], K), np.zeros(6)]); info.append(1.0)
    for f in range(12):
        et.append(1); fr.append(f); ob.append(0); meas.append(orc.ellipsoid_transform_from(e, Tcw[f])); info.append(10000.0)
    g = orc.Graph(et, fr, ob, np.array(meas), info, Tcw, 1, K)
    e0 = orc.ellipsoid_exp_update(e, rng.normal(0, 0.05, 9))
    out, st = orc.optimize_objects(g, e0[None], 10)
    assert st.chi2_final < 1e-8 * max(1.0, st.chi2_initial) and st.iterations >= 2
    assert np.allclose(orc.min_log_error_9dof(out[0], e), 0, atol=1e-5)


def test_symmetry_cost_zero_at_true_plane():  # C.7
    rng = np.random.default_rng(5)
    half = rng.uniform([-0.3, 0.02, -0.3], [0.3, 0.4, 0.3], size=(150, 3))
    ang, dis = np.deg2rad(3.0), 0.05
    n = np.array([np.sin(ang), -np.cos(ang), 0.0]); pl = np.r_[n, -dis]
    S = half + (dis - half @ n)[:, None] * n[None, :] + np.abs(half[:, 1:2]) * n[None, :]  # one side of the plane
    mirror = S - 2 * (S @ n - dis)[:, None] * n[None, :]
    cloud = np.concatenate([S, mirror]).astype(np.float32).astype(np.float64)  # float-representable => NN distance exactly 0
    depth = np.zeros((480, 640), np.uint16)  # range image 0 everywhere: inside-bbox points cost 0, others pure NN
    T_oc = np.array([0, 0, -3.0, 0, 0, 0, 1.0])
    err = orc.symmetry_error(depth, [100, 100, 500, 400], T_oc, cloud, pl, 0.0, 1)
    assert 0 <= err < 1e-9
    bad = orc.symmetry_error(depth, [0, 0, 1, 1], T_oc, cloud, np.r_[n, -(dis + 0.1)], 0.0, 1)
    assert bad > 1e-3


def test_crop_mask_count():  # C.8
    depth, fo, bb, lb, tw = synth.make_batch(1)
    out, _, _, _ = orc.estimate_ellipsoids(depth, fo, bb, lb, tw, synth.GROUND_PLANE_WORLD)
    x1, y1, x2, y2 = [int(v) for v in bb[0]]
    sub = depth[0][y1:y2:3, x1:x2:3].astype(np.int64)
    assert out[0].n_crop == int(((sub > 500) & (sub <= 30000)).sum())


def test_voxel_grid_lattice():  # C.9
    g = np.arange(0, 8) * 0.005 + 0.00125
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    out = orc.voxel_grid(pts, 0.01)
    assert len(out) == 64
    keys = np.floor(out / np.float32(0.01)).astype(int) @ np.array([1, 4, 16])
    assert (np.diff(keys) > 0).all()  # ascending voxel index, x fastest
    assert np.allclose(out[0], [0.00375, 0.00375, 0.00375], atol=1e-6)
    # unstable std::sort variant differs by float rounding only
    assert np.allclose(orc.voxel_grid(pts, 0.01, 1), out, atol=1e-7)


def test_euclidean_clustering_gap():  # C.10
    rng = np.random.default_rng(2)
    def blob(n, x0):
        g = np.stack(np.meshgrid(np.arange(12), np.arange(12), indexing="ij"), -1).reshape(-1, 2)[:n] * 0.01
        return np.c_[g[:, 0] + x0, g[:, 1], np.zeros(n)]
    a = blob(144, 0.0)
    for gap, nclu in ((0.021, 2), (0.019, 1)):
        b = blob(144, 0.11 + gap)
        k, lab = orc.euclidean_clusters(np.r_[a, b].astype(np.float32), 0.02, 100)
        assert k == nclu
    k, lab = orc.euclidean_clusters(np.r_[a, blob(50, 1.0)].astype(np.float32), 0.02, 100)
    assert k == 1 and (lab[144:] == -1).all()


def test_eigen3_and_ldlt():
    rng = np.random.default_rng(0)
    for _ in range(20):
        B = rng.normal(size=(3, 3)); A = B @ B.T
        ev, V = orc.eigen3(A)
        assert np.allclose(ev, np.linalg.eigvalsh(A), rtol=1e-12, atol=1e-14)
        assert np.allclose(A @ V, V * ev[None, :], atol=1e-12) and np.allclose(V.T @ V, np.eye(3), atol=1e-13)
    B = rng.normal(size=(18, 18)); A = B @ B.T + np.eye(18); b = rng.normal(size=18)
    ok, x = orc.ldlt_solve(A, b)
    assert ok and np.allclose(A @ x, b, atol=1e-9)
    ok, _ = orc.ldlt_solve(-A, b)
    assert not ok


def test_plane_analytic():
    depth, fo, bb, lb, tw = synth.make_batch(1)
    pl, found, ninl = orc.extract_ground_plane(depth)
    exp = orc.plane_to_world([0, 0, 1, 0.0], orc.se3_inverse(tw[0]))
    assert found[0] == 1 and ninl[0] > 20000
    assert np.allclose(pl[0], exp, atol=2e-3)
    # noise-free, hole-free frame: tight agreement
    rng = np.random.default_rng(0)
    R, t, pose7 = synth.camera_pose()
    d, _ = synth.render_depth(R, t, [], rng, noise=False, dropout=0.0)
    pl, found, _ = orc.extract_ground_plane(d)
    exp = orc.plane_to_world([0, 0, 1, 0.0], orc.se3_inverse(pose7))
    # pcl::computeMeanAndCovarianceMatrix accumulates 9 sums in FLOAT over ~250k pixels: the fit itself carries ~1e-3
    # error in the reference (hence the GPU path must reproduce the exact sequential float order)
    assert found[0] == 1 and np.allclose(pl[0], exp, atol=3e-3)


def test_calibrate_measurement():
    assert orc.calibrate_measurement([50, 50, 300, 300]) == 0
    assert orc.calibrate_measurement([5, 50, 300, 300]) == 1
    assert orc.calibrate_measurement([50, 50, 630, 300]) == 1


# ---- SLAM mode: free poses + odometry edges ----
def test_odometry_edge_is_zero_at_the_measured_relative_pose():
    rng = np.random.default_rng(3)
    v1 = orc.se3_exp(rng.normal(0, 0.5, 6)); v2 = orc.se3_exp(rng.normal(0, 0.5, 6))
    meas = orc.se3_mul(v2, orc.se3_inverse(v1))                      # Optimizer.cpp:141-143
    assert np.abs(orc.odom_error(meas, v1, v2)).max() < 1e-12        # log(C v1 v2^-1) = 0, types_six_dof_expmap.h:89-98
    xi = np.array([0.01, -0.02, 0.03, 0.1, 0.0, -0.05])
    e = orc.odom_error(orc.se3_mul(orc.se3_exp(xi), meas), v1, v2)   # left-perturbed measurement -> error = xi
    assert np.abs(e - xi).max() < 1e-9


def test_slam_mode_recovers_perturbed_poses():
    from helpers import make_slam_problem
    p = make_slam_problem(16, 5, 4, seed=11, pose_noise=0.01, noise=0.0)   # noise-free measurements: the truth is the optimum
    g = orc.Graph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], p["K"])
    e, T, st = orc.optimize_slam(g, p["init"], 10, odom_meas=p["odom"])
    assert st.chi2_final < 1e-9 * st.chi2_initial
    d1 = max(np.abs(orc.se3_log(orc.se3_mul(a, orc.se3_inverse(b)))).max() for a, b in zip(T[1:], p["Tcw_true"][1:]))
    assert d1 < 1e-7 and np.abs(e - p["truth"]).max() < 1e-7, d1
    assert np.array_equal(T[0], p["Tcw"][0])                         # first pose is fixed (Optimizer.cpp:137)


def test_slam_mode_with_fixed_poses_limit_equals_mapping_mode():
    """With a single frame there is no free pose and no odometry edge: SLAM mode must reproduce mapping mode exactly."""
    from helpers import make_ba_problem
    p = make_ba_problem(1, 3, 6, seed=5)
    g = orc.Graph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], p["K"])
    e0, s0 = orc.optimize_objects(g, p["init"], 5)
    e1, T, s1 = orc.optimize_slam(g, p["init"], 5)
    assert np.array_equal(e0, e1) and (s0.iterations, s0.trials_total) == (s1.iterations, s1.trials_total)


def test_slam_mode_noise_floor():
    """Documents why SLAM-mode state parity is asserted at 1e-3 (not 1e-4) on noisy graphs: the reference algorithm (numeric Jacobians
    with delta = 1e-9, weak odometry information) amplifies a 1e-13 measurement change -- the size of a libm ulp difference -- to
    ~1e-4 in the state after two iterations, while chi2 stays put.  Mapping mode is an order of magnitude better conditioned."""
    from helpers import make_slam_problem, K
    p = make_slam_problem(30, 6, 8, seed=7, pose_noise=0.01)
    g0 = orc.Graph(p["type"], p["frame"], p["object"], p["meas"], p["info"], p["Tcw"], p["n_objects"], K)
    m2 = p["meas"].copy(); m2[:, :4] += 1e-13 * np.sign(m2[:, :4])
    g1 = orc.Graph(p["type"], p["frame"], p["object"], m2, p["info"], p["Tcw"], p["n_objects"], K)
    e0, T0, s0 = orc.optimize_slam(g0, p["init"], 2, odom_meas=p["odom"])
    e1, T1, s1 = orc.optimize_slam(g1, p["init"], 2, odom_meas=p["odom"])
    assert abs(s1.chi2_final - s0.chi2_final) <= 1e-5 * s0.chi2_final
    assert np.abs(e1 - e0).max() < 1e-3 and np.abs(T1 - T0).max() < 1e-3        # bounded ...
    assert np.abs(e1 - e0).max() > 1e-6                                          # ... but far above the 1e-13 input change
    m0, _ = orc.optimize_objects(g0, p["init"], 2); m1, _ = orc.optimize_objects(g1, p["init"], 2)
    assert np.abs(m1 - m0).max() < np.abs(e1 - e0).max()


def test_oracle_never_reads_pixels_outside_the_boxes():
    """Restated reference behaviour the CUDA host entry relies on (it uploads only the boxes' union rectangle per frame): crop,
    centre samples and the range-image reads of the symmetry cost all stay inside the int-truncated bbox."""
    from eslam_b200 import synth
    from helpers import poison_outside_boxes
    depth, fo, bb, lb, tw = synth.make_batch(2, "one_box")
    a, _, _, _ = orc.estimate_ellipsoids(depth, fo, bb, lb, tw, synth.GROUND_PLANE_WORLD)
    b, _, _, _ = orc.estimate_ellipsoids(poison_outside_boxes(depth, fo, bb), fo, bb, lb, tw, synth.GROUND_PLANE_WORLD)
    assert bytes(a) == bytes(b) and all(o.status == 0 for o in a)
