"""The reference's incremental pipeline (main.py:34-128: INIT1-3, EXT1-5; the second test adds the BA step) replayed on the
exact synthetic Dino tracks through the drop-in modules: F-RANSAC -> E -> relative pose -> triangulation -> for every
further view the 2D<->3D match loop + PnP-RANSAC of Tables.addNewView + triangulation of the new points.  Everything
numeric runs on the GPU; the data are exact, so every recovered pose must equal the ground-truth camera (expressed in
the frame of camera 0 with unit baseline) and every 3-D point the ground-truth point."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(dino, i, j):
    y1, y2 = dino["x2d"][i].T, dino["x2d"][j].T
    ok = np.logical_and(np.any(y1 != -1, axis=1), np.any(y2 != -1, axis=1))
    return np.ascontiguousarray(y1[ok]), np.ascontiguousarray(y2[ok]), np.flatnonzero(ok)


def test_main_py_chain_on_clean_dino(rg, dino, pnp_golden):
    fun, lab3 = rg.fun, rg.lab3
    Tables, CameraPose = rg.tables.Tables, rg.help_classes.CameraPose
    Ps = dino["Ps"]
    C = Ps[None]                                              # fun.getCameraMatrices() layout (1, 36, 3, 4)
    y1, y2, _ = _pair(dino, 0, 1)
    # INIT1 (main.py:38-42): F by RANSAC (the gold-standard LM stage is SciPy, kept off here)
    F = fun.getFFromLabCode(y1.T, y2.T, r=2000, seed=0, refine=False)
    # INIT2 (main.py:54-62)
    E, K = fun.getEAndK(C, F)
    assert np.allclose(K, pnp_golden["K"][-1], rtol=1e-9)
    T = Tables()
    T.K = K
    y1h, y2h = fun.MakeHomogenous(K, y1), fun.MakeHomogenous(K, y2)
    R, t = fun.relative_camera_pose(E, y1h[0, :2].T, y2h[0, :2].T)
    Rg, tg = pnp_golden["R"], pnp_golden["t"]
    R01 = Rg[1] @ Rg[0].T
    t01 = tg[1] - R01 @ tg[0]
    s = np.linalg.norm(t01)
    # the Dino cameras (K with positive diagonal) see the object at NEGATIVE depth, so the cheirality test of
    # fun.py:240-255 selects the point-reflected twin (R, -t, -X): same images, z > 0.  sg carries that sign.
    sg = np.sign(np.median((dino["X3d"] @ Rg[0].T + tg[0])[:, 2]))
    s = s * sg
    assert np.abs(R - R01).max() < 1e-7 and np.abs(t - t01 / s).max() < 1e-7
    # INIT3 (main.py:68-76)
    C1, C2 = CameraPose(), CameraPose(R, t)
    v1, v2 = T.addView(0, C1), T.addView(1, C2)
    T.triangulateAndAddPoints(v1, v2, C1, C2, y1h, y2h)
    assert T.T_points.size == len(y1) and T.T_obs.size == 2 * len(y1)
    # EXT1-5 (main.py:111-128) for the next views
    for i in range(1, 6):
        a, b, _ = _pair(dino, i, i + 1)
        ah, bh = fun.MakeHomogenous(K, a), fun.MakeHomogenous(K, b)
        n_obs_before = T.T_obs.size
        A1, A2 = T.addNewView(K, i + 1, ah, bh, a, b, r=256, reproj_px=1.5, seed=i)
        view = T.T_views[-1]
        Rk = Rg[i + 1] @ Rg[0].T
        tk = (tg[i + 1] - Rk @ tg[0]) / s
        # no bundle adjustment in between: the DLT poses inherit the error of the points they were computed from, so
        # the tolerance is the drift of an incremental chain on exact data, not a kernel tolerance
        assert np.abs(view.camera_pose.R - Rk).max() < 1e-5, i
        assert np.abs(view.camera_pose.t - tk).max() < 1e-5 * np.abs(tk).max(), i
        assert T.T_obs.size > n_obs_before and len(A1) == len(A2)
        added = T.addNewPoints(fun.MakeHomogenous(K, A1), fun.MakeHomogenous(K, A2), i, i + 1)
        assert added == len(A1)                               # exact data: every putative pair satisfies E
    # every reconstructed point is a ground-truth point in the camera-0 frame, unit-baseline scale
    Xgt = (dino["X3d"] @ Rg[0].T + tg[0]) / s
    rec = np.array([p.point for p in T.T_points])
    d = np.abs(rec[:, None, :] - Xgt[None, :, :]).max(axis=2).min(axis=1)
    assert d.max() < 1e-5 * np.abs(Xgt).max()
    Rs, ts = T.getCamerasForEvaluation()
    assert Rs.shape == (7, 3, 3) and ts.shape == (7, 3)


def test_main_py_chain_with_bundle_adjustment(rg, dino, pnp_golden):
    """main.py:95-128 in the reference's order — BundleAdjustment2, then addNewView / addNewPoints — on the exact Dino
    tracks.  Without BA the chain drifts (1e-5, test above); with BA after every view the total reprojection cost of the
    tables goes back to rounding level each time, and a perturbed reconstruction is pulled back onto the observations."""
    from oracle import ba_path as oba
    fun = rg.fun
    Tables, CameraPose = rg.tables.Tables, rg.help_classes.CameraPose
    y1, y2, _ = _pair(dino, 0, 1)
    F = fun.getFFromLabCode(y1.T, y2.T, r=2000, seed=0, refine=False)
    E, K = fun.getEAndK(dino["Ps"][None], F)
    T = Tables()
    T.K = K
    y1h, y2h = fun.MakeHomogenous(K, y1), fun.MakeHomogenous(K, y2)
    R, t = fun.relative_camera_pose(E, y1h[0, :2].T, y2h[0, :2].T)
    C1, C2 = CameraPose(), CameraPose(R, t)
    v1, v2 = T.addView(0, C1), T.addView(1, C2)
    T.triangulateAndAddPoints(v1, v2, C1, C2, y1h, y2h)

    def table_cost():
        cams = np.stack([v.camera_pose.GetCameraMatrix() for v in T.T_views])
        pts = np.stack([p.point for p in T.T_points])
        uv, ci, pi = T.observationArrays()
        return oba.cost(cams, pts, uv, ci, pi)

    for i in range(1, 5):
        before = table_cost()
        info = T.BundleAdjustment2()                          # main.py:100
        after = table_cost()
        assert abs(after - info["cost"]) <= 1e-9 * max(after, 1e-30) + 1e-24
        assert after <= before and after < 1e-18              # exact observations: the drift of the chain is removed
        assert np.array_equal(T.T_views[0].camera_pose.GetCameraMatrix(), np.hstack([np.eye(3), np.zeros((3, 1))]))
        a, b, _ = _pair(dino, i, i + 1)
        ah, bh = fun.MakeHomogenous(K, a), fun.MakeHomogenous(K, b)
        A1, A2 = T.addNewView(K, i + 1, ah, bh, a, b, r=256, reproj_px=1.5, seed=i)
        T.addNewPoints(fun.MakeHomogenous(K, A1), fun.MakeHomogenous(K, A2), i, i + 1)
    # a perturbed reconstruction (3-D points off by 1 % of the scene depth) is pulled back onto the exact observations
    rng = np.random.default_rng(5)
    scale = np.abs(np.stack([p.point for p in T.T_points])).max()
    for p in T.T_points:
        p.point = p.point + 0.01 * scale * rng.standard_normal(3)
    c0 = table_cost()
    info = T.BundleAdjustment2(max_iter=100, ftol=1e-12)
    assert info["cost"] < 1e-10 * c0 and info["status"] in (2, 3)


def test_whole_main_py_sequence_with_bundle_adjustment(rg, dino, pnp_golden):
    """tools/run_main_dropin.py = main.py:30-181 statement by statement: all 35 views, BA before every new view.  On the
    exact tracks the reconstruction must reproject onto every observation, and the cameras must be the ground-truth
    cameras up to the projective gauge BA leaves free (the reference's own R_eval_clean.npy is not orthonormal either):
    checked through the reprojection of the GROUND-TRUTH points mapped by the best 3-D homography."""
    import importlib.util
    import os
    from oracle import ba_path as oba
    spec = importlib.util.spec_from_file_location(
        "run_main_dropin", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "run_main_dropin.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    T, Rs, ts, times, total, ba_log = mod.run(34, bundle_adjust=True, r_f=2000, verbose=False)
    assert Rs.shape == (35, 3, 3) and ts.shape == (35, 3) and len(ba_log) == 33
    cams = np.stack([v.camera_pose.GetCameraMatrix() for v in T.T_views])
    pts = np.stack([p.point for p in T.T_points])
    uv, ci, pi = T.observationArrays()
    rms = np.sqrt(2 * oba.cost(cams, pts, uv, ci, pi) / (2 * len(uv)))
    assert rms < 1e-7                                        # C-normalised units: 3e-4 px
    assert np.array_equal(cams[0], np.hstack([np.eye(3), np.zeros((3, 1))]))
    # every view's observations carry the same image points as the ground-truth tracks of that view
    assert len(T.T_points) > 600 and len(uv) > 3500
    for it in ba_log:
        assert it[4] < 1e-12                                 # every BA converged onto the exact observations
    # all 35 recovered rotations are the ground-truth rotations (relative to view 0); without BA the chain drifts until
    # PnP-RANSAC finds no consensus any more around view 20
    assert mod.rotation_errors_deg(Rs).max() < 1e-4


def test_match_last_view_equals_reference_loop(rg, dino, pnp_golden):
    """Tables.matchLastView against the literal double loop of tables.py:116-124 (numpy oracle)."""
    from oracle import geom_path as og
    fun = rg.fun
    Tables, CameraPose = rg.tables.Tables, rg.help_classes.CameraPose
    K = pnp_golden["K"][0]
    T = Tables(); T.K = K
    y1, y2, _ = _pair(dino, 0, 1)
    y1h, y2h = fun.MakeHomogenous(K, y1), fun.MakeHomogenous(K, y2)
    v1, v2 = T.addView(0, CameraPose()), T.addView(1, CameraPose())
    for k, (a, b) in enumerate(zip(y1h, y2h)):
        p = T.addPoint(np.array([k, 0.0, 1.0]))
        T.addObs(a, v1, p); T.addObs(b, v2, p)
    q, _, _ = _pair(dino, 1, 2)
    qh = fun.MakeHomogenous(K, q)
    got = T.matchLastView(qh)
    obs_idx = T.T_views[-1].observations_index
    coords = np.array([T.T_obs[v].image_coordinates for v in obs_idx])
    ref = og.match_first_within(coords, qh, 1e-4)
    ref = np.where(ref >= 0, obs_idx[np.maximum(ref, 0)], -1)
    assert np.array_equal(got, ref) and (got >= 0).sum() > 10


def test_two_view_init_batched_equals_oracle_chain(rg, dino, pnp_golden):
    """rg_two_view_init (main.py:54-76 for P pairs in one call) against the same chain spelled out with the oracle:
    E = K^T F K, MakeHomogenous, relative_camera_pose on the first correspondence, triangulate_optimal per point."""
    from oracle import geom_path as og
    K = pnp_golden["K"][0]
    Kinv = np.linalg.inv(K)
    Ps = dino["Ps"]
    rng = np.random.default_rng(3)
    pairs, Fs = [], []
    for (i, j, noise) in ((0, 1, 0.0), (4, 5, 0.0), (10, 11, 0.4), (20, 22, 0.0), (30, 31, 1.0)):
        a, b, _ = _pair(dino, i, j)
        pairs.append(np.hstack([a + rng.normal(0, noise, a.shape), b + rng.normal(0, noise, b.shape)]))
        Fs.append(og.fmatrix_from_cameras(Ps[i], Ps[j]))
    pairs.insert(2, np.zeros((0, 4))); Fs.insert(2, Fs[0])          # an empty pair in the middle of the batch
    res = rg.batched.two_view_init(pairs, np.stack(Fs), K)
    for p, (pts, F) in enumerate(zip(pairs, Fs)):
        if len(pts) == 0:
            assert res["which"][p] == -1 and res["X"][p].shape == (0, 3)
            continue
        h1 = (Kinv @ np.column_stack([pts[:, :2], np.ones(len(pts))]).T).T
        h2 = (Kinv @ np.column_stack([pts[:, 2:], np.ones(len(pts))]).T).T
        E = K.T @ F @ K
        R, t = og.relative_camera_pose(E, h1[0, :2], h2[0, :2])
        assert np.abs(res["R"][p] - R).max() < 1e-9 and np.abs(res["t"][p] - t).max() < 1e-9, p
        C1 = np.hstack([np.eye(3), np.zeros((3, 1))]); C2 = np.hstack([R, t[:, None]])
        X = og.triangulate_optimal_batch(C1, C2, h1[:, :2], h2[:, :2])
        err = np.abs(res["X"][p] - X).max(axis=1) / np.abs(X).max()
        assert np.mean(err < 1e-8) >= 0.98 and np.median(err) < 1e-10, (p, np.sort(err)[-3:])


def test_ransac_then_two_view_init_with_masks(rg, dino, pnp_golden):
    """F-RANSAC winners and inlier masks fed straight into the batched initialisation: outliers come back as NaN, inliers
    reproject onto their measurements."""
    from oracle import geom_path as og
    K = pnp_golden["K"][0]
    rng = np.random.default_rng(8)
    pairs = []
    for i in (0, 7, 15):
        a, b, _ = _pair(dino, i, i + 1)
        pts = np.hstack([a, b]) + rng.normal(0, 0.2, (len(a), 4))
        pts[::5, 2:] = rng.uniform(0, 600, (len(pts[::5]), 2))            # 20 % gross outliers
        pairs.append(pts)
    fr = rg.batched.f_ransac_pairs(pairs, n_hyp=2000, thr=1.5, seed=1, host_sampling=True)
    res = rg.batched.two_view_init(pairs, fr["F"], K, masks=fr["mask"])
    Kinv = np.linalg.inv(K)
    for p, pts in enumerate(pairs):
        m = fr["mask"][p].astype(bool)
        assert res["which"][p] >= 0 and m.sum() > 0.6 * len(pts)
        X = res["X"][p]
        assert np.isnan(X[~m]).all() and np.isfinite(X[m]).all()
        R, t = res["R"][p], res["t"][p]
        y1 = (Kinv @ np.column_stack([pts[m, :2], np.ones(m.sum())]).T).T[:, :2]
        y2 = (Kinv @ np.column_stack([pts[m, 2:], np.ones(m.sum())]).T).T[:, :2]
        # (no reprojection bound: with its f = f' = 1 simplification the reference's "optimal" correction moves noisy
        #  C-normalised points by up to ~1e-2; the contract is equality with the reference algorithm, checked here)
        C1 = np.hstack([np.eye(3), np.zeros((3, 1))]); C2 = np.hstack([R, t[:, None]])
        sub = np.arange(0, m.sum(), 7)
        ref = og.triangulate_optimal_batch(C1, C2, y1[sub], y2[sub])
        err = np.abs(X[m][sub] - ref).max(axis=1) / np.abs(ref).max()
        assert np.mean(err < 1e-8) >= 0.95 and np.median(err) < 1e-10
        x2 = X[m] @ R.T + t
        assert (X[m][:, 2] > 0).mean() > 0.95 and (x2[:, 2] > 0).mean() > 0.95
