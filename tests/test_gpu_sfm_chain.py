"""The reference's incremental pipeline (main.py:34-128: INIT1-3, EXT1-5, without bundle adjustment) replayed on the
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
