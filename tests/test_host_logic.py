"""Host-side logic that needs no GPU: samplers, helper formulas, argument validation, synthetic data, and a numpy
float32 model of the scorer's guard band (the device code's bound must dominate the observed FP32 error)."""
import numpy as np
import pytest

from oracle import f_path as orc


def test_fast_sampler_rows_are_distinct_and_reproducible(rg):
    s = rg.sampling
    a = s.fast(37, 5000, 8, seed=3)
    b = s.fast(37, 5000, 8, seed=3)
    assert a.dtype == np.int32 and a.shape == (5000, 8)
    assert np.array_equal(a, b)
    assert (np.sort(a, axis=1)[:, 1:] != np.sort(a, axis=1)[:, :-1]).all()
    assert a.min() >= 0 and a.max() < 37
    assert not np.array_equal(a, s.fast(37, 5000, 8, seed=4))
    assert s.fast(8, 10, 8, seed=0).shape == (10, 8)           # N == k still works
    with pytest.raises(ValueError):
        s.fast(7, 10, 8)


def test_reference_stream_equals_the_reference_draw(rg):
    np.random.seed(11)
    a = rg.sampling.reference_stream(257, 50, 8)
    np.random.seed(11)
    b = np.stack([np.random.choice(np.arange(0, 257, 1), 8, replace=False) for _ in range(50)])
    assert np.array_equal(a, b)


def test_golden_indices_are_the_seed0_reference_stream(rg, f_golden, noisy01):
    np.random.seed(0)
    a = rg.sampling.reference_stream(noisy01[0].shape[1], 64, 8)
    assert np.array_equal(a, f_golden["ransac_idx"][:64])


def test_ransac_helpers(rg):
    r = rg.ransac
    assert r.calc_p(0.5, 3, 10) == pytest.approx(1 - (1 - 0.125) ** 10)
    assert r.calc_r(0.5, 3, 0.99) == pytest.approx(np.log(0.01) / np.log(1 - 0.125))
    assert r.norm_p([2.0, 4.0, 2.0]) == [1.0, 2.0, 1.0]
    assert r.cart([2.0, 4.0, 2.0]) == [1.0, 2.0]
    assert r.dpp([1, 2, 1], [2, 2, 2]) == pytest.approx(1.0)
    assert r.dpp_squared([1, 2, 1], [4, 2, 2]) == pytest.approx(2.0)
    assert np.allclose(r.calc_y_prim(np.array([1.0, 0, 0]), np.eye(3), np.array([0, 0, 1.0])), [1, 0, 1])
    with pytest.raises(ValueError):
        r.gen_rnd_indices(3, 6)
    assert sorted(r.gen_rnd_indices(6, 6)) == list(range(6))


def test_ransac_robust_argument_errors(rg):
    r = rg.ransac
    D = np.zeros((10, 2, 3))
    with pytest.raises(ValueError, match="Not implemented yet"):
        r.ransac_robust(D, D, 5, 1e-3, 4)
    with pytest.raises(ValueError, match="No PnP algorithm"):
        r.ransac_robust(D, D, 5, 1e-3, 5)
    with pytest.raises(ValueError):
        r.ransac_robust(D, D, 5, 1e-3, 3)                  # the reference's p3p raises too
    with pytest.raises(ValueError):
        rg.pnp.p3p(None, None, None)
    with pytest.raises(ValueError):
        rg.pnp.pnp_minimize(np.ones((5, 4)), np.ones((5, 3)), 5)


def test_lab3_shape_errors_before_touching_the_gpu(rg):
    with pytest.raises(ValueError, match="same shape"):
        rg.lab3.fmatrix_stls(np.zeros((2, 8)), np.zeros((2, 9)))
    with pytest.raises(ValueError, match="same sizes"):
        rg.lab3.fmatrix_residuals(np.eye(3), np.zeros((2, 8)), np.zeros((2, 9)))
    with pytest.raises(ValueError):
        rg.fun.f_ransac(np.zeros((2, 8)), np.zeros((2, 9)))


def test_synthetic_scenes(rg):
    pts, labels = rg.synth.two_view(2000, seed=1)
    assert pts.shape == (2000, 4) and labels.sum() == 1400
    F = orc.fmatrix_stls(pts[labels][:200, :2].T, pts[labels][:200, 2:].T)
    d = orc.distance(F, pts[:, :2].T, pts[:, 2:].T)
    assert np.mean(d[labels] < 1.5) > 0.9 and np.mean(d[~labels] < 1.5) < 0.1
    X, y, (R, t) = rg.synth.pnp_scene(1000, seed=2)
    e = np.sum((y - (X @ R.T + t)[:, :2] / (X @ R.T + t)[:, 2:3]) ** 2, axis=1)
    assert np.mean(e[300:] < (1.5 / 3217) ** 2) > 0.9 and np.mean(e[:300] < (1.5 / 3217) ** 2) < 0.05
    assert abs(np.linalg.det(R) - 1) < 1e-9
    assert len(rg.synth.multi_pair(2, 100)) == 2
    y1, y2 = rg.synth.dino_noisy_pair(0, 1)
    assert y1.shape == (257, 2)
    assert [rg.synth.dino_clean_pair(i, i + 1)[0].shape[0] for i in range(3)] == [37, 52, 70]


# ---- guard band model ------------------------------------------------------------------------------------------
def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _hyp32_model(F, c1, c2, thr, B):
    """numpy transcription of make_hyp32 (csrc/f_kernels.cuh) for the EPI_MAX mode."""
    M1 = np.array([[thr, 0, c1[0]], [0, thr, c1[1]], [0, 0, 1.0]])
    M2 = np.array([[thr, 0, c2[0]], [0, thr, c2[1]], [0, 0, 1.0]])
    Ft = M1.T @ F @ M2
    w = np.array([B, B, 1.0])
    Ft = Ft / np.sum(np.abs(Ft) * np.outer(w, w))
    a = np.abs(Ft)
    rho = a @ w                      # rho0, rho1 (rows), w-weighted
    # image-2 normal l2 = (f0 x0 + f3 x1 + f6, f1 x0 + f4 x1 + f7) rotated so that its second component loses the x0 term
    f = Ft.ravel()
    hh = np.hypot(f[0], f[1])
    if hh > 0:
        g = np.array([hh, (f[0] * f[3] + f[1] * f[4]) / hh, (f[0] * f[6] + f[1] * f[7]) / hh,
                      (f[0] * f[4] - f[1] * f[3]) / hh, (f[0] * f[7] - f[1] * f[6]) / hh])
    else:
        g = np.array([0.0, f[3], f[6], f[4], f[7]])
    kap = np.array([abs(g[0]) * B + abs(g[1]) * B + abs(g[2]), abs(g[3]) * B + abs(g[4])])
    S1, S2 = rho[0] ** 2 + rho[1] ** 2, kap[0] ** 2 + kap[1] ** 2
    eps = 2.0 ** -24
    Mb, Smax = min(S1, S2), max(S1, S2)
    G = 1.25 * 16 * eps * np.sqrt(Mb) + 2 * (100 * eps * eps + 10 * eps * Smax + 2 * eps * Mb)
    return Ft, g, G


def test_guard_band_dominates_fp32_error(rg):
    """|q32 - q_exact| <= G for every evaluation whose decision is in doubt; the observed error stays far below G."""
    pts, _ = rg.synth.two_view(4000, seed=21)
    p1, p2 = pts[:, :2].T, pts[:, 2:].T
    idx = rg.sampling.fast(4000, 40, 8, seed=1)
    thr = 1.5
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    c = 0.5 * (lo + hi)
    B = np.max(hi - c) / thr * (1 + 1e-6)
    xt = ((pts[:, :2] - c[:2]) / thr)
    yt = ((pts[:, 2:] - c[2:]) / thr)
    x32, y32 = xt.astype(np.float32), yt.astype(np.float32)
    worst = 0.0
    for sel in idx:
        F = orc.fmatrix_stls(p1[:, sel], p2[:, sel])
        Ft, g64, G = _hyp32_model(F, c[:2], c[2:], thr, B)
        f = Ft.astype(np.float32).ravel()
        g = g64.astype(np.float32)
        x0, x1, y0, y1 = x32[:, 0], x32[:, 1], y32[:, 0], y32[:, 1]
        bc = lambda v: np.full_like(x0, v)
        l1x = _fma32(bc(f[0]), y0, _fma32(bc(f[1]), y1, bc(f[2])))
        l1y = _fma32(bc(f[3]), y0, _fma32(bc(f[4]), y1, bc(f[5])))
        l1z = _fma32(bc(f[6]), y0, _fma32(bc(f[7]), y1, bc(f[8])))
        r = _fma32(l1x, x0, _fma32(l1y, x1, l1z))
        l2x = _fma32(bc(g[0]), x0, _fma32(bc(g[1]), x1, bc(g[2])))       # rotated normal: |l2'| = |l2|, one FMA fewer
        l2y = _fma32(bc(g[3]), x1, bc(g[4]))
        s1 = _fma32(l1x, l1x, (l1y * l1y).astype(np.float32))
        s2 = _fma32(l2x, l2x, (l2y * l2y).astype(np.float32))
        q32 = _fma32(r, r, -np.minimum(s1, s2)).astype(np.float64)
        # exact value in the same frame, float64 (error ~1e-16, negligible against 2^-24)
        xh = np.column_stack([xt, np.ones(len(xt))])
        yh = np.column_stack([yt, np.ones(len(yt))])
        L1 = yh @ Ft.T
        L2 = xh @ Ft
        rr = np.sum(L1 * xh, axis=1)
        q = rr * rr - np.minimum(L1[:, 0] ** 2 + L1[:, 1] ** 2, L2[:, 0] ** 2 + L2[:, 1] ** 2)
        doubt = np.sign(q32) != np.sign(q)
        assert np.all(np.abs(q32[doubt]) <= G), "a wrong FP32 decision escaped the guard band"
        near = np.abs(q) <= 4 * G
        if near.any():
            worst = max(worst, np.max(np.abs(q32[near] - q[near])) / G)
        # and the decision in that frame is the reference decision
        d = orc.distance(F, p1, p2)
        assert np.array_equal(q < 0, d < thr) or np.count_nonzero((q < 0) != (d < thr)) <= 1
    assert worst < 1.0


def test_tables_bookkeeping_matches_reference_semantics(rg):
    """tables.Tables / help_classes mirrors: the table bookkeeping is pure host logic (no GPU)."""
    Tables = rg.tables.Tables
    hc = rg.help_classes
    T = Tables()
    T.K = np.eye(3)
    v0 = T.addView(0, hc.CameraPose())
    v1 = T.addView(1, hc.CameraPose(np.eye(3), np.array([1.0, 0.0, 0.0])))
    assert (v0, v1) == (0, 1) and T.T_views.size == 2
    # the reference initialises observations_index to [0] (help_classes.py:27, 62): kept
    assert T.T_views[0].observations_index.tolist() == [0]
    p = T.addPoint(np.array([0.0, 0.0, 5.0]))
    T.addObs(np.array([0.0, 0.0, 1.0]), v0, p)
    T.addObs(np.array([0.2, 0.0, 1.0]), v1, p)
    assert T.T_obs.size == 2 and T.T_obs[1].view_index == 1 and T.T_obs[1].color is None
    assert T.T_views[1].observations_index.tolist() == [0, 1] and T.T_points[0].observations_index.tolist() == [0, 0, 1]
    yij, Rktk, xj = T.getObsAsArrays()
    assert yij.shape == (2, 3) and Rktk.shape == (2, 3, 4) and xj.shape == (2, 4) and xj[0, 3] == 1.0
    assert np.array_equal(Rktk[1], np.hstack([np.eye(3), [[1.0], [0.0], [0.0]]]))
    Rs, ts = T.getCamerasForEvaluation()
    assert Rs.shape == (2, 3, 3) and np.array_equal(ts[1], [1.0, 0.0, 0.0])
    assert np.allclose(T.T_views[1].getWorldPosition(), [-1.0, 0.0, 0.0])
    C = hc.CameraPose().GetCameraMatrix()
    assert C.shape == (3, 4) and np.array_equal(C[:, :3], np.eye(3))
    # fun.getEFromCameras / crossProductMat are host helpers
    E = rg.fun.getEFromCameras(T.T_views[0].camera_pose, T.T_views[1].camera_pose)
    assert np.allclose(E, rg.fun.crossProductMat(np.array([1.0, 0.0, 0.0])))


def test_fast_batch_is_the_per_pair_sampler_in_one_buffer(rg):
    """sampling.fast_batch: same index sets as fast(n_p, H, k, seed + p), stored as consecutive row blocks of one array, which
    runtime._consecutive recognises (no concatenation on the way to the library); anything else falls back to a copy."""
    import numpy as np
    from tsbb15_b200 import runtime as rt, sampling
    ns = [300, 257, 40, 9]
    blocks = sampling.fast_batch(ns, 500, 8, seed=11)
    for p, n in enumerate(ns):
        assert blocks[p].dtype == np.int32 and np.array_equal(blocks[p], sampling.fast(n, 500, 8, 11 + p))
    whole = rt._consecutive(blocks)
    assert whole is not None and whole.shape == (2000, 8) and np.array_equal(whole, np.concatenate(blocks))
    assert whole.ctypes.data == blocks[0].ctypes.data                       # a view of the same memory, not a copy
    assert rt._consecutive([blocks[0], blocks[2]]) is None                  # a gap
    assert rt._consecutive([blocks[1], blocks[0]]) is None                  # wrong order
    assert rt._consecutive([blocks[0], blocks[1].astype(np.int64)]) is None
    assert rt._consecutive([blocks[0][:, :4], blocks[1][:, :4]]) is None    # not contiguous
    assert rt._consecutive([]) is None
