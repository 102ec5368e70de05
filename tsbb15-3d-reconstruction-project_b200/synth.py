"""Synthetic correspondences of the shapes BASELINE.json names, plus access to the
Dino fixtures (data/dino_data.npz inside the package: INPUT data, not an oracle artefact — exported from the reference's BAdino2.mat / imgdata/points.txt by
oracle/gen_golden.py).  Pure numpy host code; nothing here is on the hot path.
"""
from __future__ import annotations

import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DINO_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "dino_data.npz")

# bounding box of the Dino model in world units
DINO_BBOX = np.array([[-0.045, 0.045], [-0.08, 0.03], [-0.72, -0.54]])

_dino_cache = None


def dino() -> dict:
    """{'Ps' (36,3,4), 'x2d' (36,2,676), 'X3d' (676,3), 'tracks' (4983,72), 'Fmatrix', 'clean_data_eval'}."""
    global _dino_cache
    if _dino_cache is None:
        with np.load(DINO_FIXTURE) as z:
            d = {k: z[k] for k in z.files}
        d["tracks"] = d.pop("tracks_x100").astype(np.float64) / 100.0
        _dino_cache = d
    return _dino_cache


def dino_clean_pair(i1: int, i2: int):
    """Clean (exact synthetic) correspondences of views i1, i2 as (N,2),(N,2) — correspondences.py:24-36."""
    x2d = dino()["x2d"]
    y1, y2 = x2d[i1].T, x2d[i2].T
    ok = np.logical_and(np.any(y1 != -1, axis=1), np.any(y2 != -1, axis=1))
    return np.array(y1[ok]), np.array(y2[ok])


def dino_noisy_pair(i1: int, i2: int):
    """Noisy tracked correspondences from imgdata/points.txt (the commented branch correspondences.py:19-22)."""
    tr = dino()["tracks"]
    y1, y2 = tr[:, 2 * i1:2 * i1 + 2], tr[:, 2 * i2:2 * i2 + 2]
    ok = np.logical_and(np.any(y1 != -1, axis=1), np.any(y2 != -1, axis=1))
    return np.array(y1[ok]), np.array(y2[ok])


def dino_view_2d3d(i: int):
    """Visible world points of view i and their C-normalised image points: X (n,3), y (n,2), K (3,3)."""
    d = dino()
    vis = np.any(d["x2d"][i] != -1, axis=0)
    X = d["X3d"][vis]
    px = d["x2d"][i][:, vis].T
    K = calibration(d["Ps"][i])
    yh = np.linalg.solve(K, np.vstack([px.T, np.ones(px.shape[0])])).T
    return X, yh[:, :2] / yh[:, 2:3], K


def calibration(P: np.ndarray) -> np.ndarray:
    """Upper-triangular K (positive diagonal, K[2,2] = 1) of a 3x4 camera by RQ of its left 3x3 block."""
    M = P[:, :3]
    # RQ via QR of the reversed transpose
    Q, R = np.linalg.qr(np.flipud(M).T)
    K = np.fliplr(np.flipud(R.T))
    D = np.diag(np.sign(np.diag(K)))
    K = K @ D
    return K / K[2, 2]


def project(P: np.ndarray, X: np.ndarray) -> np.ndarray:
    Xh = np.hstack([X, np.ones((X.shape[0], 1))])
    x = Xh @ P.T
    return x[:, :2] / x[:, 2:3]


def two_view(n: int, seed: int = 1, cams=(0, 1), outlier_frac: float = 0.3, sigma_px: float = 0.5,
             image_size=(640.0, 480.0)):
    """Config 3/5 pair: X ~ U(Dino bbox) seen by two Dino cameras, N(0, sigma^2) pixel noise in both images, the first
    ``outlier_frac`` of the image-2 points replaced by U(image).  Returns (N,4) rows (x0,x1,y0,y1) and the bool
    ground-truth inlier labels."""
    Ps = dino()["Ps"]
    rng = np.random.default_rng(seed)
    X = rng.uniform(DINO_BBOX[:, 0], DINO_BBOX[:, 1], size=(n, 3))
    x = project(Ps[cams[0]], X) + rng.normal(0.0, sigma_px, size=(n, 2))
    y = project(Ps[cams[1]], X) + rng.normal(0.0, sigma_px, size=(n, 2))
    n_out = int(round(outlier_frac * n))
    y[:n_out] = rng.uniform([0.0, 0.0], image_size, size=(n_out, 2))
    labels = np.ones(n, dtype=bool)
    labels[:n_out] = False
    return np.ascontiguousarray(np.hstack([x, y])), labels


def multi_pair(n_pairs: int, n: int, first_pair: int = 0):
    """Config 5: pair p uses seed 1000+p and a random pair of distinct Dino cameras drawn from that seed."""
    out = []
    for p in range(first_pair, first_pair + n_pairs):
        rng = np.random.default_rng(1000 + p)
        c = rng.choice(36, size=2, replace=False)
        pts, _ = two_view(n, seed=1000 + p, cams=(int(c[0]), int(c[1])))
        out.append(pts)
    return out


def pnp_scene(n: int, seed: int = 4, cam: int = 17, outlier_frac: float = 0.3, sigma_px: float = 0.5):
    """Config 4: X ~ U(Dino bbox), C-normalised image points of Dino camera ``cam`` with noise sigma_px / f, the first
    ``outlier_frac`` image points replaced by uniform points of the normalised image.  Returns X (N,3), y (N,2), (R,t)."""
    P = dino()["Ps"][cam]
    K = calibration(P)
    Rt = np.linalg.solve(K, P)
    if np.linalg.det(Rt[:, :3]) < 0:
        Rt = -Rt
    s = np.cbrt(np.linalg.det(Rt[:, :3]))
    Rt = Rt / s
    rng = np.random.default_rng(seed)
    X = rng.uniform(DINO_BBOX[:, 0], DINO_BBOX[:, 1], size=(n, 3))
    yc = X @ Rt[:, :3].T + Rt[:, 3]
    y = yc[:, :2] / yc[:, 2:3] + rng.normal(0.0, sigma_px / K[0, 0], size=(n, 2))
    n_out = int(round(outlier_frac * n))
    lo, hi = y[n_out:].min(axis=0), y[n_out:].max(axis=0)
    y[:n_out] = rng.uniform(lo, hi, size=(n_out, 2))
    return np.ascontiguousarray(X), np.ascontiguousarray(y), (Rt[:, :3], Rt[:, 3])


def ba_scene(n_views: int, n_points: int, seed: int = 7, track: int = 8, sigma: float = 0.5 / 3217.0, start_pt: float = 2e-3,
             start_cam: float = 1e-3):
    """A bundle-adjustment problem in the layout Tables.BundleAdjustment2 works on (tables.py:298-315): n_views cameras
    [R | t] on a ring around a point cloud (first view = [I | 0]), every point seen by `track` consecutive views
    (wrapping), C-normalised observations with noise sigma, start values off the truth by start_pt / start_cam.
    Returns cams (V,3,4), pts (P,3), uv (O,2), cam_idx (O,), pt_idx (O,), and the noise-free truth (cams, pts)."""
    rng = np.random.default_rng(seed)
    track = min(track, n_views)
    X = rng.uniform(-1.0, 1.0, (n_points, 3)) * np.array([1.0, 0.6, 1.0])
    cams_w = np.zeros((n_views, 3, 4))
    for k in range(n_views):
        a = 2.0 * np.pi * k / n_views
        c = np.array([6.0 * np.sin(a), 0.3 * np.sin(3 * a), -6.0 * np.cos(a)])        # camera centre on a ring of radius 6
        z = -c / np.linalg.norm(c)
        x = np.cross([0.0, 1.0, 0.0], z)
        x /= np.linalg.norm(x)
        R = np.stack([x, np.cross(z, x), z])
        cams_w[k, :, :3] = R
        cams_w[k, :, 3] = -R @ c
    R0, t0 = cams_w[0, :, :3], cams_w[0, :, 3]
    Xc = X @ R0.T + t0                                                               # frame of view 0
    true_cams = np.zeros_like(cams_w)
    for k in range(n_views):
        Rk = cams_w[k, :, :3] @ R0.T
        true_cams[k, :, :3] = Rk
        true_cams[k, :, 3] = cams_w[k, :, 3] - Rk @ t0
    first = rng.integers(0, n_views, n_points)
    cam_idx = ((first[:, None] + np.arange(track)[None, :]) % n_views).ravel()
    pt_idx = np.repeat(np.arange(n_points), track)
    Xh = np.hstack([Xc, np.ones((n_points, 1))])
    y = np.einsum("oab,ob->oa", true_cams[cam_idx], Xh[pt_idx])
    uv = y[:, :2] / y[:, 2:] + sigma * rng.standard_normal((len(cam_idx), 2))
    cams0 = true_cams.copy()
    cams0[1:] += start_cam * rng.standard_normal((n_views - 1, 3, 4)) * np.array([1, 1, 1, 0.1])
    pts0 = Xc + start_pt * rng.standard_normal(Xc.shape)
    return cams0, pts0, uv, cam_idx.astype(np.int32), pt_idx.astype(np.int32), (true_cams, Xc)
