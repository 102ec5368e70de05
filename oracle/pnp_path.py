"""TEST INFRASTRUCTURE ONLY — CPU (numpy, float64) restatement of the reference's PnP path.

Status of the reference code for this path (verified in the build container):
  * ``pnp.pnp_minimize`` (pnp.py:163-196) is unfinished: it builds a list ``A`` and returns undefined names;
    the algorithm is specified only by the docstring outline pnp.py:132-152, which is what is restated here.
  * ``ransac.ransac_robust`` (ransac.py:37-113) cannot run (wrong arity at :82, ``len(D_high[0])`` at :77,
    ``C_est[0]`` on an empty list at :108); its *semantics* (sample from D_high, score D_med and D_high with the
    squared distance of p-normalised points, inclusive ``thresh >= e``, keep the largest consensus) are restated.
Because the reference cannot execute, there is no reference output to compare with: PARITY FOR THE PnP ROWS IS PINNED
ONLY BY the exact synthetic Dino data (``BAdino2.mat``: reprojection residual <= 3.4e-13 px), for which any correct
DLT-PnP must return the ground-truth pose that the reference's own ``fun.camera_resectioning`` (fun.py:260-280)
extracts from the same camera matrices (golden ``pnp_R`` / ``pnp_t`` in tests/golden/).

Outline restated (pnp.py:132-152):
   1-6   for every correspondence k and every row r_l of [y_k]_x : append vec(r_l x_k^T) (3x4, row-major) to A
   7-9   c0 = right singular vector of the smallest singular value of A ("homogeneous method"), C0 = reshape 3x4 = (A|b)
   10-13 tau = sign(det A);  tau*A = U S V^T;  R = U V^T;  lambda = 3 tau / trace(S);  t = lambda b
"""
from __future__ import annotations

import numpy as np


def cross_matrix(v: np.ndarray) -> np.ndarray:
    return np.array([[0.0, -v[2], v[1]],
                     [v[2], 0.0, -v[0]],
                     [-v[1], v[0], 0.0]])


def design_matrix(X_h: np.ndarray, y_h: np.ndarray) -> np.ndarray:
    """(3m, 12) matrix of pnp.py:138-143.  X_h (m,4) homogeneous world points, y_h (m,3) C-normalised homogeneous."""
    m = X_h.shape[0]
    A = np.empty((3 * m, 12))
    for k in range(m):
        Yx = cross_matrix(y_h[k])
        for l in range(3):
            A[3 * k + l] = np.outer(Yx[l], X_h[k]).ravel()
    return A


def pnp_minimize(X_h: np.ndarray, y_h: np.ndarray, m: int | None = None):
    """R (3,3), t (3,) minimising the algebraic error (pnp.py:132-152)."""
    X_h = np.asarray(X_h, dtype=np.float64)
    y_h = np.asarray(y_h, dtype=np.float64)
    if m is not None:
        X_h, y_h = X_h[:m], y_h[:m]
    if X_h.shape[0] < 6:
        raise ValueError("pnp_minimize needs m >= 6 correspondences")
    A = design_matrix(X_h, y_h)
    Vt = np.linalg.svd(A)[2]
    C0 = Vt[-1].reshape(3, 4)
    A3, b = C0[:, :3], C0[:, 3]
    tau = np.sign(np.linalg.det(A3))
    U, S, Vt3 = np.linalg.svd(tau * A3)
    R = U @ Vt3
    lam = 3.0 * tau / np.sum(S)
    return R, lam * b


def reprojection_error_sq(R: np.ndarray, t: np.ndarray, X: np.ndarray, y_h: np.ndarray) -> np.ndarray:
    """e_k = || pnorm(y_k) - pnorm(R x_k + t) ||^2   (ransac.py:21-35, 96-101; pnorm divides by the last entry)."""
    yp = X @ R.T + t
    with np.errstate(divide='ignore', invalid='ignore'):
        a = y_h[:, :2] / y_h[:, 2:3] - yp[:, :2] / yp[:, 2:3]
    return np.sum(a * a, axis=1)


def consensus(R, t, X, y_h, thresh: float) -> np.ndarray:
    """boolean membership, inclusive comparison ``thresh >= e`` (ransac.py:104-105); NaN -> not a member."""
    with np.errstate(invalid='ignore'):
        return thresh >= reprojection_error_sq(R, t, X, y_h)


def solve_hypotheses(X: np.ndarray, y_h: np.ndarray, idx: np.ndarray):
    """Poses (H, 3, 4) = (R | t) for every injected sample (ransac.py:77-82 with pnp_minimize as the solver)."""
    X_h = np.hstack([X, np.ones((X.shape[0], 1))])
    out = np.empty((idx.shape[0], 3, 4))
    for h, sel in enumerate(idx):
        try:
            R, t = pnp_minimize(X_h[sel], y_h[sel])
            out[h, :, :3] = R
            out[h, :, 3] = t
        except np.linalg.LinAlgError:
            out[h] = np.nan
    return out


def sample_gap(X: np.ndarray, y_h: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """(sigma_11 - sigma_12) / sigma_1 of each sample's design matrix: how well-determined the minimiser is."""
    X_h = np.hstack([X, np.ones((X.shape[0], 1))])
    out = np.empty(idx.shape[0])
    for h, sel in enumerate(idx):
        s = np.linalg.svd(design_matrix(X_h[sel], y_h[sel]), compute_uv=False)
        out[h] = (s[10] - s[11]) / s[0] if s[0] > 0 else 0.0
    return out


def score_hypotheses(poses: np.ndarray, X: np.ndarray, y_h: np.ndarray, thresh: float) -> np.ndarray:
    counts = np.empty(poses.shape[0], dtype=np.int32)
    for h, P in enumerate(poses):
        counts[h] = int(np.count_nonzero(consensus(P[:, :3], P[:, 3], X, y_h, thresh)))
    return counts


def pnp_ransac(X: np.ndarray, y_h: np.ndarray, idx: np.ndarray, thresh: float, n_sel: int | None = None) -> dict:
    """ransac.py:72-111 with injected samples.  Selection = largest consensus over the first ``n_sel`` correspondences
    (the reference compares ``len(C[0])``, i.e. the D_med part, ransac.py:108), strict > so the first maximum wins."""
    X = np.asarray(X, dtype=np.float64)
    y_h = np.asarray(y_h, dtype=np.float64)
    n_sel = X.shape[0] if n_sel is None else n_sel
    poses = solve_hypotheses(X, y_h, idx)
    counts = score_hypotheses(poses, X[:n_sel], y_h[:n_sel], thresh)
    best = int(np.argmax(counts)) if counts.size and counts.max() > 0 else -1
    out = {"poses": poses, "counts": counts, "best": best, "R": None, "t": None,
           "mask": np.zeros(X.shape[0], dtype=np.uint8)}
    if best >= 0:
        out["R"], out["t"] = poses[best][:, :3], poses[best][:, 3]
        out["mask"] = consensus(out["R"], out["t"], X, y_h, thresh).astype(np.uint8)
    return out


def rotation_angle(Ra: np.ndarray, Rb: np.ndarray) -> float:
    """geodesic distance between two rotations (radians); chordal form, accurate for tiny angles."""
    chord = np.linalg.norm(np.asarray(Ra) - np.asarray(Rb)) / (2.0 * np.sqrt(2.0))
    return float(2.0 * np.arcsin(min(1.0, chord)))
