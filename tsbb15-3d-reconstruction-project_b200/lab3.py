"""Mirror of the reference's ``lab3`` module for the functions on (and right next to) the hot path.

On the hot path — arithmetic runs in CUDA through the C ABI, no CPU fallback:
    fmatrix_stls(pl, pr)        reference lab3.py:269-329
    fmatrix_residuals(F, x, y)  reference lab3.py:188-227

Next to the hot path — CUDA as well, one thread per correspondence, FP64:
    triangulate_optimal(C1, C2, x1, x2)   reference lab3.py:382-475   (+ triangulate_optimal_batch: all points, one call)
    triangulate_linear(C1, C2, x1, x2)    reference lab3.py:477-503   (+ triangulate_linear_batch)
    fmatrix_from_cameras(C1, C2)          reference lab3.py:331-351

Host-side helpers kept in numpy because they run once per image pair inside SciPy's Levenberg-Marquardt callback of
the gold-standard refinement when it runs on the host as in the reference (fun.py:342-369, `refine=True`; the device
version of that stage is fun.gold_standard_device): homog, project, cross_matrix,
fmatrix_cameras, fmatrix_epipoles, fmatrix_residuals_gs.  Same names, argument layouts, return shapes and errors.
"""
from __future__ import annotations

import numpy as np

from . import runtime as _rt


# ---------------------------------------------------------------------------------------------------------------
# hot path (GPU)
# ---------------------------------------------------------------------------------------------------------------
def fmatrix_stls(pl, pr):
    """Estimate the fundamental matrix with the normalised 8-point algorithm (N >= 8), pl^T F pr = 0.

    pl, pr : (2, N) left / right image coordinates.  Raises ValueError on a shape mismatch like the reference
    (lab3.py:283-284)."""
    pl = np.asarray(pl, dtype=np.float64)
    pr = np.asarray(pr, dtype=np.float64)
    if not pl.shape == pr.shape:
        raise ValueError('pl and pr must have same shape')
    if pl.ndim != 2 or pl.shape[0] != 2:
        raise ValueError('pl and pr must be (2, N) arrays')
    return _rt.fmatrix_stls(pl, pr)


def fmatrix_residuals(F, x, y):
    """Signed distances between the points and their epipolar lines, (2, N): row 0 for x, row 1 for y.

    F : (3, 3) with x^T F y = 0;  x, y : (2, N).  Raises ValueError if the sizes differ (lab3.py:207-208)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if not x.shape == y.shape:
        raise ValueError('x and y must have same sizes')
    F = np.asarray(F, dtype=np.float64)
    if F.shape != (3, 3):
        raise ValueError('F must be a (3, 3) matrix')
    return _rt.fmatrix_residuals(F, x, y)


# ---------------------------------------------------------------------------------------------------------------
# host helpers (numpy), semantics of the reference
# ---------------------------------------------------------------------------------------------------------------
def homog(x):
    """Append a row of ones: (d, n) -> (d+1, n); 1-D input is treated as one point (lab3.py:30-50)."""
    x = np.asarray(x, dtype=np.float64)
    flat = x.ndim != 2
    pts = x.reshape(-1, 1) if flat else x
    out = np.ones((pts.shape[0] + 1, pts.shape[1]))
    out[:-1] = pts
    return out.ravel() if flat else out


def project(x, C):
    """Pinhole projection of world point(s) x (3,) / (3, n) through camera C (3, 4) (lab3.py:52-72)."""
    C = np.asarray(C)
    if not C.shape == (3, 4):
        raise ValueError('C is not a valid camera matrix')
    y = C @ homog(x)
    return y[:2] / y[2]


def cross_matrix(v):
    """[v]_x such that [v]_x b == v x b (lab3.py:110-129)."""
    v = np.asarray(v).ravel()
    if not v.size == 3:
        raise ValueError('Can only handle 3D vectors')
    M = np.zeros((3, 3), dtype=np.result_type(v.dtype, np.float64))
    M[0, 1], M[0, 2] = -v[2], v[1]
    M[1, 0], M[1, 2] = v[2], -v[0]
    M[2, 0], M[2, 1] = -v[1], v[0]
    return M


def fmatrix_from_cameras(C1, C2):
    """F of a camera pair: e = C1 n with n the centre of C2, F = [e]_x C1 C2^+ (lab3.py:331-351).  GPU; the sign of F
    is arbitrary exactly as in the reference (it comes from the sign of an SVD null vector)."""
    C1 = np.asarray(C1, dtype=np.float64)
    C2 = np.asarray(C2, dtype=np.float64)
    if C1.shape != (3, 4) or C2.shape != (3, 4):
        raise ValueError('C1 and C2 must be (3, 4) camera matrices')
    return _rt.fmatrix_from_cameras(C1, C2)[0]


def fmatrix_cameras(F):
    """One camera pair consistent with F, second camera fixed to [I | 0] (lab3.py:353-380)."""
    e1 = np.linalg.svd(F)[0][:, -1]
    C1 = np.hstack([cross_matrix(e1) @ F, e1.reshape(3, 1)])
    C2 = np.hstack([np.eye(3), np.zeros((3, 1))])
    return C1, C2


def fmatrix_epipoles(F):
    """Inhomogeneous epipoles (e1^T F = 0, F e2 = 0), each a (2,) array (lab3.py:505-527)."""
    U, _, Vt = np.linalg.svd(F)
    e1 = U[:, -1] / U[-1, -1]
    e2 = Vt[-1] / Vt[-1, -1]
    return e1[:2], e2[:2]


def _points2(x, name):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1 or (x.ndim == 2 and x.shape[1] == 1 and x.shape[0] in (2, 3)):
        x = x.reshape(1, -1)
    if x.ndim != 2 or x.shape[1] not in (2, 3):
        raise ValueError(f'{name} must hold 2-D image points ((2,), (3,) homogeneous, or (N, 2))')
    if x.shape[1] == 3:                      # tables.py:170 passes homogeneous points; like the reference only the
        x = x[:, :2]                         # first two components are read (lab3.py:401-403), no division
    return np.ascontiguousarray(x)


def triangulate_linear_batch(C1, C2, x1, x2):
    """Linear triangulation of N correspondences in one GPU call: x1, x2 (N, 2) -> (N, 3)."""
    return _rt.triangulate(C1, C2, [_points2(x1, 'x1')], [_points2(x2, 'x2')], method=_rt.TRI_LINEAR)[0]


def triangulate_optimal_batch(C1, C2, x1, x2):
    """Optimal triangulation of N correspondences in one GPU call: x1, x2 (N, 2) -> (N, 3).  Replaces the Python loops
    around lab3.triangulate_optimal at fun.py:352 and tables.py:170, 243."""
    return _rt.triangulate(C1, C2, [_points2(x1, 'x1')], [_points2(x2, 'x2')], method=_rt.TRI_OPTIMAL)[0]


def triangulate_linear(C1, C2, x1, x2):
    """Homogeneous linear triangulation of one correspondence (lab3.py:477-503): (2,) + (2,) -> (3,)."""
    return triangulate_linear_batch(C1, C2, x1, x2)[0]


def triangulate_optimal(C1, C2, x1, x2):
    """Hartley-Sturm triangulation of one correspondence as the reference performs it (lab3.py:382-475): both image
    points are moved to the origin, both epipoles rotated onto the x axis, the stationary points of the summed squared
    distances to a pencil of corresponding epipolar lines are the roots of a degree-6 polynomial, the best of them (or
    the asymptote) gives the corrected points, which are triangulated linearly.  Like the reference the epipole scale
    factors are f = f' = 1 (lab3.py:421) and the real parts of ALL roots are tried.  Returns (3,)."""
    return triangulate_optimal_batch(C1, C2, x1, x2)[0]


def fmatrix_residuals_gs(params, pl, pr):
    """Gold-standard residual vector for scipy.optimize.least_squares (lab3.py:230-266).

    params = [C1.ravel() (12), X.T.ravel() (3N)];  second camera is [I | 0].  Returns (4N,) ordered
    leftx, lefty, rightx, righty.  Raises ValueError if the parameter vector does not match N."""
    params = np.asarray(params, dtype=np.float64)
    C1 = params[:12].reshape(3, 4)
    C2 = np.hstack([np.eye(3), np.zeros((3, 1))])
    X = params[12:].reshape(-1, 3).T
    if not X.shape[1] == pl.shape[1]:
        raise ValueError('Wrong size of parameter vector')
    r1 = pl - project(X, C1)
    r2 = pr - project(X, C2)
    return np.concatenate((r1.ravel(), r2.ravel()))
