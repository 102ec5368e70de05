"""Mirror of the reference's ``lab3`` module for the functions on (and right next to) the hot path.

On the hot path — arithmetic runs in CUDA through the C ABI, no CPU fallback:
    fmatrix_stls(pl, pr)        reference lab3.py:269-329
    fmatrix_residuals(F, x, y)  reference lab3.py:188-227

Host-side helpers kept in numpy because the reference keeps them on the host too and they run once per image pair
inside the gold-standard refinement of ``fun.getFFromLabCode`` (fun.py:342-369; SURVEY.md section 8f rows N3/N4,
"next"): homog, project, cross_matrix, fmatrix_from_cameras, fmatrix_cameras, fmatrix_epipoles, triangulate_linear,
triangulate_optimal, fmatrix_residuals_gs.  Same names, argument layouts, return shapes and error behaviour.
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
    """F of a camera pair: e = C1 n with n the centre of C2, F = [e]_x C1 C2^+ (lab3.py:331-351)."""
    centre = np.linalg.svd(C2)[2][3]
    e = C1 @ centre
    return cross_matrix(e) @ (C1 @ np.linalg.pinv(C2))


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


def triangulate_linear(C1, C2, x1, x2):
    """Homogeneous linear triangulation (lab3.py:477-503)."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    if x1.shape[0] == 2:
        x1, x2 = homog(x1), homog(x2)
    M = np.vstack([cross_matrix(x1) @ C1, cross_matrix(x2) @ C2])
    X = np.linalg.svd(M)[2][-1]
    return X[:3] / X[-1]


def triangulate_optimal(C1, C2, x1, x2):
    """Hartley-Sturm triangulation as the reference performs it (lab3.py:382-475).

    Both image points are moved to the origin, both epipoles are rotated onto the x axis, the stationary points of the
    summed squared distances to a pencil of corresponding epipolar lines are the roots of a degree-6 polynomial, the
    best of them (or the asymptote) gives the corrected points, which are triangulated linearly.  Like the reference
    the epipole scale factors are taken as f = f' = 1 (lab3.py:421) and the real parts of ALL roots are tried."""
    x1 = np.asarray(x1, dtype=np.float64).ravel()
    x2 = np.asarray(x2, dtype=np.float64).ravel()

    def shift(p):
        T = np.eye(3)
        T[0, 2], T[1, 2] = p[0], p[1]
        return T

    def epipole_rotation(e):
        return np.array([[e[0], e[1], 0.0], [-e[1], e[0], 0.0], [0.0, 0.0, 1.0]])

    T1, T2 = shift(x1), shift(x2)
    F = T1.T @ fmatrix_from_cameras(C1, C2) @ T2
    e1, e2 = fmatrix_epipoles(F)
    R1 = epipole_rotation(e1 / np.linalg.norm(e1))
    R2 = epipole_rotation(e2 / np.linalg.norm(e2))
    F = R1 @ F @ R2.T
    a, b, c, d = F[1, 1], F[1, 2], F[2, 1], F[2, 2]
    f1 = f2 = 1.0

    # g(t) = t ((at+b)^2 + f1^2 (ct+d)^2)^2 - (ad-bc) (1 + f2^2 t^2)^2 (at+b)(ct+d)
    P = np.polynomial.polynomial
    atb, ctd = np.array([b, a]), np.array([d, c])                 # ascending coefficients
    quad = P.polyadd(P.polymul(atb, atb), f1 ** 2 * P.polymul(ctd, ctd))
    term1 = P.polymul([0.0, 1.0], P.polymul(quad, quad))
    one = np.array([1.0, 0.0, f2 ** 2])
    term2 = (a * d - b * c) * P.polymul(P.polymul(one, one), P.polymul(atb, ctd))
    g = P.polysub(term1, term2)
    g = np.concatenate([g, np.zeros(7 - g.size)])[:7]
    roots = np.real(np.roots(g[::-1]))                            # np.roots wants descending order

    def cost(t):
        return t ** 2 / (1 + f2 ** 2 * t ** 2) + (c * t + d) ** 2 / ((a * t + b) ** 2 + f1 ** 2 * (c * t + d) ** 2)

    values = [cost(t) for t in roots]
    values.append(1.0 / f2 ** 2 + c ** 2 / (a ** 2 + f1 ** 2 * c ** 2))       # t -> infinity
    k = int(np.argmin(values))
    if k < roots.size:
        t = roots[k]
        l1 = np.array([-f1 * (c * t + d), a * t + b, c * t + d])
        l2 = np.array([t * f2, 1.0, -t])
    else:
        l1 = np.array([-f1 * c, a, c])
        l2 = np.array([f2, 0.0, -1.0])

    def foot(l):                                                  # closest point of the line to the origin
        return np.array([-l[0] * l[2], -l[1] * l[2], l[0] ** 2 + l[1] ** 2]).reshape(3, 1)

    x1n = T1 @ (R1.T @ foot(l1))
    x2n = T2 @ (R2.T @ foot(l2))
    return triangulate_linear(C1, C2, x1n, x2n)


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
