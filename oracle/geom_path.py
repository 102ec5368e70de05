"""TEST INFRASTRUCTURE ONLY — CPU (numpy, float64) restatement of the reference's two-view geometry that sits either
side of the RANSAC hot path.

Nothing under ``oracle/`` is imported by the product package; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs use it, and only as the checker / the timed CPU baseline.

Parity pinning: every function here is compared with the *unmodified* reference function it restates in
``tests/test_oracle_vs_reference.py`` (runs where /root/reference is mounted) and with golden vectors generated from
the reference by ``oracle/gen_golden.py`` (``tests/golden/geom_golden.npz``), including the reference's own artefact
``clean_data_eval.npy`` (relative pose of the clean Dino pair (0, 1)).

Reference lines restated (paths relative to the reference root):
  * F from a camera pair ............. lab3.py:331-351  (``fmatrix_from_cameras``)
  * camera pair from F ............... lab3.py:353-380  (``fmatrix_cameras``)
  * epipoles ......................... lab3.py:505-527  (``fmatrix_epipoles``)
  * optimal (Hartley-Sturm) triang. .. lab3.py:382-475  (``triangulate_optimal``)
  * linear triangulation ............. lab3.py:477-503  (``triangulate_linear``)
  * det-corrected SVD ................ fun.py:186-207   (``specSVD``)
  * relative pose + cheirality ....... fun.py:209-258   (``relative_camera_pose``)
  * RQ camera decomposition .......... fun.py:174-184, 260-283 (``specRQ``, ``camera_resectioning``)
  * E = K^T F K ...................... fun.py:91-102    (``getEAndK``)
  * 2D<->3D match loop ............... tables.py:116-135 (``Tables.addNewView``, first observation within 1e-4)
"""
from __future__ import annotations

import numpy as np


def cross_matrix(v) -> np.ndarray:
    """lab3.py:110-129."""
    v = np.asarray(v, dtype=np.float64).ravel()
    if v.size != 3:
        raise ValueError('Can only handle 3D vectors')
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def fmatrix_from_cameras(C1, C2) -> np.ndarray:
    """lab3.py:331-351: n = centre of C2 (last right singular vector), e = C1 n, F = [e]x C1 pinv(C2)."""
    C1 = np.asarray(C1, dtype=np.float64)
    C2 = np.asarray(C2, dtype=np.float64)
    n = np.linalg.svd(C2)[2][3]
    return cross_matrix(C1 @ n) @ (C1 @ np.linalg.pinv(C2))


def fmatrix_cameras(F):
    """lab3.py:353-380: C1 = ([e1]x F | e1) with e1 the left null vector, C2 = [I | 0]."""
    F = np.asarray(F, dtype=np.float64)
    e1 = np.linalg.svd(F)[0][:, -1]
    return np.hstack([cross_matrix(e1) @ F, e1[:, None]]), np.hstack([np.eye(3), np.zeros((3, 1))])


def fmatrix_epipoles(F):
    """lab3.py:505-527 (without the reference's in-place division of the SVD factors)."""
    U, _, Vt = np.linalg.svd(np.asarray(F, dtype=np.float64))
    return U[:2, -1] / U[2, -1], Vt[-1, :2] / Vt[-1, 2]


def triangulate_linear(C1, C2, x1, x2) -> np.ndarray:
    """lab3.py:477-503.  x1, x2: (2,) inhomogeneous or (3,) / (3,1) homogeneous (then used as they are, unscaled)."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    if x1.shape[0] == 2:
        x1 = np.append(x1.ravel(), 1.0)
        x2 = np.append(x2.ravel(), 1.0)
    M = np.vstack([cross_matrix(x1) @ C1, cross_matrix(x2) @ C2])
    X = np.linalg.svd(M)[2][-1]
    return X[:3] / X[3]


def sextic(a, b, c, d) -> np.ndarray:
    """Coefficients (descending, t^6 .. t^0) of the stationarity polynomial of lab3.py:421-437 for f = f' = 1:
    g(t) = t ((at+b)^2 + (ct+d)^2)^2 - (ad-bc) (1+t^2)^2 (at+b)(ct+d)."""
    k1 = b * c - a * d
    return np.array([
        a * c * k1,
        (a * a + c * c) ** 2 + k1 * (b * c + a * d),
        4.0 * (a * a + c * c) * (a * b + c * d) + 2.0 * a * c * k1 + b * d * k1,
        2.0 * (4.0 * a * b * c * d + 3.0 * a * a * b * b + c * c * (3.0 * d * d + 2.0 * b * b)),
        -a * a * c * d + a * b * (4.0 * b * b + c * c + 2.0 * d * d) + 2.0 * c * d * (2.0 * d * d + 3.0 * b * b),
        b ** 4 - a * a * d * d + d ** 4 + b * b * (c * c + 2.0 * d * d),
        b * d * k1])


def triangulate_optimal(C1, C2, x1, x2) -> np.ndarray:
    """lab3.py:382-475.  Like the reference: epipole scale factors f = f' = 1 (lab3.py:421), real parts of ALL roots
    of the sextic are candidates (lab3.py:440), first minimum of the cost wins (np.argmin), the corrected points are
    kept homogeneous and unscaled when they enter the linear triangulation."""
    x1 = np.asarray(x1, dtype=np.float64).ravel()[:2]
    x2 = np.asarray(x2, dtype=np.float64).ravel()[:2]
    T1 = np.array([[1.0, 0.0, x1[0]], [0.0, 1.0, x1[1]], [0.0, 0.0, 1.0]])
    T2 = np.array([[1.0, 0.0, x2[0]], [0.0, 1.0, x2[1]], [0.0, 0.0, 1.0]])
    F = T1.T @ fmatrix_from_cameras(C1, C2) @ T2
    e1, e2 = fmatrix_epipoles(F)
    e1 = e1 / np.linalg.norm(e1)
    e2 = e2 / np.linalg.norm(e2)
    R1 = np.array([[e1[0], e1[1], 0.0], [-e1[1], e1[0], 0.0], [0.0, 0.0, 1.0]])
    R2 = np.array([[e2[0], e2[1], 0.0], [-e2[1], e2[0], 0.0], [0.0, 0.0, 1.0]])
    F = R1 @ F @ R2.T
    a, b, c, d = F[1, 1], F[1, 2], F[2, 1], F[2, 2]
    roots = np.real(np.roots(sextic(a, b, c, d)))
    with np.errstate(divide='ignore', invalid='ignore'):
        cost = [t * t / (1.0 + t * t) + (c * t + d) ** 2 / ((a * t + b) ** 2 + (c * t + d) ** 2) for t in roots]
        cost.append(1.0 + c * c / (a * a + c * c))
    k = int(np.argmin(cost))
    if k < roots.size:
        t = roots[k]
        l1 = np.array([-(c * t + d), a * t + b, c * t + d])
        l2 = np.array([t, 1.0, -t])
    else:
        l1 = np.array([-c, a, c])
        l2 = np.array([1.0, 0.0, -1.0])
    foot = lambda l: np.array([-l[0] * l[2], -l[1] * l[2], l[0] ** 2 + l[1] ** 2])
    return triangulate_linear(C1, C2, T1 @ (R1.T @ foot(l1)), T2 @ (R2.T @ foot(l2)))


def triangulate_optimal_batch(C1, C2, x1, x2) -> np.ndarray:
    """(N, 2), (N, 2) -> (N, 3): the per-correspondence loops of fun.py:352, tables.py:170, 243."""
    return np.stack([triangulate_optimal(C1, C2, a, b) for a, b in zip(np.asarray(x1), np.asarray(x2))]) \
        if len(x1) else np.zeros((0, 3))


def triangulate_linear_batch(C1, C2, x1, x2) -> np.ndarray:
    return np.stack([triangulate_linear(C1, C2, a, b) for a, b in zip(np.asarray(x1), np.asarray(x2))]) \
        if len(x1) else np.zeros((0, 3))


def spec_svd(M):
    """fun.py:186-207: SVD with the last columns of U and V flipped so that det U = det V = +1 (sigma_3 gets the sign).
    Returns U, S, V^T."""
    U, S, Vt = np.linalg.svd(np.asarray(M, dtype=np.float64))
    V = Vt.T.copy()
    U = U.copy()
    S = S.copy()
    du, dv = np.linalg.det(U), np.linalg.det(V)
    U[:, -1] *= du
    V[:, -1] *= dv
    S[-1] *= du * dv
    return U, S, V.T


def pose_candidates(E):
    """The four (R, t) of fun.py:213-235 in the reference's order: (V W U^T, v3), (V W^T U^T, v3), (V W U^T, -v3),
    (V W^T U^T, -v3)."""
    U, _, Vt = spec_svd(E)
    V = Vt.T
    W = np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    Ra, Rb = V @ W @ U.T, V @ W.T @ U.T
    v3 = V[:, -1]
    return [(Ra, v3), (Rb, v3), (Ra, -v3), (Rb, -v3)]


def relative_camera_pose(E, y1, y2):
    """fun.py:209-258: first candidate for which the optimally triangulated point of the one correspondence (y1, y2)
    (C-normalised) lies in front of both cameras; None if there is none (the reference falls off the end)."""
    C0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    for R, t in pose_candidates(E):
        x1 = triangulate_optimal(C0, np.hstack([R, t[:, None]]), y1, y2)
        x2 = R @ x1 + t
        if x1[-1] > 0 and x2[-1] > 0:
            return R, t
    return None


def spec_rq(M):
    """fun.py:174-184 (the reference's ``det(Q) == -1`` test is an exact float comparison; restated as is)."""
    import scipy.linalg
    U, Q = scipy.linalg.rq(np.asarray(M, dtype=np.float64))
    if np.linalg.det(Q) == -1:
        U[0, :] *= -1.0
        Q[:, 0] *= -1.0
    return U, Q


def camera_resectioning(C):
    """fun.py:260-283: C = K [R | t] with K upper triangular (positive diagonal, K[2,2] = 1), det R = +1."""
    C = np.asarray(C, dtype=np.float64)
    A, b = C[:, :3], C[:, 3]
    U, Q = spec_rq(A)
    t = np.linalg.solve(U, b)
    U = U / U[2, 2]
    D = np.diag(np.sign(np.diag(U)))
    K = U @ D
    if np.linalg.det(D) == 1:
        return K, D @ Q, D @ t
    return K, -1.0 * D @ Q, -1.0 * D @ t


def essential_from_F(K, F) -> np.ndarray:
    """fun.py:100-101: E = K^T F K."""
    K = np.asarray(K, dtype=np.float64)
    return K.T @ np.asarray(F, dtype=np.float64) @ K


def match_first_within(obs, y, tol=1e-4) -> np.ndarray:
    """tables.py:116-124: for every row y[i] the index of the FIRST row of ``obs`` with ||obs[v] - y[i]|| < tol
    (np.linalg.norm, strict <), -1 if there is none."""
    obs = np.asarray(obs, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    out = np.full(y.shape[0], -1, dtype=np.int32)
    for i in range(y.shape[0]):
        for v in range(obs.shape[0]):
            if np.linalg.norm(obs[v] - y[i]) < tol:
                out[i] = v
                break
    return out
