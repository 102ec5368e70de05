"""TEST INFRASTRUCTURE ONLY — CPU (numpy / SciPy, float64) restatement of the reference's bundle adjustment
``Tables.BundleAdjustment2`` (the multi-view half).

Nothing under ``oracle/`` is imported by the product package; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs use it, and only as the checker / the timed CPU baseline.

Reference lines restated (paths relative to the reference root):
  * residual vector ................. tables.py:266-296  (``EpsilonBA``: r[2i] = u_i - c1.x / c3.x, r[2i+1] = v_i - c2.x / c3.x,
                                                         c = the 3x4 matrix of the observation's view, x = its 3-D point, w = 1)
  * parameter vector ................ tables.py:302-315, fun.py:282-289  (all 3x4 camera matrices raveled, then all points)
  * sparsity pattern ................ tables.py:346-380  (12 + 3 columns per residual row; the first camera's 12 columns
                                                         cleared = the first view is held fixed)
  * the solver call ................. tables.py:317      (scipy least_squares, method='trf', x_scale='jac', ftol=1e-4,
                                                         jac_sparsity = the mask, finite-difference Jacobian)
Parity pinning: ``tests/golden/ba_golden.npz`` holds runs of the UNMODIFIED ``Tables.BundleAdjustment2`` (on tables built
from the Dino data with the reference's own record classes): the residual vector the reference evaluates at the start,
SciPy's solution / cost / nfev / status.  ``tests/test_oracle_golden.py`` checks ``residuals`` / ``sparsity_mask`` /
``bundle_adjust_scipy`` against it.  ``bundle_adjust_lm`` is NOT a restatement of the reference: it is a numpy version of
the Levenberg-Marquardt / Schur-complement iteration that the CUDA path runs, kept here to check that path step by step.
"""
from __future__ import annotations

import numpy as np


def pack(cams, pts) -> np.ndarray:
    """tables.py:315: x0 = [Rktk.ravel(), xj.ravel()]."""
    return np.hstack([np.asarray(cams, dtype=np.float64).ravel(), np.asarray(pts, dtype=np.float64).ravel()])


def unpack(x, n_C, n_P):
    """fun.reshapeToCamera3DPoints2 (fun.py:282-289)."""
    x = np.asarray(x, dtype=np.float64)
    return x[:n_C * 12].reshape(n_C, 3, 4), x[n_C * 12:].reshape(n_P, 3)


def residuals(x, u, v, cam_idx, pt_idx, n_C, n_P) -> np.ndarray:
    """tables.py:266-296, vectorised over the observation table: interleaved (u, v) residuals, length 2 * n_obs."""
    C, X = unpack(x, n_C, n_P)
    Xh = np.concatenate([X, np.ones((n_P, 1))], axis=1)
    y = np.einsum("oab,ob->oa", C[cam_idx], Xh[pt_idx])
    r = np.empty(2 * len(u))
    r[0::2] = u - y[:, 0] / y[:, 2]
    r[1::2] = v - y[:, 1] / y[:, 2]
    return r


def sparsity_mask(cam_idx, pt_idx, n_C, n_P):
    """tables.py:346-380 as a scipy.sparse lil_matrix of ints."""
    from scipy.sparse import lil_matrix
    n_obs = len(cam_idx)
    A = lil_matrix((2 * n_obs, n_C * 12 + n_P * 3), dtype="int")
    i = np.arange(n_obs)
    for s in range(12):
        A[2 * i, cam_idx * 12 + s] = 1
        A[2 * i + 1, cam_idx * 12 + s] = 1
    for s in range(3):
        A[2 * i, n_C * 12 + pt_idx * 3 + s] = 1
        A[2 * i + 1, n_C * 12 + pt_idx * 3 + s] = 1
    A[:, 0:12] = 0
    return A


def bundle_adjust_scipy(cams, pts, uv, cam_idx, pt_idx):
    """tables.py:298-333 with the reference's SciPy call.  Returns new cameras, new points and the OptimizeResult."""
    from scipy.optimize import least_squares
    n_C, n_P = len(cams), len(pts)
    sol = least_squares(residuals, pack(cams, pts), args=(uv[:, 0], uv[:, 1], cam_idx, pt_idx, n_C, n_P),
                        jac_sparsity=sparsity_mask(cam_idx, pt_idx, n_C, n_P), x_scale="jac", ftol=1e-4, method="trf")
    C, X = unpack(sol.x, n_C, n_P)
    return C, X, sol


def cost(cams, pts, uv, cam_idx, pt_idx) -> float:
    r = residuals(pack(cams, pts), uv[:, 0], uv[:, 1], cam_idx, pt_idx, len(cams), len(pts))
    return 0.5 * float(r @ r)


def _inv_sym3(V):
    """Adjugate inverse of a stack of symmetric 3x3 (the device uses the same formula)."""
    return np.linalg.inv(V)


def bundle_adjust_lm(cams, pts, uv, cam_idx, pt_idx, n_fixed=1, max_iter=50, ftol=1e-12, lam=1e-3, trace=None):
    """Numpy version of the device algorithm (csrc/ba_kernels.cuh): Levenberg-Marquardt with Marquardt scaling over all
    12 entries of every non-fixed camera and all points; the 3x3 point blocks are eliminated (Schur complement), the
    reduced camera system is solved by Cholesky; accept / reject with lambda /10, x10; converged when an accepted step
    lowers the cost by no more than ftol * cost, or brings it to the rounding level of the observations.
    Returns cameras, points, cost, iterations, status (2 converged, 3 lambda overflow, 4 max_iter)."""
    C = np.array(cams, dtype=np.float64).reshape(-1, 3, 4)
    X = np.array(pts, dtype=np.float64).reshape(-1, 3)
    uv = np.asarray(uv, dtype=np.float64)
    cam_idx = np.asarray(cam_idx, dtype=np.int64)
    pt_idx = np.asarray(pt_idx, dtype=np.int64)
    n_C, n_P, n_O = len(C), len(X), len(uv)
    n_free = n_C - n_fixed
    c = cost(C, X, uv, cam_idx, pt_idx)
    floor = 0.5 * float(np.sum(uv * uv)) * (64.0 * np.finfo(np.float64).eps) ** 2      # residuals = rounding noise of uv
    status, it = 4, 0
    if n_free <= 0 and n_P == 0:
        return C, X, c, 0, 2
    for it in range(1, max_iter + 1):
        Xh = np.concatenate([X, np.ones((n_P, 1))], axis=1)
        xo = Xh[pt_idx]                                              # (O, 4)
        y = np.einsum("oab,ob->oa", C[cam_idx], xo)
        iy = 1.0 / y[:, 2]
        pred = y[:, :2] * iy[:, None]
        r = uv - pred
        P2 = np.zeros((n_O, 2, 3))
        P2[:, 0, 0] = iy; P2[:, 1, 1] = iy
        P2[:, 0, 2] = -pred[:, 0] * iy; P2[:, 1, 2] = -pred[:, 1] * iy
        Q = np.einsum("oka,oac->okc", P2, C[cam_idx][:, :, :3])       # d pred / d X
        M = np.einsum("oka,okb->oab", P2, P2)
        T = np.einsum("oka,okc->oac", P2, Q)
        V = np.zeros((n_P, 3, 3)); gp = np.zeros((n_P, 3))
        np.add.at(V, pt_idx, np.einsum("oka,okb->oab", Q, Q))
        np.add.at(gp, pt_idx, np.einsum("oka,ok->oa", Q, r))
        Vd = V + lam * np.einsum("pab,ab->pab", V, np.eye(3))
        seen = np.zeros(n_P, dtype=bool); seen[pt_idx] = True
        Vd[~seen] = np.eye(3)
        Vi = _inv_sym3(Vd)
        Vi[~seen] = 0.0
        e = np.einsum("pab,pb->pa", Vi, gp)
        XX = np.einsum("oa,ob->oab", xo, xo)
        # camera blocks
        n = 12 * n_C
        S = np.zeros((n, n)); rhs = np.zeros(n); dU = np.zeros(n)
        U = np.einsum("oab,ocd->oacbd", M, XX).reshape(n_O, 12, 12)
        z = np.einsum("oka,ok->oa", P2, r) - np.einsum("oac,oc->oa", T, e[pt_idx])
        for o in range(n_O):
            k = cam_idx[o]
            S[12 * k:12 * k + 12, 12 * k:12 * k + 12] += U[o]
            dU[12 * k:12 * k + 12] += np.diag(U[o])
            rhs[12 * k:12 * k + 12] += np.kron(z[o], xo[o])
        order = np.argsort(pt_idx, kind="stable")
        starts = np.searchsorted(pt_idx[order], np.arange(n_P + 1))
        for j in range(n_P):
            oj = order[starts[j]:starts[j + 1]]
            for o in oj:
                for o2 in oj:
                    k, l = cam_idx[o], cam_idx[o2]
                    S[12 * k:12 * k + 12, 12 * l:12 * l + 12] -= np.kron(T[o] @ Vi[j] @ T[o2].T, XX[o])
        S[np.arange(n), np.arange(n)] += np.where(dU > 0.0, lam * dU, 1.0)      # no information: unit diagonal, zero step
        f0 = 12 * n_fixed
        dc = np.zeros(n)
        ok = True
        if n_free > 0:
            try:
                L = np.linalg.cholesky(S[f0:, f0:])
                dc[f0:] = np.linalg.solve(L.T, np.linalg.solve(L, rhs[f0:]))
            except np.linalg.LinAlgError:
                ok = False
        if ok:
            dC = dc.reshape(n_C, 3, 4)
            q = np.einsum("oab,ob->oa", dC[cam_idx], xo)             # dC_k Xh per observation
            back = np.zeros((n_P, 3))
            np.add.at(back, pt_idx, np.einsum("oac,oa->oc", T, q))
            dX = np.einsum("pab,pb->pa", Vi, gp - back)
            Cn, Xn = C + dC, X + dX
            cn = cost(Cn, Xn, uv, cam_idx, pt_idx)
        if trace is not None:
            trace.append(dict(it=it, lam=lam, cost=c, cost_trial=cn if ok else np.nan, ok=ok))
        if ok and cn < c:
            gain = c - cn
            conv = gain <= ftol * c or cn <= floor
            C, X, c = Cn, Xn, cn
            lam = max(lam * 0.1, 1e-15)
            if conv:
                status = 2
                break
        else:
            lam *= 10.0
            if lam > 1e12:
                status = 3
                break
    return C, X, c, it, status


def dino_scene(Ps, x2d, X3d, n_views, seed=3, sig_obs=0.5 / 3217.0, sig_pt=2e-3, sig_cam=1e-3, camera_resectioning=None):
    """A bundle-adjustment problem in the layout main.py builds (first view = [I | 0], C-normalised observations) from
    the exact synthetic Dino data: the first n_views cameras, every point seen by at least two of them, observation
    noise sig_obs, start values off the truth by sig_pt (points) / sig_cam (cameras 1..).
    Returns cams (n_C,3,4), pts (n_P,3), uv (n_O,2), cam_idx, pt_idx."""
    if camera_resectioning is None:
        from .geom_path import camera_resectioning
    rng = np.random.default_rng(seed)
    KRt = [camera_resectioning(Ps[i]) for i in range(n_views)]
    K, R0, t0 = KRt[0]
    Ki = np.linalg.inv(K)
    vis = x2d[:n_views, 0, :] != -1
    use = np.flatnonzero(vis.sum(0) >= 2)
    cams = np.zeros((n_views, 3, 4))
    for i in range(n_views):
        _, R, t = KRt[i]
        Rn = R @ R0.T
        tn = t - Rn @ t0
        if i > 0:
            Rn = Rn + sig_cam * rng.standard_normal((3, 3))
            tn = tn + sig_cam * 0.1 * rng.standard_normal(3)
        cams[i, :, :3] = Rn
        cams[i, :, 3] = tn
    pts, uv, ci, pi = [], [], [], []
    for jn, j in enumerate(use):
        pts.append(R0 @ X3d[j] + t0 + sig_pt * rng.standard_normal(3))
        for i in range(n_views):
            if vis[i, j]:
                y = Ki @ np.array([x2d[i, 0, j], x2d[i, 1, j], 1.0])
                y = y / y[2]
                uv.append(y[:2] + sig_obs * rng.standard_normal(2))
                ci.append(i)
                pi.append(jn)
    return cams, np.array(pts), np.array(uv), np.array(ci, dtype=np.int64), np.array(pi, dtype=np.int64)
