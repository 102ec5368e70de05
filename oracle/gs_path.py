"""TEST INFRASTRUCTURE ONLY — CPU (numpy / SciPy, float64) restatement of the gold-standard stage of
``fun.getFFromLabCode``.

Nothing under ``oracle/`` is imported by the product package; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs use it, and only as the checker / the timed CPU baseline.

Reference lines restated (paths relative to the reference root):
  * residual vector ................. lab3.py:230-266  (``fmatrix_residuals_gs``)
  * the stage itself ................ fun.py:342-369   (cameras from F, optimal triangulation, scipy least_squares with
                                                       xtol=2.22e-14, tr_solver='lsmr', F from the refined cameras)
Parity pinning: ``tests/golden/gs_golden.npz`` holds the reference's own run on the noisy Dino pair (0,1) (residuals at
the start, SciPy's cost / nfev / status / solution, F_gold); ``tests/test_oracle_golden.py`` checks the restatement
against it.  ``gold_standard_lm`` is NOT a restatement of the reference: it is a dense numpy version of the
Levenberg-Marquardt / Schur-complement iteration that the CUDA path runs, kept here to check that path step by step.
"""
from __future__ import annotations

import numpy as np

from . import geom_path as g


def fmatrix_residuals_gs(params, pl, pr) -> np.ndarray:
    """lab3.py:230-266: params = [C1.ravel(), X.T.ravel()], C2 = [I | 0]; returns leftx, lefty, rightx, righty (4N)."""
    params = np.asarray(params, dtype=np.float64)
    C1 = params[:12].reshape(3, 4)
    X = params[12:].reshape(-1, 3).T
    if X.shape[1] != pl.shape[1]:
        raise ValueError('Wrong size of parameter vector')
    y = C1 @ np.vstack([X, np.ones(X.shape[1])])
    return np.concatenate(((pl - y[:2] / y[2]).ravel(), (pr - X[:2] / X[2]).ravel()))


def start_point(F, in1, in2):
    """fun.py:345-355: cameras from F and the optimally triangulated inliers."""
    C1, C2 = g.fmatrix_cameras(F)
    X = g.triangulate_optimal_batch(C1, C2, in1.T, in2.T).T
    return C1, C2, X


def gold_standard_scipy(F, in1, in2):
    """fun.py:342-369 with the reference's SciPy call.  Returns F_gold and the OptimizeResult."""
    from scipy.optimize import least_squares
    C1, C2, X = start_point(F, in1, in2)
    sol = least_squares(fmatrix_residuals_gs, np.hstack((C1.ravel(), X.T.ravel())), xtol=2.22e-14, tr_solver='lsmr',
                        args=(in1, in2))
    return g.fmatrix_from_cameras(sol.x[:12].reshape(3, 4), C2), sol


def cost(C1, X, in1, in2) -> float:
    r = fmatrix_residuals_gs(np.hstack((C1.ravel(), X.T.ravel())), in1, in2)
    return 0.5 * float(r @ r)


def gold_standard_lm(F, in1, in2, max_iter=50, ftol=1e-12, lam=1e-3):
    """Dense numpy version of the device algorithm: Levenberg-Marquardt with Marquardt scaling, point blocks eliminated
    by the Schur complement, accept / reject with lambda /10, x10.  Returns F_gold, cost, iterations, C1, X."""
    C1, C2, X = start_point(F, in1, in2)
    N = in1.shape[1]
    c = cost(C1, X, in1, in2)
    it = 0
    for it in range(1, max_iter + 1):
        Xh = np.vstack([X, np.ones(N)])
        y = C1 @ Xh
        rl = in1 - y[:2] / y[2]
        rr = in2 - X[:2] / X[2]
        S = np.zeros((12, 12)); rhs = np.zeros(12); dA = np.zeros(12)
        blocks = []
        for i in range(N):
            P2 = np.array([[1, 0, -y[0, i] / y[2, i]], [0, 1, -y[1, i] / y[2, i]]]) / y[2, i]
            Q = P2 @ C1[:, :3]
            Pr = np.array([[1, 0, -X[0, i] / X[2, i]], [0, 1, -X[1, i] / X[2, i]]]) / X[2, i]
            M, T = P2.T @ P2, P2.T @ Q
            D = Q.T @ Q + Pr.T @ Pr
            gp = -(Q.T @ rl[:, i]) - Pr.T @ rr[:, i]
            Di = np.linalg.inv(D + lam * np.diag(np.diag(D)))
            XX = np.outer(Xh[:, i], Xh[:, i])
            S += np.kron(M - T @ Di @ T.T, XX)
            dA += np.kron(np.diag(M), np.diag(XX))
            rhs += np.kron(P2.T @ rl[:, i] + T @ Di @ gp, Xh[:, i])
            blocks.append((T, Di, gp))
        try:
            dc = np.linalg.solve(S + lam * np.diag(dA), rhs).reshape(3, 4)
        except np.linalg.LinAlgError:
            lam *= 10.0
            continue
        Xn = X.copy()
        for i in range(N):
            T, Di, gp = blocks[i]
            Xn[:, i] = X[:, i] + Di @ (-gp - T.T @ (dc @ Xh[:, i]))
        cn = cost(C1 + dc, Xn, in1, in2)
        if cn < c:
            gain = c - cn
            conv = gain <= ftol * c
            C1, X, c = C1 + dc, Xn, cn
            lam = max(lam * 0.1, 1e-15)
            if conv:
                break
        else:
            lam *= 10.0
            if lam > 1e12:
                break
    return g.fmatrix_from_cameras(C1, C2), c, it, C1, X
