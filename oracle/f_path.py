"""TEST INFRASTRUCTURE ONLY — CPU (numpy, float64) restatement of the reference's F-matrix RANSAC path.

Nothing under ``oracle/`` is imported by the product package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it,
and only as the checker / the timed CPU baseline.

Parity pinning: every function here is checked against the *unmodified* reference code
(``tests/test_oracle_vs_reference.py``, runs when /root/reference is mounted) and against golden
vectors generated from the reference by ``oracle/gen_golden.py`` (``tests/golden/*.npz``, travel to the
GPU box), including the reference's own shipped artefact ``Fmatrix.npy``.

Reference lines restated (paths relative to the reference root):
  * 8-point solver ............ lab3.py:269-329  (``fmatrix_stls``)
  * epipolar residuals ........ lab3.py:188-227  (``fmatrix_residuals``), ``homog`` lab3.py:30-50
  * RANSAC loop body .......... fun.py:303-328   (sample -> solve -> score -> threshold -> best update)
The arithmetic bottoms out in ``np.linalg.svd`` (LAPACK gesdd through numpy; numpy is unpinned by the
reference — no requirements file), exactly as in the reference.
"""
from __future__ import annotations

import numpy as np

EPI_MAX = 0   # reference criterion: max(|d1|, |d2|) < thr      (fun.py:316-317)
SAMPSON = 1   # extension named by BASELINE.json north_star: sqrt(r^2 / (|l1|^2 + |l2|^2)) < thr


def _hartley(x: np.ndarray) -> np.ndarray:
    """Isotropic scaling homography of lab3.py:288-295: zero mean, mean squared radius 2."""
    n = x.shape[1]
    mean = x.mean(axis=1)
    centred = x - mean[:, None]
    scale = np.sqrt(np.sum(centred ** 2) / (2.0 * n))
    inv = 1.0 / scale
    return np.array([[inv, 0.0, -mean[0] / scale],
                     [0.0, inv, -mean[1] / scale],
                     [0.0, 0.0, 1.0]])


def fmatrix_stls(pl: np.ndarray, pr: np.ndarray) -> np.ndarray:
    """8-point (N >= 8) DLT estimate of F with pl^T F pr = 0   (lab3.py:269-329)."""
    pl = np.asarray(pl, dtype=np.float64)
    pr = np.asarray(pr, dtype=np.float64)
    if pl.shape != pr.shape:
        raise ValueError('pl and pr must have same shape')
    n = pl.shape[1]
    S = _hartley(pl)
    T = _hartley(pr)
    # lab3.py:301-309 — only the affine part of the homography is applied
    X = pl[0] * S[0, 0] + pl[1] * S[0, 1] + S[0, 2]
    Y = pl[0] * S[1, 0] + pl[1] * S[1, 1] + S[1, 2]
    x = pr[0] * T[0, 0] + pr[1] * T[0, 1] + T[0, 2]
    y = pr[0] * T[1, 0] + pr[1] * T[1, 1] + T[1, 2]
    A = np.stack([X * x, X * y, X, Y * x, Y * y, Y, x, y, np.ones(n)], axis=1)   # (N, 9), lab3.py:312-315
    Vt = np.linalg.svd(A)[2]                     # full_matrices=True -> (9, 9): row 8 is the null direction
    Fs = Vt[-1].reshape(3, 3)
    U, s, Vt3 = np.linalg.svd(Fs)                # rank-2 projection, lab3.py:321-324
    s = s.copy()
    s[2] = 0.0
    Fs = U @ (np.diag(s) @ Vt3)
    return S.T @ (Fs @ T)                        # lab3.py:327


def fmatrix_residuals(F: np.ndarray, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Signed point-to-epipolar-line distances in both images, (2, N)   (lab3.py:188-227)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.shape != y.shape:
        raise ValueError('x and y must have same sizes')
    n = x.shape[1]
    xh = np.vstack([x, np.ones((1, n))])
    yh = np.vstack([y, np.ones((1, n))])
    l1 = F @ yh
    l2 = F.T @ xh
    n1 = np.sqrt(l1[0] ** 2 + l1[1] ** 2)
    n2 = np.sqrt(l2[0] ** 2 + l2[1] ** 2)
    with np.errstate(divide='ignore', invalid='ignore'):
        r1 = np.sum(l1 * xh, axis=0) / n1
        r2 = np.sum(l2 * yh, axis=0) / n2
    return np.vstack([r1, r2])


def distance(F: np.ndarray, x: np.ndarray, y: np.ndarray, mode: int = EPI_MAX) -> np.ndarray:
    """Per-correspondence scalar that is compared with the threshold."""
    if mode == EPI_MAX:
        with np.errstate(invalid='ignore'):
            return np.max(np.abs(fmatrix_residuals(F, x, y)), axis=0)      # fun.py:316
    # Sampson distance (first-order geometric error); not in the reference -> parity unpinned for this
    # mode, checked only against this restatement.
    n = x.shape[1]
    xh = np.vstack([x, np.ones((1, n))])
    yh = np.vstack([y, np.ones((1, n))])
    l1 = F @ yh
    l2 = F.T @ xh
    r = np.sum(l1 * xh, axis=0)
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.sqrt(r * r / (l1[0] ** 2 + l1[1] ** 2 + l2[0] ** 2 + l2[1] ** 2))


def inliers(F, x, y, thr: float, mode: int = EPI_MAX) -> np.ndarray:
    """Index set S = flatnonzero(d < thr): strict <, NaN/inf => outlier   (fun.py:317)."""
    with np.errstate(invalid='ignore'):
        return np.flatnonzero(distance(F, x, y, mode) < thr)


def solve_hypotheses(p1: np.ndarray, p2: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """F for every injected 8-index sample: fun.py:309-311 with the index draw factored out."""
    idx = np.asarray(idx)
    out = np.empty((idx.shape[0], 3, 3))
    for h, sel in enumerate(idx):
        with np.errstate(divide='ignore', invalid='ignore'):
            try:
                out[h] = fmatrix_stls(p1[:, sel], p2[:, sel])
            except np.linalg.LinAlgError:      # LAPACK refuses NaN input (coincident sample points)
                out[h] = np.nan
    return out


def sample_condition(p1: np.ndarray, p2: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """sigma_8 / sigma_1 of each sample's 8x9 design matrix (how well-determined the null vector is).

    Used by parity tests to decide which hypotheses may legitimately differ between LAPACK and the GPU
    (SURVEY.md section 7 'hard parts': rank-deficient samples make V[-1] arbitrary)."""
    idx = np.asarray(idx)
    out = np.empty(idx.shape[0])
    for h, sel in enumerate(idx):
        pl, pr = p1[:, sel], p2[:, sel]
        with np.errstate(divide='ignore', invalid='ignore'):
            S, T = _hartley(pl), _hartley(pr)
            X = pl[0] * S[0, 0] + S[0, 2]; Y = pl[1] * S[1, 1] + S[1, 2]
            x = pr[0] * T[0, 0] + T[0, 2]; y = pr[1] * T[1, 1] + T[1, 2]
            A = np.stack([X * x, X * y, X, Y * x, Y * y, Y, x, y, np.ones(8)], axis=1)
            if not np.all(np.isfinite(A)):
                out[h] = 0.0
                continue
            s = np.linalg.svd(A, compute_uv=False)
        out[h] = s[7] / s[0] if s[0] > 0 else 0.0
    return out


def score_hypotheses(F_all: np.ndarray, p1: np.ndarray, p2: np.ndarray, thr: float,
                     mode: int = EPI_MAX) -> np.ndarray:
    """Inlier count of every hypothesis over all N correspondences (fun.py:315-317)."""
    counts = np.empty(F_all.shape[0], dtype=np.int32)
    for h, F in enumerate(F_all):
        counts[h] = inliers(F, p1, p2, thr, mode).size
    return counts


def select_first_max(counts: np.ndarray) -> int:
    """Strict-> best update without the tie rule: lowest index among the maxima; -1 if all zero."""
    if counts.size == 0 or counts.max() <= 0:
        return -1
    return int(np.argmax(counts))


def select_reference_rule(counts, F_all, p1, p2, mode: int = EPI_MAX) -> int:
    """The reference's best-update including its tie rule (fun.py:320-328), verbatim semantics:

        if |S| > |S_best|:                      take it, remember std(d)
        elif |S| == |S_best| and norm(std_best) > norm(d):   take it, remember std(d)

    with S_best = [] and d_best = [] initially (``np.linalg.norm([]) == 0``).  ``d`` is the full
    length-N vector of max-abs distances of the *candidate*; ``std_best`` is the scalar population
    standard deviation of the incumbent's ``d``.
    """
    best, best_count, best_std = -1, 0, None
    for h in range(len(counts)):
        c = int(counts[h])
        if c > best_count:
            best, best_count = h, c
            best_std = float(np.std(distance(F_all[h], p1, p2, mode)))
        elif c == best_count:
            incumbent = 0.0 if best_std is None else abs(best_std)
            d = distance(F_all[h], p1, p2, mode)
            if incumbent > np.linalg.norm(d):
                best, best_count = h, c
                best_std = float(np.std(d))
    return best


def f_ransac(p1: np.ndarray, p2: np.ndarray, idx: np.ndarray, thr: float = 1.5,
             mode: int = EPI_MAX, tie: str = "reference") -> dict:
    """The loop of fun.py:303-328 with the sample indices injected (``idx`` is (H, 8) int).

    Returns every intermediate the GPU path is compared on."""
    p1 = np.asarray(p1, dtype=np.float64)
    p2 = np.asarray(p2, dtype=np.float64)
    F_all = solve_hypotheses(p1, p2, idx)
    counts = score_hypotheses(F_all, p1, p2, thr, mode)
    if tie == "reference":
        best = select_reference_rule(counts, F_all, p1, p2, mode)
    else:
        best = select_first_max(counts)
    mask = np.zeros(p1.shape[1], dtype=np.uint8)
    F_best = None
    if best >= 0:
        F_best = F_all[best]
        mask[inliers(F_best, p1, p2, thr, mode)] = 1
    return {"F_all": F_all, "counts": counts, "best": best, "F": F_best, "mask": mask}


def normalise_F(F: np.ndarray, ref: np.ndarray | None = None) -> np.ndarray:
    """Unit Frobenius norm and a fixed sign (F is only defined up to scale/sign)."""
    F = np.asarray(F, dtype=np.float64)
    nrm = np.linalg.norm(F)
    Fn = F / nrm if nrm > 0 else F
    if ref is not None:
        if np.sum(Fn * ref) < 0:
            Fn = -Fn
    else:
        k = np.argmax(np.abs(Fn))
        if Fn.flat[k] < 0:
            Fn = -Fn
    return Fn
