"""numpy twin of csrc/philox.cuh: the counter-based draws the library makes on the DEVICE, replayed on the host bit for bit.

* ``sample_indices``  — the (H, k) sample index sets of a seeded RANSAC call (``rg_f_ransac_dev2 / _host2`` with
  ``idx = NULL``; reference draw: ``np.random.choice(N, 8, replace=False)`` per trial, fun.py:305-308).  The north star wants
  ONE seed to feed the GPU path and the oracle with the same samples: with device-side drawing the 4-11 MB of indices per
  batch never cross PCIe, and the oracle gets them from this function.
* ``synth_two_view``  — the synthetic correspondences of BASELINE configs 3-5 (``rg_synth_two_view_dev``).

Only integer arithmetic and single IEEE double operations in a fixed order are used on both sides, so equality is exact
(tests/test_philox.py on the CPU, tests/test_gpu_philox.py against the device).
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
DOM_SAMPLE, DOM_POINT, DOM_CAMS = 0x53414D50, 0x504F494E, 0x43414D53


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox4x32-10 (Salmon et al., SC'11) on arrays of counters; returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & _MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def sample_indices(n_points: int, n_hyp: int, k: int = 8, seed: int = 0, pair_id: int = 0, hyp_first: int = 0) -> np.ndarray:
    """(n_hyp, k) int32: k distinct indices of [0, n_points) per hypothesis — exactly what the device draws for
    (seed, global pair id, hyp_first + h).  Draw j takes r = floor(u_j (n - j) / 2^32) and maps it to the r-th index not
    chosen before, so there is no rejection loop and the cost does not depend on n."""
    if n_points < k:
        raise ValueError("Cannot generate more indices than the amount of values in the set from which they are "
                         "extracted. n should therefore be smaller or equal to set_length")
    if not 1 <= k <= 8:
        raise ValueError("k must be in [1, 8]")
    h = np.arange(hyp_first, hyp_first + n_hyp, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    u = list(philox4x32_10(h, 0, pair_id, DOM_SAMPLE, k0, k1))
    if k > 4:
        u += list(philox4x32_10(h, 1, pair_id, DOM_SAMPLE, k0, k1))
    out = np.empty((n_hyp, k), dtype=np.int64)
    srt = np.empty((n_hyp, k), dtype=np.int64)
    for j in range(k):
        r = ((u[j].astype(np.uint64) * np.uint64(n_points - j)) >> np.uint64(32)).astype(np.int64)
        for t in range(j):
            r = r + (r >= srt[:, t])
        out[:, j] = r
        v = r.copy()
        for t in range(j):
            s = srt[:, t].copy()
            sw = v < s
            srt[:, t] = np.where(sw, v, s)
            v = np.where(sw, s, v)
        srt[:, j] = v
    return out.astype(np.int32)


def sample_indices_batch(n_points_list, n_hyp: int, k: int = 8, seed: int = 0, first_pair: int = 0) -> list:
    return [sample_indices(int(n), n_hyp, k, seed, first_pair + p) for p, n in enumerate(n_points_list)]


def _u01(u):
    return (u.astype(np.float64) + 0.5) * 2.3283064365386963e-10


def _lerp(lo, hi, t):
    return lo + (hi - lo) * t


def _ih12(words):
    s = np.zeros(words[0].shape, dtype=np.int64)
    for w in words:
        s += (w & np.uint32(0xFFFF)).astype(np.int64) + (w >> np.uint32(16)).astype(np.int64)
    return (s - 393210).astype(np.float64) * 1.52587890625e-05


def _project(C, X0, X1, X2):
    a = ((C[0, 0] * X0 + C[0, 1] * X1) + C[0, 2] * X2) + C[0, 3]
    b = ((C[1, 0] * X0 + C[1, 1] * X1) + C[1, 2] * X2) + C[1, 3]
    w = ((C[2, 0] * X0 + C[2, 1] * X1) + C[2, 2] * X2) + C[2, 3]
    return a / w, b / w


def camera_pair(pair_id: int, n_cams: int, seed_base: int = 1000):
    seed = seed_base + pair_id
    c = philox4x32_10(0, 0, pair_id, DOM_CAMS, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    c1 = (int(c[0]) * n_cams) >> 32
    c2 = (int(c[1]) * (n_cams - 1)) >> 32
    c2 += 1 if c2 >= c1 else 0
    return c1, c2


def synth_two_view(n: int, pair_id: int, cams, bbox, seed_base: int = 1000, outlier_frac: float = 0.3,
                   sigma_px: float = 0.5, image_size=(640.0, 480.0)):
    """(n, 4) rows (x0, x1, y0, y1) of global pair ``pair_id`` — the host replay of rg_synth_two_view_dev — and the
    camera pair drawn for it.  cams: (n_cams, 3, 4); bbox: (3, 2) lo/hi of the world box."""
    cams = np.asarray(cams, dtype=np.float64).reshape(-1, 3, 4)
    bbox = np.asarray(bbox, dtype=np.float64).reshape(3, 2)
    c1, c2 = camera_pair(pair_id, cams.shape[0], seed_base)
    seed = seed_base + pair_id
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    i = np.arange(n, dtype=np.uint64)
    w = []
    for b in range(8):
        w += list(philox4x32_10(i, b, pair_id, DOM_POINT, k0, k1))
    X0 = _lerp(bbox[0, 0], bbox[0, 1], _u01(w[0]))
    X1 = _lerp(bbox[1, 0], bbox[1, 1], _u01(w[1]))
    X2 = _lerp(bbox[2, 0], bbox[2, 1], _u01(w[2]))
    xu, xv = _project(cams[c1], X0, X1, X2)
    yu, yv = _project(cams[c2], X0, X1, X2)
    xu = xu + sigma_px * _ih12(w[8:14])
    xv = xv + sigma_px * _ih12(w[14:20])
    yu = yu + sigma_px * _ih12(w[20:26])
    yv = yv + sigma_px * _ih12(w[26:32])
    n_out = int(np.floor(outlier_frac * n + 0.5))
    yu[:n_out] = image_size[0] * _u01(w[3][:n_out])
    yv[:n_out] = image_size[1] * _u01(w[4][:n_out])
    return np.ascontiguousarray(np.stack([xu, xv, yu, yv], axis=1)), (c1, c2)
