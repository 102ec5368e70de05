"""Host-side sample index generation (reference: fun.py:305-308, ransac.py:12-19).

North-star contract: sample index sets are drawn on the HOST from one seed and fed to both the GPU path and the
oracle, so inlier counts and the selected hypothesis can be compared bit for bit.
"""
from __future__ import annotations

import random as _random

import numpy as np


def reference_stream(n_points: int, n_hyp: int, k: int = 8) -> np.ndarray:
    """Exactly the reference's draw: ``np.random.choice(np.arange(N), k, replace=False)`` once per trial from the
    GLOBAL numpy RandomState (fun.py:305-306).  ``np.random.seed(s)`` before this call reproduces the index sets the
    reference itself would use after the same seed.  O(N) per draw, like the reference."""
    index_points = np.arange(0, n_points, 1)
    out = np.empty((n_hyp, k), dtype=np.int32)
    for h in range(n_hyp):
        out[h] = np.random.choice(index_points, k, replace=False)
    return out


def fast(n_points: int, n_hyp: int, k: int = 8, seed: int | None = 0) -> np.ndarray:
    """(H, k) index sets without replacement within a row, from ``np.random.default_rng(seed)``; vectorised
    rejection of rows that contain a repeated index (cost independent of N)."""
    if n_points < k:
        raise ValueError("Cannot generate more indices than the amount of values in the set from which they are "
                         "extracted. n should therefore be smaller or equal to set_length")
    rng = np.random.default_rng(seed)
    out = rng.integers(0, n_points, size=(n_hyp, k), dtype=np.int64)
    while True:
        s = np.sort(out, axis=1)
        bad = np.flatnonzero((s[:, 1:] == s[:, :-1]).any(axis=1))
        if bad.size == 0:
            break
        out[bad] = rng.integers(0, n_points, size=(bad.size, k), dtype=np.int64)
    return out.astype(np.int32)


_PINNED: dict = {}


def fast_batch(n_points_list, n_hyp: int, k: int = 8, seed: int = 0, pinned: bool = False, device=None) -> list:
    """``fast(n_p, n_hyp, k, seed + p)`` for every pair / view p, drawn into ONE contiguous (sum H, k) int32 array and
    returned as the list of its row blocks: the batched entry points recognise consecutive blocks and hand the parent array
    to the library without concatenating (for the Dino sequence the 11 MB copy was two thirds of the host call).
    pinned=True: draw into a page-locked buffer that is CACHED per (thread, size) and reused by the next call of the same
    shape (cudaMallocHost / cudaFreeHost cost milliseconds and synchronise the device, so a fresh buffer per call would be a
    net loss; the returned blocks are only valid until the next pinned call of that shape from the same thread).
    Callers that do not need host-visible samples should let the library draw them on the device instead
    (runtime.f_ransac_batched(idx_list=None, n_hyp=..., sample_seed=...))."""
    n_points_list = [int(n) for n in n_points_list]
    shape = (len(n_points_list) * int(n_hyp), k)
    if pinned:
        import threading
        from . import _cabi
        key = (threading.get_ident(), shape, device)
        parent = _PINNED.get(key)
        if parent is None:
            parent = _cabi.pinned_empty(shape, np.int32, device=device)
            _PINNED[key] = parent
    else:
        parent = np.empty(shape, dtype=np.int32)
    out = []
    for p, n in enumerate(n_points_list):
        blk = parent[p * n_hyp:(p + 1) * n_hyp]
        blk[...] = fast(n, n_hyp, k, seed + p)
        out.append(blk)
    return out


def gen_rnd_indices(set_length: int, n: int) -> list:
    """ransac.gen_rnd_indices (ransac.py:12-19): shuffle range(set_length) with the global ``random`` module state and
    keep the first n."""
    if set_length < n:
        raise ValueError("Cannot generate more indices than the amount of values in the set from which they are "
                         "extracted. n should therefore be smaller or equal to set_length")
    order = list(range(set_length))
    _random.shuffle(order)
    return order[0:n]
