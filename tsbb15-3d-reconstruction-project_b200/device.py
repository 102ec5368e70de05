"""Device-resident entry points (the ``_dev`` half of the C ABI) for callers that keep their batches in HBM.

torch is used ONLY as a carrier of device memory and streams (``data_ptr()`` / ``current_stream()``); every kernel that
runs is in librg_b200.so.  All calls are asynchronous on torch's current stream.  ``parallel.py`` builds the multi-GPU
decompositions on top of these, ``bench.py`` times them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi as cabi
from ._cabi import FLAG_REUSE_POINTS, MODE_EPI_MAX, SCORE_FP32_GUARDED, SOLVER_QR, TIE_FIRST

_vp = C.c_void_p
_pi32 = C.POINTER(C.c_int32)


def _torch():
    import torch
    return torch


def _stream():
    return _torch().cuda.current_stream().cuda_stream


def _dev_index(t) -> int:
    return int(t.device.index if t.device.index is not None else _torch().cuda.current_device())


def offsets(counts) -> np.ndarray:
    """int32 prefix table [0, c0, c0+c1, ...] (host) of per-pair counts."""
    counts = np.asarray(counts, dtype=np.int64).ravel()
    off = np.zeros(counts.size + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    if off[-1] >= 2 ** 31:
        raise ValueError("batch too large for int32 offsets")
    return off.astype(np.int32)


def synth_two_view(P: int, N: int, first_pair: int = 0, seed_base: int = 1000, cams=None, bbox=None, outlier_frac: float = 0.3,
                   sigma_px: float = 0.5, image_size=(640.0, 480.0), device=None, out=None):
    """(P, N, 4) float64 CUDA tensor: the synthetic pairs ``first_pair .. first_pair + P`` of BASELINE config 5, generated
    on the device (rg_synth_two_view_dev); ``philox.synth_two_view`` replays any of them on the host.  Also returns the
    (P, 2) int32 tensor of the camera pairs drawn."""
    torch = _torch()
    from . import synth
    lib = cabi.load_library()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    cams = np.ascontiguousarray(synth.dino()["Ps"] if cams is None else cams, dtype=np.float64).reshape(-1, 12)
    bbox = np.ascontiguousarray(synth.DINO_BBOX if bbox is None else bbox, dtype=np.float64).reshape(6)
    pts = out if out is not None else torch.empty((P, N, 4), dtype=torch.float64, device=dev)
    cam_pair = torch.empty((P, 2), dtype=torch.int32, device=dev)
    ctx = cabi.context(dev.index)
    done = 0
    while done < P:                                       # grid.y is limited to 65535 pairs per launch
        n = min(P - done, 32768)
        cabi.check(lib.rg_synth_two_view_dev(_vp(ctx), _vp(_stream()), n, first_pair + done, N, _vp(cams.ctypes.data),
                                             cams.shape[0], _vp(bbox.ctypes.data), int(seed_base), float(outlier_frac),
                                             float(sigma_px), float(image_size[0]), float(image_size[1]),
                                             _vp(pts[done:].data_ptr()), _vp(cam_pair[done:].data_ptr())))
        done += n
    return pts, cam_pair


def sample_indices(n_points_list, n_hyp, k: int = 8, seed: int = 0, first_pair: int = 0, hyp_first: int = 0, device=None):
    """(sum H, k) int32 CUDA tensor of the index sets a seeded call draws (rg_sample_indices_dev)."""
    torch = _torch()
    lib = cabi.load_library()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    n_pts = np.ascontiguousarray(n_points_list, dtype=np.int32).ravel()
    P = n_pts.size
    hyp_off = offsets(np.full(P, n_hyp) if np.isscalar(n_hyp) else n_hyp)
    out = torch.empty((int(hyp_off[-1]), k), dtype=torch.int32, device=dev)
    cabi.check(lib.rg_sample_indices_dev(_vp(cabi.context(dev.index)), _vp(_stream()), P, n_pts.ctypes.data_as(_pi32),
                                         hyp_off.ctypes.data_as(_pi32), int(k), int(seed), int(first_pair), int(hyp_first),
                                         _vp(out.data_ptr())))
    return out


class FOutputs:
    """Device output buffers of a batched F-RANSAC call over P pairs (allocated once, reused across calls)."""

    def __init__(self, P: int, n_total: int = 0, device=None, want_mask: bool = False, want_key: bool = False):
        torch = _torch()
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        # one contiguous block per pair: [best_idx i32][best_count i32][F 9 x f64] is what the pair-sharded gather ships
        self.best_idx = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
        self.best_count = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
        self.F = torch.empty((max(P, 1), 9), dtype=torch.float64, device=dev)
        self.mask = torch.empty(max(n_total, 1), dtype=torch.uint8, device=dev) if want_mask else None
        self.key = torch.empty(max(P, 1), dtype=torch.int64, device=dev) if want_key else None
        self.P = P


def f_ransac(d_pts, pair_off, d_idx, hyp_off, out: FOutputs, thr=1.5, mode=MODE_EPI_MAX, tie_mode=TIE_FIRST, solver=SOLVER_QR,
             score_path=SCORE_FP32_GUARDED, flags=0, seed=0, first_pair=0, hyp_first=0):
    """rg_f_ransac_dev2 on CUDA tensors.  d_idx None = samples drawn on the device from ``seed``."""
    lib = cabi.load_library()
    P = len(pair_off) - 1
    ctx = cabi.context(_dev_index(d_pts))
    cabi.check(lib.rg_f_ransac_dev2(
        _vp(ctx), _vp(_stream()), P, _vp(d_pts.data_ptr()), pair_off.ctypes.data_as(_pi32),
        _vp(d_idx.data_ptr()) if d_idx is not None else None, hyp_off.ctypes.data_as(_pi32), float(thr), int(mode),
        int(tie_mode), int(solver), int(score_path), int(flags), int(seed), int(first_pair), int(hyp_first),
        _vp(out.best_idx.data_ptr()), _vp(out.best_count.data_ptr()), _vp(out.F.data_ptr()),
        _vp(out.mask.data_ptr()) if out.mask is not None else None, _vp(out.key.data_ptr()) if out.key is not None else None))
    return out


def f_inlier_mask(d_pts, pair_off, d_F, thr=1.5, mode=MODE_EPI_MAX, out=None):
    """uint8 inlier mask of one F per pair (reference criterion), all on the device."""
    torch = _torch()
    lib = cabi.load_library()
    P = len(pair_off) - 1
    if out is None:
        out = torch.empty(max(int(pair_off[-1]), 1), dtype=torch.uint8, device=d_pts.device)
    cabi.check(lib.rg_f_inlier_mask_dev(_vp(cabi.context(_dev_index(d_pts))), _vp(_stream()), P, _vp(d_pts.data_ptr()),
                                        pair_off.ctypes.data_as(_pi32), _vp(d_F.data_ptr()), float(thr), int(mode),
                                        _vp(out.data_ptr())))
    return out


class PnpOutputs:
    def __init__(self, V: int, n_total: int = 0, device=None, want_mask: bool = False, want_key: bool = False):
        torch = _torch()
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.best_idx = torch.empty(max(V, 1), dtype=torch.int32, device=dev)
        self.best_count = torch.empty(max(V, 1), dtype=torch.int32, device=dev)
        self.Rt = torch.empty((max(V, 1), 12), dtype=torch.float64, device=dev)
        self.mask = torch.empty(max(n_total, 1), dtype=torch.uint8, device=dev) if want_mask else None
        self.key = torch.empty(max(V, 1), dtype=torch.int64, device=dev) if want_key else None


def pnp_ransac(d_X, d_y, view_off, d_idx, hyp_off, out: PnpOutputs, thr2, n=6, n_vote=None, score_path=SCORE_FP32_GUARDED,
               hyp_first=0):
    """rg_pnp_ransac_batched_dev2 on CUDA tensors."""
    lib = cabi.load_library()
    V = len(view_off) - 1
    vote = None if n_vote is None else np.ascontiguousarray(n_vote, dtype=np.int32)
    cabi.check(lib.rg_pnp_ransac_batched_dev2(
        _vp(cabi.context(_dev_index(d_X))), _vp(_stream()), V, _vp(d_X.data_ptr()), _vp(d_y.data_ptr()),
        view_off.ctypes.data_as(_pi32), vote.ctypes.data_as(_pi32) if vote is not None else None, _vp(d_idx.data_ptr()),
        hyp_off.ctypes.data_as(_pi32), int(n), float(thr2), int(score_path), int(hyp_first), _vp(out.best_idx.data_ptr()),
        _vp(out.best_count.data_ptr()), _vp(out.Rt.data_ptr()), _vp(out.mask.data_ptr()) if out.mask is not None else None,
        _vp(out.key.data_ptr()) if out.key is not None else None))
    return out


class P2PExchange:
    """The cross-GPU argmax of the hypothesis-split mode as ONE kernel over NVLink peer memory (csrc/p2p_api.cu) instead of
    NCCL collectives.  Collective constructor: every rank of ``group`` builds it at the same point; the 64-byte CUDA IPC
    handles are all-gathered through torch.distributed (any backend)."""

    def __init__(self, group=None, device=None):
        torch = _torch()
        import torch.distributed as dist
        self.lib = cabi.load_library()
        self.dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.ctx = cabi.context(self.dev.index)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # every step below is followed by a collective that ALL ranks reach even when a local CUDA call failed, so that a
        # rank whose IPC mapping is refused (containers without shared IPC namespaces) makes everybody raise, not hang
        h = (C.c_char * 64)()
        err = None
        try:
            cabi.check(self.lib.rg_p2p_create(_vp(self.ctx), self.rank, self.world, h))
        except Exception as e:           # noqa: BLE001 - reported to every rank below
            err = repr(e)
        mine = (err, bytes(h.raw))
        got = [mine]
        if self.world > 1:
            got = [None] * self.world
            dist.all_gather_object(got, mine, group=group)
        if err is None and all(g[0] is None for g in got):
            try:
                cabi.check(self.lib.rg_p2p_connect(_vp(self.ctx), b"".join(g[1] for g in got)))
            except Exception as e:       # noqa: BLE001
                err = repr(e)
        errs = [err]
        if self.world > 1:
            errs = [None] * self.world
            dist.all_gather_object(errs, err, group=group)          # also the barrier: every peer mapped every buffer
        bad = [(r, e) for r, e in enumerate(errs) if e is not None] + [(r, g[0]) for r, g in enumerate(got) if g[0] is not None]
        if bad:
            self.lib.rg_p2p_destroy(_vp(self.ctx))
            raise cabi.RGError(f"peer-memory exchange unavailable (rank {bad[0][0]}: {bad[0][1]})")
        self.status = torch.zeros(1, dtype=torch.int32, device=self.dev)

    def argmax(self, key, payload, best_idx, best_count, payload_out):
        """key (P,) int64, payload (P, npay) float64 -> winner's global index / count / payload on every rank (async)."""
        P = key.numel()
        npay = payload.shape[1] if payload is not None else 0
        cabi.check(self.lib.rg_p2p_argmax_exchange(
            _vp(self.ctx), _vp(_stream()), P, _vp(key.data_ptr()), _vp(payload.data_ptr()) if npay else None, npay,
            _vp(best_idx.data_ptr()), _vp(best_count.data_ptr()), _vp(payload_out.data_ptr()) if npay else None,
            _vp(self.status.data_ptr())))

    def check(self):
        s = int(self.status.item())
        if s:
            raise cabi.RGError(f"p2p argmax exchange: rank {s - 1} did not arrive (timeout)")

    def close(self):
        self.lib.rg_p2p_destroy(_vp(self.ctx))
