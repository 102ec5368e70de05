"""Multi-GPU sharding of the RANSAC path: one process per GPU, ``torch.distributed`` (NCCL over NVLink) as plumbing.

Two decompositions:

* by image pair (``f_ransac_pairs_sharded``): pairs are independent, rank g takes a contiguous block of pairs, there is
  NO collective on the data path; the tiny per-pair results (best index, count, F) are gathered afterwards.
* by hypothesis block inside one pair (``f_ransac_split_hypotheses``): every rank holds all N correspondences, scores
  its slice of the hypotheses, and ONE 8-byte max-all-reduce on the key ``(count << 32) | (0xFFFFFFFF - index)`` picks
  the global winner with the lowest index among equal counts (= first-maximum rule of fun.py:320); the owner of the
  winner broadcasts F and the inlier mask.

``compute`` can be injected so the distributed logic is testable on CPU with the gloo backend (tests/ use the oracle
there); by default it is the CUDA library.
"""
from __future__ import annotations

import numpy as np

from . import runtime as _rt


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition: the first ``n_units % world`` ranks get one extra unit."""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def argmax_key(count: int, index: int) -> int:
    """Monotone key: larger count wins, then the LOWER hypothesis index."""
    return (int(count) << 32) | (0xFFFFFFFF - int(index))


def key_decode(key: int) -> tuple[int, int]:
    return int(key) >> 32, 0xFFFFFFFF - (int(key) & 0xFFFFFFFF)


def _dist():
    import torch.distributed as dist
    return dist


def _world(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _device_for_collectives(group=None):
    import torch
    dist = _dist()
    if dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


class NcclComm:
    """A raw ``ncclComm_t`` for the C-ABI reduction ``rg_argmax_allreduce`` (include/rg_b200.h) — what a host that does
    not use torch would create itself.  The unique id is made on rank 0 and handed to the other ranks through
    ``exchange`` (default: ``torch.distributed.broadcast_object_list`` on whatever backend is initialised)."""

    def __init__(self, rank: int, world: int, exchange=None, lib_name: str | None = None):
        import ctypes as C
        import os
        self._C = C
        self.lib = C.CDLL(lib_name or os.environ.get("RG_NCCL_LIB") or "libnccl.so.2", mode=C.RTLD_GLOBAL)

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_char * 128)]

        self.lib.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        self.lib.ncclCommDestroy.argtypes = [C.c_void_p]
        uid = UniqueId()
        if rank == 0 and self.lib.ncclGetUniqueId(C.byref(uid)) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
        raw = bytes(C.string_at(C.byref(uid), 128))
        if exchange is None:
            def exchange(b):
                box = [b]
                _dist().broadcast_object_list(box, src=0)
                return box[0]
        raw = exchange(raw)
        C.memmove(C.byref(uid), raw, 128)
        self.handle = C.c_void_p()
        rc = self.lib.ncclCommInitRank(C.byref(self.handle), int(world), uid, int(rank))
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed ({rc})")
        self.rank, self.world = rank, world

    def close(self):
        if self.handle:
            self.lib.ncclCommDestroy(self.handle)
            self.handle = None


def argmax_allreduce_c(best_idx, best_count, index_offset, comm: NcclComm, stream=0):
    """(best_idx, best_count) of this rank's hypothesis block -> global winner, entirely through the C ABI on the device:
    rg_argmax_pack_dev -> rg_argmax_allreduce (ONE ncclAllReduce(max) of 8 bytes per pair) -> rg_argmax_unpack_dev."""
    import ctypes as C
    import torch
    from . import _cabi as cabi
    lib = cabi.load_library()
    vp = C.c_void_p
    bi = torch.as_tensor(np.atleast_1d(np.asarray(best_idx, dtype=np.int32))).cuda()
    bc = torch.as_tensor(np.atleast_1d(np.asarray(best_count, dtype=np.int32))).cuda()
    P = bi.numel()
    key = torch.empty(P, dtype=torch.int64, device=bi.device)
    st = stream or torch.cuda.current_stream().cuda_stream
    cabi.check(lib.rg_argmax_pack_dev(vp(st), P, vp(bi.data_ptr()), vp(bc.data_ptr()), int(index_offset), vp(key.data_ptr())))
    cabi.check(lib.rg_argmax_allreduce(comm.handle, vp(st), vp(key.data_ptr()), P))
    cabi.check(lib.rg_argmax_unpack_dev(vp(st), P, vp(key.data_ptr()), vp(bi.data_ptr()), vp(bc.data_ptr())))
    torch.cuda.current_stream().synchronize()
    return bi.cpu().numpy(), bc.cpu().numpy()


def f_ransac_pairs_sharded(pts_list, idx_list, thr=1.5, group=None, gather=True, compute=None, **kw) -> dict:
    """Pair-sharded batched F-RANSAC.  Every rank passes the SAME full lists (or at least its own shard filled in);
    returns, on every rank, arrays over all pairs when ``gather`` is True, else only the local shard's results."""
    rank, world = _world(group)
    P = len(pts_list)
    lo, hi = shard_range(P, rank, world)
    compute = compute or _rt.f_ransac_batched
    local = compute(pts_list[lo:hi], idx_list[lo:hi], thr=thr, **kw)
    out_local = {"range": (lo, hi), "best_idx": np.asarray(local["best_idx"]), "best_count": np.asarray(local["best_count"]),
                 "F": np.asarray(local["F"])}
    if not gather or world == 1:
        return out_local
    import torch
    dist = _dist()
    dev = _device_for_collectives(group)
    # fixed-size payload per pair: [best_idx, best_count, F(9)] as float64 -> one all_gather of padded blocks
    per = (P + world - 1) // world
    buf = torch.full((per, 11), float("nan"), dtype=torch.float64)
    n_loc = hi - lo
    if n_loc:
        buf[:n_loc, 0] = torch.from_numpy(out_local["best_idx"].astype(np.float64))
        buf[:n_loc, 1] = torch.from_numpy(out_local["best_count"].astype(np.float64))
        buf[:n_loc, 2:] = torch.from_numpy(out_local["F"].reshape(n_loc, 9))
    buf = buf.to(dev)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    best_idx = np.full(P, -1, dtype=np.int32)
    best_count = np.zeros(P, dtype=np.int32)
    F = np.full((P, 3, 3), np.nan)
    for r in range(world):
        a, b = shard_range(P, r, world)
        if b > a:
            blk = gathered[r][:b - a].cpu().numpy()
            best_idx[a:b] = blk[:, 0].astype(np.int32)
            best_count[a:b] = blk[:, 1].astype(np.int32)
            F[a:b] = blk[:, 2:].reshape(-1, 3, 3)
    return {"range": (lo, hi), "best_idx": best_idx, "best_count": best_count, "F": F}


def f_ransac_split_hypotheses(pts, idx, thr=1.5, group=None, compute=None, nccl_comm: NcclComm | None = None, **kw) -> dict:
    """One pair, hypotheses split across ranks, one max-all-reduce of an int64 key (first-maximum selection).
    nccl_comm: reduce through the library's own C entry point ``rg_argmax_allreduce`` on that communicator instead of
    ``torch.distributed.all_reduce`` (same key, same result)."""
    import torch
    rank, world = _world(group)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    H = idx.shape[0]
    lo, hi = shard_range(H, rank, world)
    compute = compute or _rt.f_ransac_batched
    kw = dict(kw)
    kw["want_mask"] = True
    local = compute([pts], [idx[lo:hi]], thr=thr, **kw)
    li = int(local["best_idx"][0])
    key = argmax_key(int(local["best_count"][0]), lo + li) if li >= 0 else 0
    if world == 1:
        cnt, gi = key_decode(key) if key else (0, -1)
        return {"best_idx": gi, "best_count": cnt, "F": local["F"][0], "mask": local["mask"][0], "owner": 0}
    dist = _dist()
    dev = _device_for_collectives(group)
    if nccl_comm is not None:
        gi_a, cnt_a = argmax_allreduce_c(local["best_idx"][:1], local["best_count"][:1], lo, nccl_comm)
        gkey = argmax_key(int(cnt_a[0]), int(gi_a[0])) if gi_a[0] >= 0 else 0
    else:
        k = torch.tensor([key], dtype=torch.int64, device=dev)
        dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
        gkey = int(k.item())
    if gkey == 0:
        n = np.asarray(pts).reshape(-1, 4).shape[0]
        return {"best_idx": -1, "best_count": 0, "F": np.full((3, 3), np.nan), "mask": np.zeros(n, np.uint8), "owner": -1}
    cnt, gi = key_decode(gkey)
    owner = next(r for r in range(world) if shard_range(H, r, world)[0] <= gi < shard_range(H, r, world)[1])
    n = np.asarray(pts).reshape(-1, 4).shape[0]
    payload = torch.zeros(9 + n, dtype=torch.float64)
    if rank == owner:
        payload[:9] = torch.from_numpy(np.asarray(local["F"][0]).reshape(9))
        payload[9:] = torch.from_numpy(np.asarray(local["mask"][0], dtype=np.float64))
    payload = payload.to(dev)
    dist.broadcast(payload, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
    payload = payload.cpu().numpy()
    return {"best_idx": gi, "best_count": cnt, "F": payload[:9].reshape(3, 3), "mask": payload[9:].astype(np.uint8),
            "owner": owner}


def pnp_ransac_split_hypotheses(X, y, idx, thr2, group=None, compute=None, **kw) -> dict:
    """PnP-RANSAC of ONE view (BASELINE config 4 at G > 1) with the pose hypotheses split across ranks: every rank holds
    all N correspondences and scores its slice of the (H, n) samples; the same 8-byte max-all-reduce on
    ``(count << 32) | ~index`` as the F path picks the winner (first maximum, ransac.py:108); its owner broadcasts
    R, t and the consensus mask."""
    import torch
    rank, world = _world(group)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    H = idx.shape[0]
    lo, hi = shard_range(H, rank, world)
    compute = compute or _rt.pnp_ransac
    kw = dict(kw)
    kw["want_mask"] = True
    n = np.asarray(X).reshape(-1, 3).shape[0]
    if hi > lo:
        local = compute(X, y, idx[lo:hi], thr2, **kw)
    else:
        local = {"best_idx": -1, "best_count": 0, "R": np.full((3, 3), np.nan), "t": np.full(3, np.nan),
                 "mask": np.zeros(n, np.uint8)}
    li = int(local["best_idx"])
    key = argmax_key(int(local["best_count"]), lo + li) if li >= 0 else 0
    if world == 1:
        cnt, gi = key_decode(key) if key else (0, -1)
        return {"best_idx": gi, "best_count": cnt, "R": local["R"], "t": local["t"], "mask": local["mask"], "owner": 0}
    dist = _dist()
    dev = _device_for_collectives(group)
    k = torch.tensor([key], dtype=torch.int64, device=dev)
    dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
    gkey = int(k.item())
    if gkey == 0:
        return {"best_idx": -1, "best_count": 0, "R": np.full((3, 3), np.nan), "t": np.full(3, np.nan),
                "mask": np.zeros(n, np.uint8), "owner": -1}
    cnt, gi = key_decode(gkey)
    owner = next(r for r in range(world) if shard_range(H, r, world)[0] <= gi < shard_range(H, r, world)[1])
    payload = torch.zeros(12 + n, dtype=torch.float64)
    if rank == owner:
        payload[:9] = torch.from_numpy(np.asarray(local["R"], dtype=np.float64).reshape(9))
        payload[9:12] = torch.from_numpy(np.asarray(local["t"], dtype=np.float64).reshape(3))
        payload[12:] = torch.from_numpy(np.asarray(local["mask"], dtype=np.float64))
    payload = payload.to(dev)
    dist.broadcast(payload, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
    payload = payload.cpu().numpy()
    return {"best_idx": gi, "best_count": cnt, "R": payload[:9].reshape(3, 3), "t": payload[9:12].copy(),
            "mask": payload[12:].astype(np.uint8), "owner": owner}


def pnp_ransac_views_sharded(X_list, y_list, idx_list, thr2, group=None, compute=None, **kw) -> dict:
    """View-sharded batched PnP-RANSAC (BASELINE config 2 at G > 1): views are independent, rank g takes a contiguous
    block, no data-path collective; (best index, count, R, t) of all views are gathered on every rank."""
    rank, world = _world(group)
    V = len(X_list)
    lo, hi = shard_range(V, rank, world)
    compute = compute or _rt.pnp_ransac_batched
    if hi > lo:
        local = compute(X_list[lo:hi], y_list[lo:hi], idx_list[lo:hi], thr2, **kw)
    else:
        local = {"best_idx": np.zeros(0, np.int32), "best_count": np.zeros(0, np.int32), "R": np.zeros((0, 3, 3)),
                 "t": np.zeros((0, 3))}
    out = {"range": (lo, hi), "best_idx": np.asarray(local["best_idx"]), "best_count": np.asarray(local["best_count"]),
           "R": np.asarray(local["R"]), "t": np.asarray(local["t"])}
    if world == 1:
        return out
    import torch
    dist = _dist()
    dev = _device_for_collectives(group)
    per = (V + world - 1) // world
    buf = torch.full((per, 14), float("nan"), dtype=torch.float64)
    n_loc = hi - lo
    if n_loc:
        buf[:n_loc, 0] = torch.from_numpy(out["best_idx"].astype(np.float64))
        buf[:n_loc, 1] = torch.from_numpy(out["best_count"].astype(np.float64))
        buf[:n_loc, 2:11] = torch.from_numpy(out["R"].reshape(n_loc, 9))
        buf[:n_loc, 11:] = torch.from_numpy(out["t"].reshape(n_loc, 3))
    buf = buf.to(dev)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    best_idx = np.full(V, -1, dtype=np.int32)
    best_count = np.zeros(V, dtype=np.int32)
    R = np.full((V, 3, 3), np.nan)
    t = np.full((V, 3), np.nan)
    for r in range(world):
        a, b = shard_range(V, r, world)
        if b > a:
            blk = gathered[r][:b - a].cpu().numpy()
            best_idx[a:b] = blk[:, 0].astype(np.int32)
            best_count[a:b] = blk[:, 1].astype(np.int32)
            R[a:b] = blk[:, 2:11].reshape(-1, 3, 3)
            t[a:b] = blk[:, 11:]
    return {"range": (lo, hi), "best_idx": best_idx, "best_count": best_count, "R": R, "t": t}
