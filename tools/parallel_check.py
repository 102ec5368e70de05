"""Launched under torchrun (one rank per GPU): checks the pair-sharded and the hypothesis-split multi-GPU paths of
parallel.py against a single-GPU run of the same problem (NCCL collectives, real CUDA library)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tsbb15_b200 as rg  # noqa: E402
from tsbb15_b200 import parallel, runtime as rt, sampling, synth  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["RG_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    pairs = synth.multi_pair(7, 3000)
    idxs = [sampling.fast(3000, 600, 8, seed=p) for p in range(7)]
    ref = rt.f_ransac_batched(pairs, idxs, thr=1.5, device=local)
    shard = parallel.f_ransac_pairs_sharded(pairs, idxs, thr=1.5, device=local)
    ok1 = (np.array_equal(shard["best_idx"], ref["best_idx"]) and np.array_equal(shard["best_count"], ref["best_count"])
           and np.allclose(shard["F"], ref["F"], rtol=0, atol=0))
    pts, _ = synth.two_view(20000, seed=5)
    idx = sampling.fast(20000, 4096, 8, seed=6)
    one = rt.f_ransac_batched([pts], [idx], thr=1.5, device=local)
    split = parallel.f_ransac_split_hypotheses(pts, idx, thr=1.5, device=local)
    ok2 = (split["best_idx"] == int(one["best_idx"][0]) and split["best_count"] == int(one["best_count"][0])
           and np.array_equal(split["mask"], one["mask"][0]) and np.array_equal(split["F"], one["F"][0]))
    # the same split with the key reduced by the library's own C entry point (rg_argmax_pack_dev / rg_argmax_allreduce /
    # rg_argmax_unpack_dev on a raw ncclComm_t) instead of torch.distributed
    comm = parallel.NcclComm(rank, world)
    split_c = parallel.f_ransac_split_hypotheses(pts, idx, thr=1.5, device=local, nccl_comm=comm)
    ok_c = (split_c["best_idx"] == split["best_idx"] and split_c["best_count"] == split["best_count"]
            and np.array_equal(split_c["mask"], split["mask"]) and np.array_equal(split_c["F"], split["F"]))
    # many keys in one reduction: every rank contributes its own (index, count) per pair
    rs = np.random.default_rng(100 + rank)
    bi = rs.integers(-1, 500, 64).astype(np.int32)
    bc = np.where(bi >= 0, rs.integers(1, 1000, 64), 0).astype(np.int32)
    gi, gc = parallel.argmax_allreduce_c(bi, bc, 1000 * rank, comm)
    allb = [torch.zeros(128, dtype=torch.int32, device="cuda") for _ in range(world)]
    dist.all_gather(allb, torch.from_numpy(np.concatenate([bi, bc])).cuda())
    keys = np.zeros(64, dtype=np.uint64)
    for r, t_ in enumerate(allb):
        a = t_.cpu().numpy()
        k_ = np.array([parallel.argmax_key(int(c_), int(i_) + 1000 * r) if i_ >= 0 and c_ > 0 else 0
                       for i_, c_ in zip(a[:64], a[64:])], dtype=np.uint64)
        keys = np.maximum(keys, k_)
    exp = [parallel.key_decode(int(k_)) if k_ else (0, -1) for k_ in keys]
    ok_c = ok_c and all(int(gc[p]) == e[0] and int(gi[p]) == e[1] for p, e in enumerate(exp))
    comm.close()
    ok2 = ok2 and ok_c
    # PnP: hypotheses of one view split over the ranks; views sharded over the ranks
    thr2 = (1.5 / 3217.0) ** 2
    X, y, _ = synth.pnp_scene(30000, seed=4)
    pidx = sampling.fast(30000, 2048, 6, seed=2)
    pone = rt.pnp_ransac(X, y, pidx, thr2, device=local)
    psplit = parallel.pnp_ransac_split_hypotheses(X, y, pidx, thr2, device=local)
    ok3 = (psplit["best_idx"] == pone["best_idx"] and psplit["best_count"] == pone["best_count"]
           and np.array_equal(psplit["mask"], pone["mask"]) and np.array_equal(psplit["R"], pone["R"]))
    views = [synth.pnp_scene(500 + 37 * v, seed=v)[:2] for v in range(5)]
    vidx = [sampling.fast(len(v[0]), 256, 6, seed=k) for k, v in enumerate(views)]
    vref = rt.pnp_ransac_batched([v[0] for v in views], [v[1] for v in views], vidx, thr2, device=local)
    vsh = parallel.pnp_ransac_views_sharded([v[0] for v in views], [v[1] for v in views], vidx, thr2, device=local)
    ok4 = (np.array_equal(vsh["best_idx"], vref["best_idx"]) and np.array_equal(vsh["best_count"], vref["best_count"])
           and np.array_equal(vsh["R"], vref["R"], equal_nan=True))
    ok2 = ok2 and ok3 and ok4
    res = torch.tensor([int(ok1), int(ok2)], device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "pairs_sharded_ok": bool(res[0].item()), "hyp_split_ok": bool(res[1].item()),
                          "c_abi_allreduce_ok": bool(ok_c), "split_owner": split["owner"], "best": split["best_idx"], "count": split["best_count"]}))
    dist.destroy_process_group()
    sys.exit(0 if int(res.min().item()) == 1 else 1)


if __name__ == "__main__":
    main()
