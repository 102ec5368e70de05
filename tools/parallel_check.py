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
    # ---- device-resident decompositions (round 2): pair-sharded sweep with gather, hypothesis split with the peer-memory
    # exchange kernel and with NCCL, F and PnP — each against a single-GPU run of the whole problem
    from tsbb15_b200 import device as dv
    detail = {}
    Pn, Nn, Hn = 11, 4000, 1024
    sh = parallel.PairShardedRansac(Pn, Nn, Hn, device=local, want_mask=True)
    sh.generate(seed_base=1000)
    sh.run(thr=1.5, sample_seed=77)
    got = sh.unpack(sh.gather(masks=True))
    full, _ = dv.synth_two_view(Pn, Nn, first_pair=0, seed_base=1000, device=local)
    o = dv.FOutputs(Pn, Pn * Nn, device=local, want_mask=True)
    dv.f_ransac(full, dv.offsets(np.full(Pn, Nn)), None, dv.offsets(np.full(Pn, Hn)), o, seed=77)
    ok5 = (np.array_equal(got["best_idx"], o.best_idx.cpu().numpy()) and np.array_equal(got["best_count"], o.best_count.cpu().numpy())
           and np.array_equal(got["F"].reshape(Pn, 9), o.F.cpu().numpy())
           and np.array_equal(np.concatenate(got["mask"]), o.mask.cpu().numpy()))
    sh.alloc_host()
    sh.h_pts[: sh.P].copy_(sh.d_pts)
    hb = sh.run_host(thr=1.5, sample_seed=77).numpy().copy()
    ok5 = ok5 and np.array_equal(hb, sh.gather(masks=False)["block"].cpu().numpy())
    detail["pair_sharded_device"] = bool(ok5)
    d1, _ = dv.synth_two_view(1, 30000, first_pair=0, seed_base=3000, device=local)
    o1 = dv.FOutputs(1, 30000, device=local, want_mask=True)
    dv.f_ransac(d1, dv.offsets([30000]), None, dv.offsets([4099]), o1, seed=5)
    ok6 = True
    for mode in ("p2p", "nccl"):
        sp = parallel.SplitHypothesesF(d1[0], 4099, exchange=mode, sample_seed=5)
        for _ in range(3):                              # repeated calls: sequence numbers / double buffering / prepared points
            sp.run(thr=1.5, want_mask=True)
        r = sp.result()
        good = (r["best_idx"] == int(o1.best_idx.item()) and r["best_count"] == int(o1.best_count.item())
                and np.array_equal(r["F"].reshape(9), o1.F.cpu().numpy().reshape(9)) and np.array_equal(r["mask"], o1.mask.cpu().numpy()))
        detail["split_f_" + mode] = bool(good)
        ok6 = ok6 and good
        if sp.p2p is not None:
            sp.p2p.close()
    dXp, dyp = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    for mode in ("p2p", "nccl"):
        spp = parallel.SplitHypothesesPnp(dXp, dyp, pidx, n=6, exchange=mode)
        spp.run(thr2); spp.run(thr2)
        r = spp.result()
        good = (r["best_idx"] == pone["best_idx"] and r["best_count"] == pone["best_count"] and np.array_equal(r["R"], pone["R"])
                and np.array_equal(r["t"], pone["t"]))
        detail["split_pnp_" + mode] = bool(good)
        ok6 = ok6 and good
        if spp.p2p is not None:
            spp.p2p.close()
    ok2 = ok2 and ok5 and ok6
    res = torch.tensor([int(ok1), int(ok2)], device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "pairs_sharded_ok": bool(res[0].item()), "hyp_split_ok": bool(res[1].item()),
                          "c_abi_allreduce_ok": bool(ok_c), "device_paths": detail, "split_owner": split["owner"], "best": split["best_idx"], "count": split["best_count"]}))
    dist.destroy_process_group()
    sys.exit(0 if int(res.min().item()) == 1 else 1)


if __name__ == "__main__":
    main()
