"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in parallel.py; the compute kernel is replaced by the
oracle so that only sharding, key packing and the collectives are exercised here."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_compute(pts_list, idx_list, thr=1.5, **kw):
    from oracle import f_path as orc
    P = len(pts_list)
    out = {"best_idx": np.full(P, -1, np.int32), "best_count": np.zeros(P, np.int32), "F": np.full((P, 3, 3), np.nan),
           "mask": []}
    for p in range(P):
        pts = np.asarray(pts_list[p])
        r = orc.f_ransac(pts[:, :2].T.copy(), pts[:, 2:].T.copy(), idx_list[p], thr, tie="first")
        out["best_idx"][p], out["best_count"][p] = r["best"], r["counts"].max() if r["best"] >= 0 else 0
        if r["best"] >= 0:
            out["F"][p] = r["F"]
        out["mask"].append(r["mask"])
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import tsbb15_b200 as rg
    from tsbb15_b200 import parallel, sampling, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pairs = synth.multi_pair(5, 300)
        idxs = [sampling.fast(300, 40, 8, seed=p) for p in range(5)]
        res = parallel.f_ransac_pairs_sharded(pairs, idxs, thr=1.5, compute=_oracle_compute)
        split = parallel.f_ransac_split_hypotheses(pairs[0], idxs[0], thr=1.5, compute=_oracle_compute)
        q.put((rank, res["range"], res["best_idx"].tolist(), res["best_count"].tolist(), res["F"].tolist(),
               split["best_idx"], split["best_count"], split["F"].tolist(), split["mask"].tolist(), split["owner"]))
    finally:
        dist.destroy_process_group()


def test_shard_range_and_keys(rg):
    par = rg.parallel
    for n in (0, 1, 5, 8, 4096):
        for w in (1, 2, 3, 8):
            rs = [par.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
    assert par.argmax_key(10, 5) > par.argmax_key(10, 6) > par.argmax_key(9, 0)
    assert par.key_decode(par.argmax_key(123, 4567)) == (123, 4567)


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from tsbb15_b200 import sampling, synth
    pairs = synth.multi_pair(5, 300)
    idxs = [sampling.fast(300, 40, 8, seed=p) for p in range(5)]
    ref = _oracle_compute(pairs, idxs, 1.5)
    assert outs[0][1] == (0, 3) and outs[1][1] == (3, 5)
    for o in outs:                                        # both ranks hold the full, identical, correct result
        assert o[2] == ref["best_idx"].tolist() and o[3] == ref["best_count"].tolist()
        assert np.allclose(np.array(o[4]), ref["F"])
        assert o[5] == int(ref["best_idx"][0]) and o[6] == int(ref["best_count"][0])
        assert np.allclose(np.array(o[7]), ref["F"][0])
        assert o[8] == ref["mask"][0].tolist()
    assert outs[0][9] == outs[1][9] in (0, 1)
