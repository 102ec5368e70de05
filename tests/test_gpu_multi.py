"""Multi-GPU host path on real hardware: world size 1 always; world size 2 under torchrun when two GPUs are visible
(the world-size-2 logic itself is covered on CPU by tests/test_parallel_gloo.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parallel_api_world_size_1(rg):
    pairs = rg.synth.multi_pair(3, 2000)
    idxs = [rg.sampling.fast(2000, 300, 8, seed=p) for p in range(3)]
    ref = rg.runtime.f_ransac_batched(pairs, idxs, thr=1.5)
    shard = rg.parallel.f_ransac_pairs_sharded(pairs, idxs, thr=1.5)
    assert shard["range"] == (0, 3)
    assert np.array_equal(shard["best_idx"], ref["best_idx"]) and np.array_equal(shard["F"], ref["F"])
    one = rg.parallel.f_ransac_split_hypotheses(pairs[0], idxs[0], thr=1.5)
    assert one["best_idx"] == int(ref["best_idx"][0]) and np.array_equal(one["mask"], ref["mask"][0])


def test_two_gpus_nccl_if_available():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tools", "parallel_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and d["pairs_sharded_ok"] and d["hyp_split_ok"] and d["c_abi_allreduce_ok"]
