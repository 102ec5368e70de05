"""Dino sequence (config 2: 35 pairs x 10 000 hypotheses, ~300 correspondences each) through the host entry point: host-call time
against the number of sub-batches (option 2), which are pipelined over two streams since round 2 (option 10)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tsbb15_b200 import runtime as rt, synth

pairs = [np.ascontiguousarray(np.hstack(synth.dino_noisy_pair(i, i + 1))) for i in range(35)]


def host_time(fn, reps=30):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps, r


ref = None
for pipe in (1, 0):
    rt.set_option(10, pipe)
    for slices in (0, 1, 2, 3, 4, 6, 8):
        rt.set_option(2, slices)
        dt, r = host_time(lambda: rt.f_ransac_batched(pairs, None, n_hyp=10000, sample_seed=7, thr=1.5))
        if ref is None:
            ref = r
        same = bool(np.array_equal(r["best_idx"], ref["best_idx"]) and np.array_equal(r["best_count"], ref["best_count"]))
        print("pipeline %d  sub-batches %d (0 = automatic): %.4f ms per host call, passes %d, same result %s"
              % (pipe, slices, dt * 1e3, rt.last_stats()["passes"], same))
rt.set_option(2, 0)
rt.set_option(10, 1)
