"""One rank's share of the config-3 hypothesis split at 8 GPUs (2048 of 16384 hypotheses x 100000 correspondences) on ONE GPU:
the launch chain SplitHypothesesF.run() issues (exchange = none), for the ncu launch list and for event timing."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsbb15_b200 import device as dv, parallel
d3, _ = dv.synth_two_view(1, 100000, first_pair=0, seed_base=3000)
sp = parallel.SplitHypothesesF(d3[0], 16384, sample_seed=20261018)
sp.lo, sp.hi = 3 * 2048, 4 * 2048                 # pretend to be rank 3 of 8
sp.hyp_off = np.array([0, 2048], dtype=np.int32)
for _ in range(3):
    sp.run(thr=1.5, want_mask=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    sp.run(thr=1.5, want_mask=True)
e1.record(); e1.synchronize()
print("plain ms", e0.elapsed_time(e1) / 50)
sp.capture(thr=1.5, want_mask=True)
e0.record()
for _ in range(50):
    sp.replay()
e1.record(); e1.synchronize()
print("graph ms", e0.elapsed_time(e1) / 50, sp.result()["best_count"])
