"""One rank's share of the config-3 hypothesis split at 8 GPUs (2048 of 16384 hypotheses x 100000 correspondences) on ONE GPU:
the launch chain SplitHypothesesF.run() issues (exchange = none), for the ncu launch list and for event timing."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsbb15_b200 import device as dv, parallel
d3, _ = dv.synth_two_view(1, 100000, first_pair=0, seed_base=3000)
sp = parallel.SplitHypothesesF(d3[0], 16384, sample_seed=20261018)
W = int(os.environ.get("R2_WORLD", "8"))          # pretend to be rank W//2 of W
sp.lo, sp.hi = (W // 2) * (16384 // W), (W // 2 + 1) * (16384 // W)
sp.hyp_off = np.array([0, 16384 // W], dtype=np.int32)
for _ in range(3):
    sp.run(thr=1.5, want_mask=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    sp.run(thr=1.5, want_mask=True)
e1.record(); e1.synchronize()
print("world", W, "gps", os.environ.get("RG_FORCE_GPS", "planner"), "plain ms", e0.elapsed_time(e1) / 50)
sp.capture(thr=1.5, want_mask=True)
e0.record()
for _ in range(50):
    sp.replay()
e1.record(); e1.synchronize()
print("graph ms", e0.elapsed_time(e1) / 50, sp.result()["best_count"])
