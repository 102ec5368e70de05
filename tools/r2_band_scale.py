"""What would a less accurate scorer cost in fix-up time?  The guard band is widened by option 9 (results stay exact: a wider
band only sends more evaluations to the FP64 recheck) on the bench's pass shape (64 pairs x 50 000 x 8 192)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsbb15_b200 import device as dv, runtime as rt
P = 16
d_pts, _ = dv.synth_two_view(P, 50000)
o = dv.FOutputs(P, P * 50000, want_mask=True)
po, ho = dv.offsets(np.full(P, 50000)), dv.offsets(np.full(P, 8192))
st = torch.cuda.current_stream().cuda_stream
ref = None
for scale in (1, 2, 4, 8, 16, 32, 64):
    rt.set_option(9, int(scale * 1000))
    dv.f_ransac(d_pts, po, None, ho, o, seed=5); torch.cuda.synchronize()
    rt.set_option(1, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dv.f_ransac(d_pts, po, None, ho, o, seed=5)
    e1.record(); e1.synchronize()
    pr = rt.profile(stream=st); rt.set_option(1, 0)
    s = rt.last_stats(stream=st)
    cnt = o.best_count.cpu().numpy().copy()
    ref = cnt if ref is None else ref
    c = max(pr["calls"], 1)
    print(json.dumps({"band_scale": scale, "ms": e0.elapsed_time(e1) / 5, "score_ms": pr["score_ms"] / c, "fixup_ms": pr["fixup_ms"] / c,
                      "flagged_groups": s["recheck_groups"], "band_evals": s["band_evals"], "overflow_hyps": s["overflow"],
                      "band_eval_fraction": s["band_evals"] / (P * 50000 * 8192.0), "same_winners": bool(np.array_equal(cnt, ref))}), flush=True)
rt.set_option(9, 1000)
