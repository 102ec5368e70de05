"""Where does the 'prepare' phase of a device-resident call go?  Variants: seeded / explicit indices / reuse of prepared points."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsbb15_b200 import device as dv, runtime as rt
P = 16
d_pts, _ = dv.synth_two_view(P, 50000)
o = dv.FOutputs(P, P * 50000, want_mask=True)
po, ho = dv.offsets(np.full(P, 50000)), dv.offsets(np.full(P, 8192))
d_idx = dv.sample_indices(np.full(P, 50000), 8192, 8, seed=5)
st = torch.cuda.current_stream().cuda_stream
def run(name, fn, reps=10):
    fn(); torch.cuda.synchronize()
    rt.set_option(1, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    pr = rt.profile(stream=st); rt.set_option(1, 0)
    c = max(pr["calls"], 1)
    print(json.dumps({"variant": name, "ms": e0.elapsed_time(e1) / reps, **{k: round(pr[k] / c, 4) for k in pr if k != "calls"}}), flush=True)
run("seeded", lambda: dv.f_ransac(d_pts, po, None, ho, o, seed=5))
run("explicit_idx", lambda: dv.f_ransac(d_pts, po, d_idx, ho, o))
run("explicit_idx_reuse", lambda: dv.f_ransac(d_pts, po, d_idx, ho, o, flags=1))
run("explicit_idx_nomask", lambda: dv.f_ransac(d_pts, po, d_idx, ho, dv.FOutputs(P)))
# one pair, config 3 shape
d3, _ = dv.synth_two_view(1, 100000)
o3 = dv.FOutputs(1, 100000, want_mask=True, want_key=True)
po3, ho3 = dv.offsets([100000]), dv.offsets([16384])
run("config3_seeded", lambda: dv.f_ransac(d3, po3, None, ho3, o3, seed=5))
run("config3_seeded_reuse", lambda: dv.f_ransac(d3, po3, None, ho3, o3, seed=5, flags=1))
ho3b = dv.offsets([2048])
run("config3_eighth_reuse", lambda: dv.f_ransac(d3, po3, None, ho3b, o3, seed=5, flags=1, hyp_first=2048))
