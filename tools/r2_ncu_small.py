"""Two device-resident config-5 calls (16 pairs) + one config-4 PnP call: the launch list for ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tsbb15_b200 import device as dv, sampling, synth
P = int(os.environ.get("R2_PAIRS", "16"))
d_pts, _ = dv.synth_two_view(P, 50000)
o = dv.FOutputs(P, P * 50000, want_mask=True)
po, ho = dv.offsets(np.full(P, 50000)), dv.offsets(np.full(P, 8192))
for _ in range(2):
    dv.f_ransac(d_pts, po, None, ho, o, seed=5)
torch.cuda.synchronize()
N4, H4 = 1000000, 8192
X4, y4, _ = synth.pnp_scene(N4, seed=4)
pi4 = sampling.fast(N4, H4, 6, seed=2)
dev = torch.device("cuda", 0)
dX, dy, dI = (torch.from_numpy(v).to(dev) for v in (X4, y4, pi4))
out = dv.PnpOutputs(1, N4, want_mask=True)
for _ in range(2):
    dv.pnp_ransac(dX, dy, np.array([0, N4], np.int32), dI, np.array([0, H4], np.int32), out, (1.5 / 3217.0) ** 2)
torch.cuda.synchronize()
print("ok", int(o.best_count.min().item()), int(out.best_count.item()))
