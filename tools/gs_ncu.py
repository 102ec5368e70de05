"""One batched device gold-standard refinement (16 pairs x 50 000 correspondences) for ncu launch lists."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tsbb15_b200 as rg  # noqa: E402
from oracle import geom_path as og  # noqa: E402

d = np.load(os.path.join(ROOT, "tsbb15-3d-reconstruction-project_b200", "data", "dino_data.npz"))
Ps = d["Ps"]
rng = np.random.default_rng(0)
P, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 50000)
pairs, Fs = [], []
for k in range(P):
    Xs = np.column_stack([rng.uniform(-0.04, 0.04, N), rng.uniform(-0.07, 0.02, N), rng.uniform(-0.7, -0.56, N), np.ones(N)])
    a, b = Xs @ Ps[k].T, Xs @ Ps[k + 1].T
    pairs.append(np.hstack([a[:, :2] / a[:, 2:], b[:, :2] / b[:, 2:]]) + rng.normal(0, 0.5, (N, 4)))
    Fs.append(og.fmatrix_from_cameras(Ps[k], Ps[k + 1]))
r = rg.runtime.gold_standard(pairs, np.stack(Fs))
print(r["iters"], r["status"], (r["cost"] / N)[:3])
