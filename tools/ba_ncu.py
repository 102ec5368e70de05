"""One device bundle adjustment of the 36-view golden Dino scene (for ncu launch lists / captures)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tsbb15_b200 as rg  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "ba_golden.npz"))
p = "v36_"
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 12
r = rg.runtime.bundle_adjust(g[p + "cams0"], g[p + "pts0"], g[p + "uv"], g[p + "cam_idx"], g[p + "pt_idx"], ftol=1e-4,
                             max_iter=n_iter)
print(r["cost"], r["iters"], r["status"])
