"""Experiment: phase breakdown (clock64 ticks seen by CTA 0) of the cluster-resident Cholesky; needs a -DRG_BA_PROF build
passed through RG_LIB."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tsbb15_b200 as rg  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "ba_golden.npz"))
p = "v36_"
lib = rg._cabi.load_library()
out = (C.c_longlong * 16)()
args = (g[p + "cams0"], g[p + "pts0"], g[p + "uv"], g[p + "cam_idx"], g[p + "pt_idx"])
rg.runtime.bundle_adjust(*args, ftol=1e-4, max_iter=2)
lib.rg_ba_prof_read(out)
n_it = 10
r = rg.runtime.bundle_adjust(*args, ftol=0.0, max_iter=n_it)
lib.rg_ba_prof_read(out)
names = ["load+sync", "owner factor", "owner panel", "cluster.sync wait", "panel fetch", "trailing", "backward"]
tot = sum(out[:7])
for k, nm in enumerate(names):
    print(f"{nm:20s} {out[k] / n_it / 1.965e3:9.1f} us/solve  {100.0 * out[k] / tot:5.1f} %")
print("total", tot / n_it / 1.965e3, "us/solve (at 1.965 GHz)", r["iters"])
