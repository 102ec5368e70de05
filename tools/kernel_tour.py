"""Small end-to-end invocations of every kernel family in one process (bundle adjustment in both placements of the camera
system and with split blocks, F-RANSAC with the tie replay, gold standard, PnP-RANSAC, triangulation): a quick tour for
profilers and for checking a new build by hand.  `python tools/kernel_tour.py [all|ba|f]`"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tsbb15_b200 as rg  # noqa: E402
from tsbb15_b200 import runtime as rt, sampling, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "ba"):
    g = np.load(os.path.join(ROOT, "tests", "golden", "ba_golden.npz"))
    for nv in (3, 8, 36):
        sc = tuple(g[f"v{nv}_" + k] for k in ("cams0", "pts0", "uv", "cam_idx", "pt_idx"))
        print("ba", nv, rt.bundle_adjust(*sc, max_iter=3)["cost"])
    sc = synth.ba_scene(60, 1500, track=8)
    print("ba 60 (L2 variant, split blocks)", rt.bundle_adjust(*sc[:5], max_iter=2)["cost"])
    rt.set_option(4, 1)
    print("ba 8 L2", rt.bundle_adjust(*tuple(g["v8_" + k] for k in ("cams0", "pts0", "uv", "cam_idx", "pt_idx")), max_iter=2)["cost"])
    rt.set_option(4, 0)
if which in ("all", "f"):
    pairs = synth.multi_pair(2, 3000)
    idx = [sampling.fast(3000, 700, 8, seed=p) for p in range(2)]
    r = rt.f_ransac_batched(pairs, idx, thr=1.5, want_counts=True, tie_mode=rg.TIE_REFERENCE)
    print("f", r["best_count"])
    gs = rt.gold_standard([pairs[0][r["mask"][0] > 0]], r["F"][:1], max_iter=5)
    print("gs", gs["cost"])
    X, y, _ = synth.pnp_scene(2000, seed=3)
    print("pnp", rt.pnp_ransac(X, y, sampling.fast(2000, 300, 6, seed=1), (1.5 / 3217.0) ** 2)["best_count"])
    d = np.load(os.path.join(ROOT, "tests", "golden", "geom_golden.npz"))
    print("tri", rt.triangulate(d["tri_C1"], d["tri_C2"], [d["tri_x1"]], [d["tri_x2"]])[0].shape)
