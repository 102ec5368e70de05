"""The reference's main.py (lines 30-181: INIT1-3, then for every further view BA -> EXT1-5) replayed through the drop-in
modules on the clean Dino tracks, timed.  main.py itself cannot run here or on the GPU box (it loads ../images/*.ppm and
plots); this script follows it statement by statement without the image / plot calls.

    python tools/run_main_dropin.py [last_view=34] [--no-ba]

Data: the exact BAdino2.mat projections, i.e. the branch of correspondences.py that the reference ships active (its
tracked points.txt branch is commented out: those tracks contain gross mismatches, INIT3 triangulates every putative
correspondence and BundleAdjustment2 has no robust loss, so the pipeline is not meant to run on them).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(last_view=34, bundle_adjust=True, r_f=10000, verbose=True):
    import tsbb15_b200 as rg
    fun = rg.fun
    Tables, CameraPose = rg.tables.Tables, rg.help_classes.CameraPose
    d = np.load(os.path.join(ROOT, "tsbb15-3d-reconstruction-project_b200", "data", "dino_data.npz"))
    x2d, Ps = d["x2d"], d["Ps"]
    pnp_kw = dict(r=256, reproj_px=1.5)

    def corr(i1, i2):                                         # correspondences.getCorrByIndices (clean branch)
        y1, y2 = x2d[i1].T, x2d[i2].T
        ok = np.logical_and(np.any(y1 != -1, axis=1), np.any(y2 != -1, axis=1))
        return np.array(y1[ok]), np.array(y2[ok])

    times = {"f_ransac": 0.0, "ba": 0.0, "add_view": 0.0, "add_points": 0.0, "init": 0.0}
    t0 = time.perf_counter()
    rg._cabi.context()                                        # CUDA context + library load: not part of the pipeline
    times["cuda_context (not in total)"] = time.perf_counter() - t0
    t_all = time.perf_counter()
    T = Tables()
    C = Ps[None]
    y1, y2 = corr(0, 1)
    t0 = time.perf_counter()
    F = fun.getFFromLabCode(y1.T, y2.T, r=r_f, seed=0, refine="device")          # main.py:39 (commented there: loads Fmatrix.npy)
    times["f_ransac"] += time.perf_counter() - t0
    t0 = time.perf_counter()
    E, K = fun.getEAndK(C, F)                                                     # main.py:54
    T.K = K
    y1h, y2h = fun.MakeHomogenous(K, y1), fun.MakeHomogenous(K, y2)
    R, t = fun.relative_camera_pose(E, y1h[0, :2].T, y2h[0, :2].T)                # main.py:62
    C1, C2 = CameraPose(), CameraPose(R, t)
    v1, v2 = T.addView(0, C1), T.addView(1, C2)
    T.triangulateAndAddPoints(v1, v2, C1, C2, y1h, y2h)                           # main.py:76
    times["init"] += time.perf_counter() - t0
    ba_log = []
    for i in range(1, last_view):                                                 # main.py:95
        if bundle_adjust:
            t0 = time.perf_counter()
            info = T.BundleAdjustment2()                                          # main.py:100
            times["ba"] += time.perf_counter() - t0
            ba_log.append((len(T.T_views), len(T.T_points), len(T.T_obs), info["iters"], info["cost"]))
        a, b = corr(i, i + 1)                                                     # main.py:113
        ah, bh = fun.MakeHomogenous(K, a), fun.MakeHomogenous(K, b)
        t0 = time.perf_counter()
        A1, A2 = T.addNewView(K, i + 1, ah, bh, a, b, seed=i, **pnp_kw)          # main.py:127
        times["add_view"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        T.addNewPoints(fun.MakeHomogenous(K, A1), fun.MakeHomogenous(K, A2), i, i + 1)  # main.py:137
        times["add_points"] += time.perf_counter() - t0
    total = time.perf_counter() - t_all
    Rs, ts = T.getCamerasForEvaluation()                                          # main.py:177
    if verbose:
        print("views", len(T.T_views), "points", len(T.T_points), "observations", len(T.T_obs))
        print("seconds: total %.3f | " % total + " | ".join(f"{k} {v:.3f}" for k, v in times.items()))
        if ba_log:
            print("last BA: views %d points %d obs %d iters %d cost %.3e" % ba_log[-1])
    return T, Rs, ts, times, total, ba_log


def rotation_errors_deg(Rs):
    """Angle between every recovered rotation (nearest rotation of the possibly non-orthonormal 3x3 block BA leaves) and
    the ground-truth rotation of that view relative to view 0 (fun.camera_resectioning of BAdino2.mat's cameras)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "pnp_golden.npz"))
    out = []
    for k, R in enumerate(Rs):
        U, _, Vt = np.linalg.svd(R)
        Rn = U @ np.diag([1.0, 1.0, np.linalg.det(U @ Vt)]) @ Vt
        Rg = g["R"][k] @ g["R"][0].T
        c = (np.trace(Rn @ Rg.T) - 1.0) / 2.0
        out.append(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))
    return np.array(out)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    T, Rs, ts, times, total, ba_log = run(int(args[0]) if args else 34, bundle_adjust="--no-ba" not in sys.argv)
    err = rotation_errors_deg(Rs)
    print("rotation error vs ground truth (deg): median %.2e max %.2e" % (np.median(err), err.max()))
