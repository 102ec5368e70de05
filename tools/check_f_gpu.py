"""Development check (GPU box): F path vs the oracle on small shapes + raw timings at config-3 shape."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tsbb15_b200 as rg  # noqa: E402
from tsbb15_b200 import runtime as rt, sampling, synth  # noqa: E402
from oracle import f_path as orc  # noqa: E402


def check_small(solver):
    pts, _ = synth.two_view(3000, seed=11)
    idx = sampling.fast(pts.shape[0], 600, 8, seed=5)
    p1, p2 = pts[:, :2].T.copy(), pts[:, 2:].T.copy()
    t0 = time.time()
    o = orc.f_ransac(p1, p2, idx, 1.5, tie="first")
    t_or = time.time() - t0
    cond = orc.sample_condition(p1, p2, idx)
    res = rt.f_ransac_batched([pts], [idx], thr=1.5, solver=solver, want_counts=True, want_F_all=True, want_flags=True)
    st = rt.last_stats()
    cg = res["counts"][0]
    Fg = res["F_all"][0]
    ferr = np.array([np.linalg.norm(orc.normalise_F(Fg[h], orc.normalise_F(o["F_all"][h])) - orc.normalise_F(o["F_all"][h]))
                     for h in range(len(idx))])
    good = cond > 1e-6
    mism = np.flatnonzero(cg != o["counts"])
    print(json.dumps({"solver": solver, "oracle_s": round(t_or, 2), "count_mismatch": int(mism.size),
                      "count_mismatch_wellcond": int(np.count_nonzero(good[mism])) if mism.size else 0,
                      "max_F_err_wellcond": float(ferr[good].max()), "median_F_err": float(np.median(ferr)),
                      "flags": int(np.count_nonzero(res["flags"][0])), "best_gpu": int(res["best_idx"][0]),
                      "best_oracle": int(o["best"]), "mask_equal": bool(np.array_equal(res["mask"][0], o["mask"])),
                      "stats": st}))
    # scoring parity with identical F: feed the ORACLE's F to the GPU scorer
    c32 = rt.epi_score_count(pts, o["F_all"], 1.5)
    st2 = rt.last_stats()
    c64 = rt.epi_score_count(pts, o["F_all"], 1.5, score_path=rg.SCORE_FP64)
    print(json.dumps({"score_fp32_vs_oracle_mismatch": int(np.count_nonzero(c32 != o["counts"])),
                      "score_fp64_vs_oracle_mismatch": int(np.count_nonzero(c64 != o["counts"])), "stats": st2}))
    for thr in (0.25, 3.0):
        oc = orc.score_hypotheses(o["F_all"], p1, p2, thr)
        c32 = rt.epi_score_count(pts, o["F_all"], thr)
        cs = rt.epi_score_count(pts, o["F_all"], thr, mode=rg.MODE_SAMPSON)
        ocs = orc.score_hypotheses(o["F_all"], p1, p2, thr, mode=orc.SAMPSON)
        print(json.dumps({"thr": thr, "epi_mismatch": int(np.count_nonzero(c32 != oc)),
                          "sampson_mismatch": int(np.count_nonzero(cs != ocs)), "stats": rt.last_stats()}))


def timing():
    import ctypes as C
    pts, _ = synth.two_view(100000, seed=1)
    idx = sampling.fast(pts.shape[0], 16384, 8, seed=2)
    for solver in (rg.SOLVER_QR, rg.SOLVER_JACOBI):
        for rep in range(3):
            t0 = time.perf_counter()
            res = rt.f_ransac_batched([pts], [idx], thr=1.5, solver=solver, want_mask=True)
            dt = time.perf_counter() - t0
        st = rt.last_stats()
        print(json.dumps({"cfg3_host_call_ms": round(dt * 1e3, 3), "solver": solver,
                          "gevals_s": round(100000 * 16384 / dt * 1e-9, 1), "best": int(res["best_idx"][0]),
                          "count": int(res["best_count"][0]), "stats": st}))
    c64 = None
    t0 = time.perf_counter()
    res64 = rt.f_ransac_batched([pts], [idx], thr=1.5, score_path=rg.SCORE_FP64, want_counts=True)
    dt64 = time.perf_counter() - t0
    res32 = rt.f_ransac_batched([pts], [idx], thr=1.5, want_counts=True)
    print(json.dumps({"cfg3_fp64_ms": round(dt64 * 1e3, 2),
                      "fp32_vs_fp64_count_mismatch": int(np.count_nonzero(res64["counts"][0] != res32["counts"][0]))}))


if __name__ == "__main__":
    print(json.dumps(rt.microbench()))
    check_small(rg.SOLVER_QR)
    check_small(rg.SOLVER_JACOBI)
    timing()
