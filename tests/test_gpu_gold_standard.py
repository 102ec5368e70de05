"""GPU tests of the device-resident gold-standard refinement against
  * the reference's own run of that stage (tests/golden/gs_golden.npz: SciPy least_squares on lab3.fmatrix_residuals_gs,
    5482 residual evaluations, stopped on ftol at cost 8.288), and
  * a dense numpy version of the same Levenberg-Marquardt / Schur iteration (oracle/gs_path.py).

Contract.  The COST is the reference's (residual vector equal to lab3.fmatrix_residuals_gs to 1e-12).  The MINIMISER is
not SciPy's trust-region solver, so F_gold is compared the way the existing getFFromLabCode test compares the LM stage:
within 2e-3 of the reference's normalised F_gold (the reference moves its own F_gold by 2e-4 when its input changes by
1e-13), and additionally the final cost must not exceed the reference's."""
import numpy as np
import pytest

from oracle import f_path as orc
from oracle import geom_path as og
from oracle import gs_path as ogs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt(rg):
    return rg.runtime


def _nerr(Fa, Fb):
    Fb = orc.normalise_F(Fb)
    return np.linalg.norm(orc.normalise_F(Fa, Fb) - Fb)


def test_residuals_gs_match_reference_golden(rt, gs_golden):
    g = gs_golden
    r = rt.fmatrix_residuals_gs(g["params0"], g["in1"], g["in2"])
    assert r.shape == g["resid0"].shape and np.abs(r - g["resid0"]).max() < 1e-12 * max(1.0, np.abs(g["resid0"]).max())
    r2 = rt.fmatrix_residuals_gs(g["scipy_params"], g["in1"], g["in2"])
    assert abs(0.5 * r2 @ r2 - float(g["scipy_cost"])) < 1e-9 * float(g["scipy_cost"])
    rng = np.random.default_rng(0)
    par = g["params0"] * (1 + 1e-3 * rng.normal(size=g["params0"].shape))
    assert np.allclose(rt.fmatrix_residuals_gs(par, g["in1"], g["in2"]), ogs.fmatrix_residuals_gs(par, g["in1"], g["in2"]),
                       rtol=1e-12, atol=1e-12)


def test_gold_standard_reaches_the_minimum_of_the_reference_cost(rt, rg, gs_golden):
    g = gs_golden
    pts = np.ascontiguousarray(np.hstack([g["in1"].T, g["in2"].T]))
    res = rt.gold_standard([pts], g["F0"][None], want_points=True)
    assert res["status"][0] == 2 and 3 <= res["iters"][0] <= 50
    # not above the reference's stopping point, and equal to the dense numpy run of the same iteration
    assert res["cost"][0] <= float(g["scipy_cost"])
    Fo, co, ito, C1o, Xo = ogs.gold_standard_lm(g["F0"], g["in1"], g["in2"])
    assert abs(res["cost"][0] - co) < 1e-9 * co
    assert _nerr(res["F"][0], Fo) < 1e-7
    assert _nerr(res["F"][0], g["F_gold"]) < 2e-3
    rms = lambda F: np.sqrt(np.mean(orc.fmatrix_residuals(F, g["in1"], g["in2"]) ** 2))
    assert rms(res["F"][0]) <= rms(g["F_gold"]) * (1 + 1e-6) < rms(g["F0"])
    # drop-in path
    F2, info = rg.fun.gold_standard_device(g["F0"], g["in1"], g["in2"], np.arange(g["in1"].shape[1]), full_output=True)
    assert _nerr(F2, res["F"][0]) < 1e-9 and abs(info["cost"] - res["cost"][0]) < 1e-9 * res["cost"][0]


def test_gold_standard_batched_with_masks_and_edge_cases(rt, rg, dino):
    """Several pairs in one call, inlier masks from F-RANSAC, one empty pair, one pair whose F is NaN."""
    rng = np.random.default_rng(4)
    pairs = []
    for i in (0, 9, 21):
        y1, y2 = dino["x2d"][i].T, dino["x2d"][i + 1].T
        ok = np.logical_and(np.any(y1 != -1, axis=1), np.any(y2 != -1, axis=1))
        pts = np.hstack([y1[ok], y2[ok]]) + rng.normal(0, 0.3, (ok.sum(), 4))
        pts[::6, 2:] = rng.uniform(0, 600, (len(pts[::6]), 2))
        pairs.append(pts)
    fr = rg.batched.f_ransac_pairs(pairs, n_hyp=3000, thr=1.5, seed=2, host_sampling=True)
    batch = pairs + [np.zeros((0, 4)), pairs[0]]
    F0 = np.concatenate([fr["F"], fr["F"][:1], np.full((1, 3, 3), np.nan)])
    masks = list(fr["mask"]) + [np.zeros(0, np.uint8), fr["mask"][0]]
    res = rt.gold_standard(batch, F0, masks=masks, want_points=True)
    assert res["status"].tolist()[-1] == 1 and np.isnan(res["F"][-1]).all()
    for p in range(3):
        m = fr["mask"][p].astype(bool)
        one = rt.gold_standard([pairs[p][m]], fr["F"][p][None])
        assert res["status"][p] in (2, 3) and abs(res["cost"][p] - one["cost"][0]) < 1e-8 * one["cost"][0]
        assert _nerr(res["F"][p], one["F"][0]) < 1e-7
        in1, in2 = pairs[p][m, :2].T.copy(), pairs[p][m, 2:].T.copy()
        c0 = ogs.cost(*ogs.start_point(fr["F"][p], in1, in2)[::2], in1, in2)
        assert res["cost"][p] < c0                                       # lower than the starting cost
        assert np.isnan(res["X"][p][~m]).all() and np.isfinite(res["X"][p][m]).all()
        rms0 = np.sqrt(np.mean(orc.fmatrix_residuals(fr["F"][p], in1, in2) ** 2))
        rms1 = np.sqrt(np.mean(orc.fmatrix_residuals(res["F"][p], in1, in2) ** 2))
        assert rms1 <= rms0 * (1 + 1e-9)
    assert res["X"][3].shape == (0, 3)
    assert rt.gold_standard([], np.zeros((0, 3, 3)))["F"].shape == (0, 3, 3)


def test_gold_standard_on_exact_data_keeps_the_exact_f(rt, dino):
    Ps = dino["Ps"]
    y1, y2 = dino["x2d"][3].T, dino["x2d"][4].T
    ok = np.logical_and(np.any(y1 != -1, axis=1), np.any(y2 != -1, axis=1))
    pts = np.ascontiguousarray(np.hstack([y1[ok], y2[ok]]))
    Ft = og.fmatrix_from_cameras(Ps[3], Ps[4])
    res = rt.gold_standard([pts], Ft[None])
    assert res["cost"][0] < 1e-12 and _nerr(res["F"][0], Ft) < 1e-9


def test_fused_single_cta_loop_equals_multi_kernel_path(rt, rg, gs_golden, dino):
    """Pairs of up to 4096 correspondences run the whole LM loop in one CTA (gs_fused); option 5 forces the multi-kernel
    path.  Same iteration, same decisions: status equal, cost and F to rounding."""
    g = gs_golden
    rng = np.random.default_rng(9)
    batch = [np.ascontiguousarray(np.hstack([g["in1"].T, g["in2"].T]))]
    F0 = [g["F0"]]
    for i in (4, 17):
        y1, y2 = dino["x2d"][i].T, dino["x2d"][i + 1].T
        ok = np.logical_and(np.any(y1 != -1, axis=1), np.any(y2 != -1, axis=1))
        batch.append(np.hstack([y1[ok], y2[ok]]) + rng.normal(0, 0.4, (ok.sum(), 4)))
        F0.append(og.fmatrix_from_cameras(dino["Ps"][i], dino["Ps"][i + 1]))
    masks = [np.ones(len(b), np.uint8) for b in batch]
    masks[1][::5] = 0
    a = rt.gold_standard(batch, np.stack(F0), masks=masks, want_points=True)

    def agrees(b):
        # the multi-kernel path sums with floating-point atomics, so its last convergence test (relative decrease against
        # ftol = 1e-12) can fall on the other side: at most one iteration apart
        ok = np.abs(a["iters"] - b["iters"]).max() <= 1 and a["status"].tolist() == b["status"].tolist()
        for p in range(3):
            ok = ok and abs(a["cost"][p] - b["cost"][p]) < 1e-9 * b["cost"][p] and _nerr(a["F"][p], b["F"][p]) < 1e-9
            m = masks[p].astype(bool)
            # the points live in a projective frame that the free gauge lets drift: compared loosely
            ok = ok and np.abs(a["X"][p][m] - b["X"][p][m]).max() < 1e-5 * np.abs(b["X"][p][m]).max()
            ok = ok and bool(np.isnan(a["X"][p][~m]).all())
        return bool(ok)

    # the atomics make the multi-kernel path's rounding (hence, rarely, its stopping iteration) vary from run to run: the
    # fused path has to agree with ONE of up to three runs of it (observed: 1 disagreement in ~15 full-suite runs)
    try:
        rt.set_option(5, 1)
        good = False
        for _ in range(3):
            good = agrees(rt.gold_standard(batch, np.stack(F0), masks=masks, want_points=True))
            if good:
                break
    finally:
        rt.set_option(5, 0)
    assert good
    # reproducible bit for bit (no atomics on this path)
    a2 = rt.gold_standard(batch, np.stack(F0), masks=masks, want_points=True)
    assert np.array_equal(a2["cost"], a["cost"]) and np.array_equal(a2["F"], a["F"])
