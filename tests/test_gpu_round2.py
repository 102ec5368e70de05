"""Round-2 GPU parity tests (run with `pytest -m gpu` on a B200): the device-side counter-based draws against their host
twin, the seeded / multi-pass / flag-list-overflow variants of the batched call against the plain one, the BASELINE
full-size configs against ORACLE slices (not only against the library's own FP64 path), all 35 noisy Dino pairs, a
property test of guarded-FP32 == FP64, the PnP solvers against each other and the oracle, and the peer-memory exchange.

Every place where a test tolerates a deviation (rank-deficient samples, rounding-noise ties) COUNTS how often the escape
was used, asserts an upper bound, and appends the count to gpurun_out/r02_parity_escapes.jsonl so that a regression cannot
hide in an escape."""
import json
import os

import numpy as np
import pytest

from oracle import f_path as orc
from oracle import pnp_path as opnp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COND_MIN = 1e-6
THR2 = (1.5 / 3217.0) ** 2


def log_escape(test, **kw):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "r02_parity_escapes.jsonl"), "a") as f:
        f.write(json.dumps({"test": test, **{k: (int(v) if isinstance(v, (np.integer, int)) else v) for k, v in kw.items()}}) + "\n")


@pytest.fixture(scope="module")
def rt(rg):
    return rg.runtime


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


# ---------------------------------------------------------------------------------------------------------------
# device-side draws == host twin
# ---------------------------------------------------------------------------------------------------------------
def test_device_philox_equals_host_twin(rg, torch):
    dv, ph = rg.device, rg.philox
    cams, bbox = rg.synth.dino()["Ps"], rg.synth.DINO_BBOX
    d_pts, d_cam = dv.synth_two_view(4, 3001, first_pair=4093, seed_base=1000)
    for p in range(4):
        h, cp = ph.synth_two_view(3001, 4093 + p, cams, bbox)
        assert np.array_equal(d_pts[p].cpu().numpy(), h), "device-generated pair differs from its host replay"
        assert tuple(int(v) for v in d_cam[p].cpu().numpy()) == cp
    for k in (6, 7, 8):
        sizes, hyps = [5000, 300, k, 64], [700, 50, 20, 0]
        got = dv.sample_indices(sizes, hyps, k, seed=2 ** 40 + 99, first_pair=4, hyp_first=3).cpu().numpy()
        exp = np.concatenate([ph.sample_indices(n, h, k, 2 ** 40 + 99, 4 + p, 3) for p, (n, h) in enumerate(zip(sizes, hyps)) if h])
        assert np.array_equal(got, exp)


def test_seeded_call_equals_host_drawn_samples(rt, rg):
    """idx = NULL + seed: the solver draws its own samples; the same samples replayed on the host and passed explicitly must
    give identical counts, winners, F and masks — for both solvers, and through the oracle."""
    pairs = [rg.philox.synth_two_view(n, p, rg.synth.dino()["Ps"], rg.synth.DINO_BBOX)[0] for p, n in enumerate((2500, 777, 64))]
    hyps = [600, 300, 40]
    for solver in (rg.SOLVER_QR, rg.SOLVER_JACOBI):
        a = rt.f_ransac_batched(pairs, None, n_hyp=hyps, sample_seed=7, first_pair=2, want_counts=True, solver=solver)
        idl = [rg.philox.sample_indices(p.shape[0], h, 8, 7, 2 + k) for k, (p, h) in enumerate(zip(pairs, hyps))]
        b = rt.f_ransac_batched(pairs, idl, want_counts=True, solver=solver)
        assert np.array_equal(a["best_idx"], b["best_idx"]) and np.array_equal(a["F"], b["F"])
        for k in range(3):
            assert np.array_equal(a["counts"][k], b["counts"][k]) and np.array_equal(a["mask"][k], b["mask"][k])
    p1, p2 = pairs[1][:, :2].T.copy(), pairs[1][:, 2:].T.copy()
    o = orc.f_ransac(p1, p2, idl[1], 1.5, tie="first")
    good = orc.sample_condition(p1, p2, idl[1]) > COND_MIN
    assert np.array_equal(a["counts"][1][good], o["counts"][good]) and int(a["best_idx"][1]) == o["best"]
    with pytest.raises(ValueError):
        rt.f_ransac_batched(pairs, None)                     # seeded call without n_hyp


def test_multipass_and_flag_list_overflow_do_not_change_results(rt, rg):
    """A large batch is processed in passes (option 6) and a full guard-band flag list falls back to an FP64 recount of the
    hypotheses concerned (option 8): every output must be identical to the single-pass, roomy-list call."""
    sizes = [5000, 3000, 4100, 2048, 900]
    pairs = [rg.synth.two_view(n, seed=60 + k)[0] for k, n in enumerate(sizes)]
    ref = rt.f_ransac_batched(pairs, None, n_hyp=700, sample_seed=3, want_mask=True, want_counts=True)
    one = rt.f_ransac_batched(pairs[:1], None, n_hyp=700, sample_seed=3, want_counts=True)
    try:
        rt.set_option(6, 1)                                  # every pair becomes its own pass (a pass holds >= 1 pair)
        mp = rt.f_ransac_batched(pairs, None, n_hyp=700, sample_seed=3, want_mask=True, want_counts=True)
        assert rt.last_stats()["passes"] == 5
        assert all(np.array_equal(a, b) for a, b in zip(mp["counts"], ref["counts"]))
    finally:
        rt.set_option(6, 0)
    try:
        rt.set_option(8, 100)                                # a list of 100 records: nearly every hypothesis overflows
        ov = rt.f_ransac_batched(pairs, None, n_hyp=700, sample_seed=3, want_mask=True)
        st = rt.last_stats()
        ov1 = rt.f_ransac_batched(pairs[:1], None, n_hyp=700, sample_seed=3, want_counts=True)
    finally:
        rt.set_option(8, 0)
    assert st["overflow"] > 100
    for out in (mp, ov):
        assert np.array_equal(out["best_idx"], ref["best_idx"]) and np.array_equal(out["best_count"], ref["best_count"])
        assert np.array_equal(out["F"], ref["F"])
        assert all(np.array_equal(a, b) for a, b in zip(out["mask"], ref["mask"]))
    assert np.array_equal(ov1["counts"][0], one["counts"][0])
    with pytest.raises(ValueError):
        rt.set_option(8, -1)


def test_reuse_points_flag_and_hypothesis_base(rg, torch):
    """RG_FLAG_REUSE_POINTS skips the prepare kernels (same points, same threshold) and hyp_index_base makes a block of a
    pair's hypotheses reproduce exactly the corresponding part of the whole call (sampling included)."""
    dv = rg.device
    d, _ = dv.synth_two_view(1, 20000, first_pair=1)
    po, H = dv.offsets([20000]), 3000
    whole = dv.FOutputs(1, 20000, want_mask=True, want_key=True)
    dv.f_ransac(d, po, None, dv.offsets([H]), whole, seed=11, first_pair=1)
    cnt_all = rg.runtime.f_ransac_batched([d[0].cpu().numpy()], None, n_hyp=H, sample_seed=11, first_pair=1,
                                          want_counts=True)["counts"][0]
    best = (-1, -1)
    for lo in (0, 1000, 2000):
        blk = dv.FOutputs(1, 20000, want_mask=False, want_key=True)
        dv.f_ransac(d, po, None, dv.offsets([1000]), blk, seed=11, first_pair=1, hyp_first=lo,
                    flags=rg.FLAG_REUSE_POINTS if lo else 0)
        key = int(blk.key.item())
        cnt, gi = rg.parallel.key_decode(key)
        assert gi == lo + int(blk.best_idx.item()) and cnt == int(blk.best_count.item()) == cnt_all[lo:lo + 1000].max()
        assert gi == lo + int(np.argmax(cnt_all[lo:lo + 1000]))
        best = max(best, (key, gi))
    assert best[1] == int(whole.best_idx.item()) and best[0] == int(whole.key.item())
    # the flag is refused when the previous pass prepared other points / another threshold
    other, _ = dv.synth_two_view(1, 20000, first_pair=2)
    with pytest.raises(ValueError):
        dv.f_ransac(other, po, None, dv.offsets([1000]), dv.FOutputs(1), seed=11, flags=rg.FLAG_REUSE_POINTS)
    with pytest.raises(ValueError):
        dv.f_ransac(d, po, None, dv.offsets([1000]), dv.FOutputs(1), seed=11, thr=2.0, flags=rg.FLAG_REUSE_POINTS)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE full-size configs against ORACLE slices
# ---------------------------------------------------------------------------------------------------------------
def test_config3_full_shape_against_an_oracle_slice(rt, rg):
    """Config 3 (100 000 x 16 384): the winner and the first 64 hypotheses against the numpy oracle (reference formula,
    LAPACK SVD), at two thresholds and under the Sampson criterion."""
    N, H = 100000, 16384
    pts, _ = rg.synth.two_view(N, seed=1)
    idx = rg.philox.sample_indices(N, H, 8, seed=2)
    p1, p2 = pts[:, :2].T.copy(), pts[:, 2:].T.copy()
    a = rt.f_ransac_batched([pts], None, n_hyp=H, sample_seed=2, want_counts=True, want_flags=True)
    sel = np.unique(np.concatenate([np.arange(64), [int(a["best_idx"][0])]]))
    F = orc.solve_hypotheses(p1, p2, idx[sel])
    good = orc.sample_condition(p1, p2, idx[sel]) > COND_MIN
    o = orc.score_hypotheses(F, p1, p2, 1.5)
    assert np.array_equal(a["counts"][0][sel][good], o[good])
    log_escape("config3_oracle_slice", checked=int(good.sum()), rank_deficient_skipped=int((~good).sum()))
    assert (~good).sum() <= 1
    assert int(a["best_count"][0]) == int(o[list(sel).index(int(a["best_idx"][0]))]) == int(a["mask"][0].sum())
    for thr, mode in ((0.25, 0), (4.0, 0), (1.5, 1)):
        b = rt.f_ransac_batched([pts], [idx[:64]], thr=thr, mode=mode, want_counts=True)
        ob = orc.score_hypotheses(F[:64], p1, p2, thr, mode)
        g = good[:64]
        assert np.array_equal(b["counts"][0][g], ob[g]), (thr, mode)


def test_config4_full_shape_against_an_oracle_slice(rt, rg):
    """Config 4 (PnP, 1 000 000 x 8 192): consensus counts of the first 48 hypotheses and of the winner against the oracle
    (pnp.py:132-152 solve by LAPACK SVD + ransac.py:96-105 scoring), both solvers."""
    N, H = 1000000, 8192
    X, y, _ = rg.synth.pnp_scene(N, seed=4)
    pidx = rg.sampling.fast(N, H, 6, seed=2)
    yh = np.hstack([y, np.ones((N, 1))])
    res = {}
    try:
        for solver in (0, 1):
            rt.set_option(7, solver)
            res[solver] = rt.pnp_ransac(X, y, pidx, THR2, want_counts=True, want_flags=True)
    finally:
        rt.set_option(7, 0)
    a, b = res[0], res[1]
    assert a["best_idx"] == b["best_idx"] and a["best_count"] == b["best_count"] and np.array_equal(a["mask"], b["mask"])
    both = (a["flags"] == 0) & (b["flags"] == 0)
    n_diff = int((a["counts"][both] != b["counts"][both]).sum())
    log_escape("config4_two_solvers", hypotheses=H, flagged=int((~both).sum()), count_differences=n_diff)
    assert n_diff == 0 and (~both).sum() <= 8
    sel = np.unique(np.concatenate([np.arange(48), [a["best_idx"]]]))
    o = opnp.pnp_ransac(X, yh, pidx[sel], THR2)
    good = (opnp.sample_gap(X, yh, pidx[sel]) > 1e-7) & (a["flags"][sel] == 0)
    assert np.array_equal(a["counts"][sel][good], o["counts"][good])
    log_escape("config4_oracle_slice", checked=int(good.sum()), skipped=int((~good).sum()))
    assert (~good).sum() <= 2
    k = list(sel).index(a["best_idx"])
    assert a["best_count"] == int(o["counts"][k]) == int(a["mask"].sum())


def test_all_35_noisy_dino_pairs_against_the_oracle(rt, rg):
    """Config 2 as stated: ALL 35 consecutive pairs of the noisy tracks (N = 207..445), 1 200 hypotheses each, one call:
    per-hypothesis counts, winner and inlier set of every pair against the oracle."""
    pairs = [rg.synth.dino_noisy_pair(i, i + 1) for i in range(35)]
    pts = [np.ascontiguousarray(np.hstack([a, b])) for a, b in pairs]
    H = 1200
    res = rt.f_ransac_batched(pts, None, n_hyp=H, sample_seed=35, want_counts=True, want_flags=True)
    n_bad_cond = n_mismatch = n_checked = 0
    for k in range(35):
        idx = rg.philox.sample_indices(pts[k].shape[0], H, 8, 35, k)
        p1, p2 = pts[k][:, :2].T.copy(), pts[k][:, 2:].T.copy()
        o = orc.f_ransac(p1, p2, idx, 1.5, tie="first")
        cond = orc.sample_condition(p1, p2, idx)
        bad = res["counts"][k] != o["counts"]
        assert not np.any(bad & (cond > COND_MIN)), k
        n_bad_cond += int((cond <= COND_MIN).sum())
        n_mismatch += int(bad.sum())
        n_checked += H
        assert int(res["best_idx"][k]) == o["best"], k
        assert np.array_equal(res["mask"][k], o["mask"]), k
    log_escape("dino_35_noisy_pairs", hypotheses=n_checked, rank_deficient=n_bad_cond, count_mismatches=n_mismatch)
    assert n_bad_cond <= 0.002 * n_checked and n_mismatch <= n_bad_cond


def test_reference_golden_escapes_are_counted(rt, rg, f_golden, noisy01):
    """The reference's own seeded run (config 1): how many of the 3000 golden hypotheses are rank deficient (the only ones
    allowed to differ) and how many actually differ — both bounded."""
    p1, p2 = noisy01
    idx = f_golden["ransac_idx"]
    pts = np.ascontiguousarray(np.concatenate([p1.T, p2.T], axis=1))
    res = rt.f_ransac_batched([pts], [idx], thr=1.5, tie_mode=rg.TIE_REFERENCE, want_counts=True)
    cond = orc.sample_condition(p1, p2, idx)
    bad = res["counts"][0] != f_golden["ransac_counts"]
    log_escape("reference_golden_config1", hypotheses=len(idx), rank_deficient=int((cond <= COND_MIN).sum()),
               count_mismatches=int(bad.sum()))
    assert (cond <= COND_MIN).sum() <= 6 and bad.sum() <= (cond <= COND_MIN).sum()
    assert not np.any(bad & (cond > COND_MIN))


def test_tie_rule_escape_is_counted(rt, rg):
    """Exact data: every hypothesis ties at N inliers and the reference's rule compares rounding noise.  Over several
    pairs, count how often the library's pick differs from the oracle's replay on the same F (allowed only when the
    decision statistics agree to 1e-6) and bound it."""
    differ = total = 0
    for i in (3, 10, 20, 30):
        y1, y2 = rg.synth.dino_clean_pair(i, i + 1)
        p1, p2 = y1.T.copy(), y2.T.copy()
        idx = rg.sampling.fast(p1.shape[1], 200, 8, seed=2 + i)
        pts = np.ascontiguousarray(np.concatenate([p1.T, p2.T], axis=1))
        res = rt.f_ransac_batched([pts], [idx], thr=1.5, tie_mode=rg.TIE_REFERENCE, want_counts=True, want_F_all=True)
        counts = res["counts"][0]
        expect = orc.select_reference_rule(counts, res["F_all"][0], p1, p2)
        got = int(res["best_idx"][0])
        total += 1
        if got != expect:
            differ += 1
            d_got = orc.distance(res["F_all"][0][got], p1, p2)
            d_exp = orc.distance(res["F_all"][0][expect], p1, p2)
            assert np.linalg.norm(d_got) == pytest.approx(np.linalg.norm(d_exp), rel=1e-6)
        assert counts[got] == counts.max()
    log_escape("tie_rule_replay", pairs=total, picks_differing_from_oracle=differ)
    assert differ <= 1


# ---------------------------------------------------------------------------------------------------------------
# property test: guarded FP32 == FP64 over random frames and thresholds
# ---------------------------------------------------------------------------------------------------------------
def test_guarded_fp32_equals_fp64_property(rt, rg):
    from hypothesis import HealthCheck, given, settings, strategies as st_

    base, _ = rg.synth.two_view(1800, seed=5)
    idx = rg.sampling.fast(1800, 96, 8, seed=6)

    @settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(log_thr=st_.floats(-3.0, 3.0), log_scale=st_.floats(-4.0, 4.0), ox=st_.floats(-1e4, 1e4), oy=st_.floats(-1e4, 1e4),
           sampson=st_.booleans(), n=st_.integers(8, 1800))
    def prop(log_thr, log_scale, ox, oy, sampson, n):
        s = 10.0 ** log_scale
        p = base[:n] * s + np.array([ox, oy, -oy, ox]) * s          # arbitrary similarity of both images
        thr = (10.0 ** log_thr) * s
        ii = idx[(idx < n).all(axis=1)]
        if len(ii) == 0:
            ii = np.arange(8, dtype=np.int32)[None]
        mode = rg.MODE_SAMPSON if sampson else rg.MODE_EPI_MAX
        a = rt.f_ransac_batched([p], [ii], thr=thr, mode=mode, want_counts=True)
        b = rt.f_ransac_batched([p], [ii], thr=thr, mode=mode, want_counts=True, score_path=rg.SCORE_FP64)
        assert np.array_equal(a["counts"][0], b["counts"][0])
        assert int(a["best_idx"][0]) == int(b["best_idx"][0]) and np.array_equal(a["mask"][0], b["mask"][0])

    prop()


# ---------------------------------------------------------------------------------------------------------------
# PnP: the thread-per-hypothesis solver
# ---------------------------------------------------------------------------------------------------------------
def test_pnp_row_solver_on_exact_and_noisy_data(rt, rg):
    """Exact Dino data: sigma_12 = 0, the smallest row of R is pure rounding noise and the minimiser comes from the
    orthogonal complement — it must still be the ground-truth camera (every sample, every view).  Noisy data: both solvers
    and the oracle agree on every unflagged hypothesis."""
    try:
        for solver in (0, 1):
            rt.set_option(7, solver)
            for v in (0, 17, 35):
                X, y, K = rg.synth.dino_view_2d3d(v)
                idx = rg.sampling.fast(X.shape[0], 64, 6, seed=v)
                r = rt.pnp_ransac(X, y, idx, THR2, want_counts=True, want_poses=True, want_flags=True)
                ok = r["flags"] == 0
                assert ok.sum() >= 60 and (r["counts"][ok] == X.shape[0]).all(), (solver, v)
        for n in (6, 7, 8):
            X, y, _ = rg.synth.pnp_scene(3000, seed=n, sigma_px=0.3, outlier_frac=0.25)
            idx = rg.sampling.fast(3000, 400, n, seed=n)
            yh = np.hstack([y, np.ones((3000, 1))])
            o = opnp.pnp_ransac(X, yh, idx, THR2)
            for solver in (0, 1):
                rt.set_option(7, solver)
                r = rt.pnp_ransac(X, y, idx, THR2, want_counts=True, want_flags=True)
                good = (opnp.sample_gap(X, yh, idx) > 1e-7) & (r["flags"] == 0)
                assert good.sum() > 380 and np.array_equal(r["counts"][good], o["counts"][good]), (n, solver)
                assert r["best_idx"] == o["best"] and np.array_equal(r["mask"], o["mask"])
    finally:
        rt.set_option(7, 0)
    with pytest.raises(ValueError):
        rt.set_option(7, 2)


# ---------------------------------------------------------------------------------------------------------------
# multi-GPU building blocks on one GPU
# ---------------------------------------------------------------------------------------------------------------
def test_peer_memory_exchange_world_size_1(rg, torch):
    """The exchange kernel with a single rank (its own buffer is its only peer): keys in, winner out, repeated calls
    (sequence numbers, double buffering), inside and outside a CUDA graph."""
    dv = rg.device
    ex = dv.P2PExchange()
    P = 5
    key = torch.tensor([rg.parallel.argmax_key(10 + p, 100 * p) for p in range(P)], dtype=torch.int64, device="cuda")
    key[3] = 0                                                    # "no hypothesis with an inlier"
    pay = torch.arange(P * 9, dtype=torch.float64, device="cuda").reshape(P, 9)
    bi = torch.empty(P, dtype=torch.int32, device="cuda"); bc = torch.empty_like(bi); po = torch.empty_like(pay)
    for _ in range(5):
        ex.argmax(key, pay, bi, bc, po)
    ex.check()
    assert bi.tolist() == [0, 100, 200, -1, 400] and bc.tolist() == [10, 11, 12, 0, 14]
    assert torch.equal(po[[0, 1, 2, 4]], pay[[0, 1, 2, 4]]) and torch.isnan(po[3]).all()
    ex.close()


def test_sharded_and_split_classes_world_size_1(rg, torch):
    par, dv = rg.parallel, rg.device
    sh = par.PairShardedRansac(5, 3000, 512, want_mask=True)
    sh.generate(seed_base=1000)
    sh.run(thr=1.5, sample_seed=9)
    got = sh.unpack(sh.gather())
    pairs = [rg.philox.synth_two_view(3000, p, rg.synth.dino()["Ps"], rg.synth.DINO_BBOX)[0] for p in range(5)]
    ref = rg.runtime.f_ransac_batched(pairs, None, n_hyp=512, sample_seed=9)
    assert np.array_equal(got["best_idx"], ref["best_idx"]) and np.array_equal(got["F"], ref["F"])
    assert all(np.array_equal(a, b) for a, b in zip(got["mask"], ref["mask"]))
    sh.alloc_host()
    sh.h_pts[: sh.P].copy_(sh.d_pts)
    blk = sh.run_host(thr=1.5, sample_seed=9).numpy()
    assert np.array_equal(blk, sh.gather()["block"].cpu().numpy())
    d = torch.from_numpy(pairs[0]).cuda()
    sp = par.SplitHypothesesF(d, 512, sample_seed=9)
    sp.run(thr=1.5); sp.run(thr=1.5)
    r = sp.result()
    assert r["best_idx"] == int(ref["best_idx"][0]) and np.array_equal(r["mask"], ref["mask"][0]) and np.array_equal(r["F"], ref["F"][0])
    sp.capture(thr=1.5)
    sp.replay(); sp.replay()
    r2 = sp.result()
    assert r2["best_idx"] == r["best_idx"] and np.array_equal(r2["mask"], r["mask"]) and np.array_equal(r2["F"], r["F"])


def test_pipelined_passes_equal_serial_passes(rt, rg, torch):
    """Calls of several passes run the fix-up / selection / mask kernels of pass k on a second stream while pass k+1 is solved
    and scored (option 10, two sets of per-pass workspaces).  Every output must be identical to the strictly serial order and
    to the single-pass call, through the host entry point and the device entry point, for both tie rules, with ragged
    pairs, an odd number of passes and calls repeated back to back (the sets are reused across calls)."""
    from tsbb15_b200 import device as dv
    sizes = [5000, 3000, 4100, 2048, 900, 7000, 1234]
    hyps = [700, 300, 513, 1100, 64, 900, 257]
    pairs = [rg.synth.two_view(n, seed=80 + k)[0] for k, n in enumerate(sizes)]
    for tie in (rg.TIE_FIRST, rg.TIE_REFERENCE):
        ref = rt.f_ransac_batched(pairs, None, n_hyp=hyps, sample_seed=5, want_mask=True, tie_mode=tie)
        assert rt.last_stats()["passes"] == 1
        outs = []
        try:
            rt.set_option(6, 1)                              # every pair its own pass
            for piped in (1, 0, 1, 1):
                rt.set_option(10, piped)
                outs.append(rt.f_ransac_batched(pairs, None, n_hyp=hyps, sample_seed=5, want_mask=True, tie_mode=tie))
                assert rt.last_stats()["passes"] == len(pairs)
        finally:
            rt.set_option(6, 0)
            rt.set_option(10, 1)
        for o in outs:
            assert np.array_equal(o["best_idx"], ref["best_idx"]) and np.array_equal(o["best_count"], ref["best_count"])
            assert np.array_equal(o["F"], ref["F"])
            assert all(np.array_equal(a, b) for a, b in zip(o["mask"], ref["mask"]))
    # device-resident entry point: one call, its passes pipelined internally; guard-band statistics must also agree
    dev = torch.device("cuda", 0)
    d_pts = torch.from_numpy(np.concatenate(pairs)).to(dev)
    po, ho = dv.offsets(sizes), dv.offsets(hyps)
    res = {}
    try:
        for label, opt6, opt10 in (("single", 0, 1), ("piped", 1, 1), ("serial", 1, 0), ("piped2", 1, 1)):
            rt.set_option(6, opt6)
            rt.set_option(10, opt10)
            o = dv.FOutputs(len(sizes), int(po[-1]), want_mask=True)
            dv.f_ransac(d_pts, po, None, ho, o, seed=9, tie_mode=rg.TIE_REFERENCE)
            st = rt.last_stats()
            res[label] = (o.best_idx.cpu().numpy(), o.best_count.cpu().numpy(), o.F.cpu().numpy(), o.mask.cpu().numpy(),
                          st["recheck_groups"], st["band_evals"], st["flips"])
            assert st["passes"] == (len(sizes) if opt6 else 1)
    finally:
        rt.set_option(6, 0)
        rt.set_option(10, 1)
    for label in ("piped", "serial", "piped2"):
        for a, b in zip(res[label][:4], res["single"][:4]):
            assert np.array_equal(a, b), label
        assert res[label][4:] == res["single"][4:], (label, res[label][4:], res["single"][4:])


def test_two_devices_in_one_process(rg, torch):
    """One process driving two GPUs (the scorer's dynamic-shared-memory attribute and occupancy are per
    device): the same batch on both devices gives identical results."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    pairs = [rg.synth.two_view(4000, seed=k)[0] for k in range(3)]
    a = rg.runtime.f_ransac_batched(pairs, None, n_hyp=600, sample_seed=1, device=0, want_counts=True)
    b = rg.runtime.f_ransac_batched(pairs, None, n_hyp=600, sample_seed=1, device=1, want_counts=True)
    assert np.array_equal(a["best_idx"], b["best_idx"]) and np.array_equal(a["F"], b["F"])
    assert all(np.array_equal(x, y) for x, y in zip(a["counts"], b["counts"]))
    X, y, _ = rg.synth.pnp_scene(5000, seed=1)
    idx = rg.sampling.fast(5000, 256, 6, seed=1)
    pa = rg.runtime.pnp_ransac(X, y, idx, THR2, device=0)
    pb = rg.runtime.pnp_ransac(X, y, idx, THR2, device=1)
    assert pa["best_idx"] == pb["best_idx"] and np.array_equal(pa["mask"], pb["mask"])
