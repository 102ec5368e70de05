"""CPU tests of the host twin of the device-side counter-based draws (tsbb15_b200/philox.py).  The device side is compared
with this twin bit for bit in tests/test_gpu_round2.py; here the twin itself is pinned: Philox4x32-10 against the
Random123 known-answer vectors, and the sampler's contract (reference draw: np.random.choice(N, 8, replace=False),
fun.py:305-308 — k DISTINCT indices of [0, N), uniform)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def philox(rg):
    return rg.philox


def test_philox4x32_10_known_answers(philox):
    # Random123 kat_vectors: philox4x32 10 rounds
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kats:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(g) for g in got) == out
    # vectorised == scalar
    c0 = np.arange(50, dtype=np.uint64)
    many = philox.philox4x32_10(c0, 7, 9, 11, 123, 456)
    for i in (0, 17, 49):
        one = philox.philox4x32_10(i, 7, 9, 11, 123, 456)
        assert all(int(many[k][i]) == int(one[k]) for k in range(4))


@pytest.mark.parametrize("k", [6, 7, 8])
def test_samples_are_distinct_in_range_and_deterministic(philox, k):
    for n in (k, k + 1, 37, 257, 100000):
        idx = philox.sample_indices(n, 2000, k, seed=5, pair_id=3)
        assert idx.shape == (2000, k) and idx.dtype == np.int32
        assert idx.min() >= 0 and idx.max() < n
        s = np.sort(idx, axis=1)
        assert (s[:, 1:] != s[:, :-1]).all(), "repeated index inside a sample"
        assert np.array_equal(idx, philox.sample_indices(n, 2000, k, seed=5, pair_id=3))
    # n == k: every sample is a permutation of all points
    assert (np.sort(philox.sample_indices(k, 50, k, seed=1), axis=1) == np.arange(k)).all()
    with pytest.raises(ValueError):
        philox.sample_indices(k - 1, 10, k)


def test_samples_depend_on_seed_pair_and_hypothesis_only(philox):
    a = philox.sample_indices(500, 300, 8, seed=9, pair_id=4, hyp_first=100)
    whole = philox.sample_indices(500, 400, 8, seed=9, pair_id=4)
    assert np.array_equal(a, whole[100:400])                   # hypothesis-split mode: a block of the same stream
    assert not np.array_equal(a, philox.sample_indices(500, 300, 8, seed=10, pair_id=4, hyp_first=100))
    assert not np.array_equal(a, philox.sample_indices(500, 300, 8, seed=9, pair_id=5, hyp_first=100))
    batch = philox.sample_indices_batch([500, 300], 64, 8, seed=9, first_pair=4)
    assert np.array_equal(batch[0], whole[:64]) and batch[1].max() < 300


def test_samples_are_uniform(philox):
    """Every index equally likely in every position (chi-square against the uniform law, 49 degrees of freedom: the 99.9 %
    quantile is 85.4), and every unordered pair of positions free of the ordering bias a naive 'skip taken' map has."""
    n, H = 50, 40000
    idx = philox.sample_indices(n, H, 8, seed=2026)
    for j in range(8):
        cnt = np.bincount(idx[:, j], minlength=n)
        chi2 = ((cnt - H / n) ** 2 / (H / n)).sum()
        assert chi2 < 95.0, (j, chi2)
    assert abs(np.mean(idx[:, 7] > idx[:, 0]) - 0.5) < 0.01


def test_synthetic_pairs_have_the_configured_statistics(philox, rg):
    cams, bbox = rg.synth.dino()["Ps"], rg.synth.DINO_BBOX
    pts, (c1, c2) = philox.synth_two_view(20000, 5, cams, bbox)
    assert pts.shape == (20000, 4) and c1 != c2 and 0 <= c1 < 36 and 0 <= c2 < 36
    again, _ = philox.synth_two_view(20000, 5, cams, bbox)
    assert np.array_equal(pts, again)
    other, _ = philox.synth_two_view(20000, 6, cams, bbox)
    assert not np.array_equal(pts, other)
    # the first 30 % of the image-2 points are uniform in the image, the rest follow the epipolar geometry with 0.5 px noise
    out = pts[:6000, 2:]
    assert out.min() >= 0 and out[:, 0].max() <= 640 and out[:, 1].max() <= 480
    assert abs(out[:, 0].mean() - 320) < 8 and abs(out[:, 1].mean() - 240) < 6
    from oracle import geom_path as og
    from oracle import f_path as orc
    F = og.fmatrix_from_cameras(cams[c1], cams[c2])
    d = orc.distance(F, pts[6000:, :2].T.copy(), pts[6000:, 2:].T.copy())
    assert 0.55 < np.sqrt(np.mean(d ** 2)) < 0.95 and np.mean(d < 1.5) > 0.9     # max of two ~N(0, 0.5..0.7) distances
    dout = orc.distance(F, pts[:6000, :2].T.copy(), pts[:6000, 2:].T.copy())
    assert np.mean(dout < 1.5) < 0.03
    # Irwin-Hall(12) noise: unit variance, light tails cut at 6 sigma
    clean, _ = philox.synth_two_view(20000, 5, cams, bbox, sigma_px=0.0, outlier_frac=0.0)
    noisy, _ = philox.synth_two_view(20000, 5, cams, bbox, sigma_px=1.0, outlier_frac=0.0)
    e = (noisy - clean).ravel()
    assert abs(e.std() - 1.0) < 0.01 and abs(e.mean()) < 0.01 and np.abs(e).max() <= 6.0
