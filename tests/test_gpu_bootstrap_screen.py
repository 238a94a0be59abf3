"""The enlargement screen (`_bootstrap_rounds_screened`): device-accumulated moments pick the rounds
that can decide max_r f_r, only those get the reference's NumPy algebra -- and the result must be the
oracle's bit for bit, for every seed, including the cases where the screen declines."""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu


def _region(n, d, seed):
    import bench
    from ultranest_b200 import mlfriends as ml
    u = bench.make_live(n, d, seed=seed)
    layer = ml.AffineLayer()
    layer.optimize(u, u)
    return ml.MLFriends(u, layer)


@pytest.mark.parametrize("n,d,nboot", [(400, 5, 12), (900, 12, 20), (2000, 10, 30), (4000, 20, 30), (1500, 33, 8)])
def test_screened_rebuild_is_the_oracle(n, d, nboot):
    from ultranest_b200 import mlfriends as ml
    before = dict(ml.screen_stats)
    for seed in range(6 if n <= 2000 else 2):
        region = _region(n, d, seed + 1)
        got = region.compute_enlargement(nbootstraps=nboot, rng=np.random.RandomState(seed))
        want = cport.compute_enlargement(region.u, region.unormed, nboot, np.random.RandomState(seed))
        assert got == want, (n, d, seed, got, want)
    used = ml.screen_stats["screened"] - before["screened"]
    exact = ml.screen_stats["exact_rounds"] - before["exact_rounds"]
    rounds = ml.screen_stats["rounds"] - before["rounds"]
    assert used > 0 and exact < rounds / 3, (used, exact, rounds)


def test_moments_match_numpy():
    from ultranest_b200 import _native
    eng = _native.get_engine()
    rng = np.random.RandomState(3)
    u = rng.uniform(0.2, 0.8, size=(777, 9))
    sel = rng.uniform(size=(5, 777)) < 0.6
    c0 = u.mean(axis=0)
    counts, sums, sxx = eng.region_bootstrap_moments(u, sel, c0, 1, 5)
    assert counts[0] == 0 and not sums[0].any()            # outside [round_lo, round_hi)
    for r in range(1, 5):
        y = u[sel[r]] - c0
        assert counts[r] == len(y)
        np.testing.assert_allclose(sums[r], y.sum(axis=0), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(sxx[r], np.triu(y.T @ y), rtol=1e-12, atol=1e-12)


def test_screen_declines_and_exact_path_takes_over():
    """Too few points per round / duplicated coordinates (singular covariance): the screen must step
    aside and the exact path must behave like the reference (value or exception)."""
    from ultranest_b200 import mlfriends as ml
    region = _region(60, 5, 3)          # N < 8 (d + 2): declined by size
    got = region.compute_enlargement(nbootstraps=6, rng=np.random.RandomState(1))
    assert got == cport.compute_enlargement(region.u, region.unormed, 6, np.random.RandomState(1))
    region = _region(800, 6, 4)
    before = ml.screen_stats["declined"]
    u2 = region.u.copy()
    u2[:, 5] = u2[:, 4]                 # exactly collinear -> singular covariance in every round
    layer = ml.ScalingLayer()
    layer.optimize(u2, u2)
    reg2 = ml.MLFriends(u2, layer)
    with pytest.raises((np.linalg.LinAlgError, AssertionError, FloatingPointError)):
        reg2.compute_enlargement(nbootstraps=5, rng=np.random.RandomState(2))
    assert ml.screen_stats["declined"] > before
