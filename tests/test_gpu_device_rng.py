"""Device-side proposal generation ("throughput mode", unb_sample.cu).  Not the reference's random
stream, so parity is tested the reference's way -- statistical windows of
tests/test_regionsampling.py:35-44,77-85 -- plus what a counter-based generator allows on top:
the device stream equals its host restatement (tests/philox_ref.py, pinned to the published
Philox known answers by tests/test_philox_cpu.py), accepted rows equal the host-side filter of
the raw draws (same membership code as `inside()`), and results do not depend on chunking."""
import numpy as np
import pytest

import philox_ref as pr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ultranest_b200 import _native
    return _native.get_engine()


def _region(upoints, layer_cls="AffineLayer", nboot=30):
    from ultranest_b200 import mlfriends as ml
    layer = getattr(ml, layer_cls)(wrapped_dims=[])
    layer.optimize(upoints, upoints)
    region = ml.MLFriends(upoints, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=nboot)
    region.create_ellipsoid()
    return region


def test_device_stream_equals_host_restatement(eng):
    from ultranest_b200 import _native
    for d in (1, 2, 5, 20, 33):
        got, cube = eng.sample_draw(_native.SAMPLE_UNIT_CUBE, 3000, d, 0x1234567890abcdef, 2**33 + 5)
        want = pr.draw_unit_cube(3000, d, 0x1234567890abcdef, 2**33 + 5)
        assert (got == want).all() and cube.all()        # pure integer -> double arithmetic: exact
    rng = np.random.RandomState(1)
    for d in (2, 5, 20):
        A = rng.normal(size=(d, d)) * 0.05
        ctr = rng.uniform(0.3, 0.7, size=d)
        got, cube = eng.sample_draw(_native.SAMPLE_WRAPPING_ELLIPSOID, 4000, d, 77, 10, ctr, A, 1.7)
        want = pr.draw_wrapping_ellipsoid(4000, d, 77, 10, ctr, A, 1.7)
        assert np.allclose(got, want, rtol=0, atol=1e-12)   # libm vs CUDA log/sincospi/pow: ulps
        assert (cube == np.logical_and(got > 0, got < 1).all(axis=1)).all()


def test_wrapping_ellipsoid_draws_are_uniform_in_the_ellipsoid(eng):
    """(r/R)^d of a uniform draw in a d-ball is uniform; directions are isotropic."""
    from scipy import stats
    from ultranest_b200 import _native
    d, m = 7, 200000
    rng = np.random.RandomState(2)
    A = rng.normal(size=(d, d))
    ctr = np.zeros(d) + 0.5
    w, _ = eng.sample_draw(_native.SAMPLE_WRAPPING_ELLIPSOID, m, d, 5, 0, ctr, A, 2.5)
    z = np.linalg.solve(A.T, (w - ctr).T).T / 2.5**0.5     # back to the unit ball
    r = np.sqrt((z**2).sum(axis=1))
    assert r.max() <= 1 + 1e-12
    assert stats.kstest(r**d, "uniform").pvalue > 1e-3
    assert np.abs(z.mean(axis=0)).max() < 5 * (1.0 / (d + 2) / m)**0.5 * 1.5
    cov = np.cov(z, rowvar=0)
    assert np.allclose(cov, np.eye(d) / (d + 2), atol=4e-3)


@pytest.mark.parametrize("layer", ["ScalingLayer", "AffineLayer"])
def test_statistical_windows_like_the_reference(eng, layer):
    """tests/test_regionsampling.py:10-50 / 53-90 with the draws made on the device."""
    np.random.seed(1)
    if layer == "ScalingLayer":
        upoints = np.random.uniform(0.2, 0.5, size=(1000, 2))
        upoints[:, 1] *= 0.1
        windows = ((0.15, 0.25), (0.015, 0.025), (0.45, 0.55), (0.045, 0.055))
    else:
        upoints = np.random.uniform(size=(1000, 2))
        upoints[:, 1] *= 0.5
        windows = ((0, 0.1), (0, 0.1), (0.95, 1.0), (0.45, 0.55))
    region = _region(upoints, layer)
    region.device_rng = True
    region.device_seed = 11
    for name in ("sample_from_boundingbox", "sample_from_wrapping_ellipsoid"):
        region.current_sampling_method = getattr(region, name)
        newpoints = region.sample(nsamples=4000 if name.endswith("ellipsoid") else 200000)
        assert len(newpoints) > 500
        lo1, lo2 = newpoints.min(axis=0)
        hi1, hi2 = newpoints.max(axis=0)
        assert windows[0][0] <= lo1 < windows[0][1], (name, lo1)
        assert windows[1][0] <= lo2 < windows[1][1], (name, lo2)
        assert windows[2][0] < hi1 <= windows[2][1], (name, hi1)
        assert windows[3][0] <= hi2 < windows[3][1], (name, hi2)
        assert region.inside(newpoints).all()
        assert (newpoints > 0).all() and (newpoints < 1).all()
    # consecutive calls never re-use draws
    a = region.sample_device(1000)
    b = region.sample_device(1000)
    assert len(a) and len(b) and not (a[:5] == b[:5]).all()


def test_accepted_rows_equal_the_host_filter_of_the_raw_draws(eng):
    """Ordered compaction + the same membership pipeline as inside(): deterministic, and
    independent of how the call is chunked (counter-based generator)."""
    import bench
    from ultranest_b200 import _native
    from ultranest_b200 import mlfriends as ml
    from ultranest_b200.likelihoods import GaussianLogLike
    u = 0.5 + (bench.make_live(1500, 6, seed=3) - 0.5) * 7.0    # wide: the unit cube cuts it
    assert (u > 0).all() and (u < 1).all()
    layer = ml.AffineLayer()
    layer.optimize(u, u)
    region = ml.MLFriends(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=10, rng=np.random.RandomState(2))
    region.maxradiussq *= 0.5          # make the neighbour test bite
    region.create_ellipsoid()
    m, seed = 50000, 424242
    for method, const in (("wrapping_ellipsoid", _native.SAMPLE_WRAPPING_ELLIPSOID),
                          ("boundingbox", _native.SAMPLE_UNIT_CUBE)):
        region._device_draws = 0
        rows = region.sample_device(m, method="sample_from_" + method, seed=seed)
        raw, cube = eng.sample_draw(const, m, 6, seed, 0, region.ellipsoid_center,
                                    region.ellipsoid_axes_T, region.enlarge)
        member = region._bind().region_inside(raw, use_ellipsoid=(const == _native.SAMPLE_UNIT_CUBE))
        want = raw[cube & member]
        assert rows.shape == want.shape and (rows == want).all(), method
        assert 0 < len(rows) < m
        # chunking does not change the result
        eng.set_option(_native.OPT_CHUNK_ROWS, 7777)
        try:
            region._device_draws = 0
            again = region.sample_device(m, method="sample_from_" + method, seed=seed)
        finally:
            eng.set_option(_native.OPT_CHUNK_ROWS, 0)
        assert (again == rows).all()
    # fused likelihood and the Lmin cut
    loglike = GaussianLogLike(0.5, 0.05)
    region._device_draws = 0
    rows, like = region.sample_device(m, method="sample_from_wrapping_ellipsoid", seed=seed, loglike=loglike)
    assert (rows == want_rows(region, eng, m, seed)).all()
    assert (like == loglike(rows)).all()
    Lmin = float(np.median(like))
    region._device_draws = 0
    rows2, like2 = region.sample_device(m, method="sample_from_wrapping_ellipsoid", seed=seed,
                                        loglike=loglike, Lmin=Lmin)
    keep = like > Lmin
    assert (rows2 == rows[keep]).all() and (like2 == like[keep]).all()


def want_rows(region, eng, m, seed):
    from ultranest_b200 import _native
    raw, cube = eng.sample_draw(_native.SAMPLE_WRAPPING_ELLIPSOID, m, region.u.shape[1], seed, 0,
                                region.ellipsoid_center, region.ellipsoid_axes_T, region.enlarge)
    return raw[cube & region._bind().region_inside(raw, use_ellipsoid=False)]


def test_device_rng_is_opt_in(eng):
    """Default: every draw stays on the host np.random stream (seeded-run identity)."""
    np.random.seed(4)
    upoints = np.random.uniform(0.3, 0.6, size=(500, 3))
    region = _region(upoints, nboot=5)
    assert region.device_rng is False
    np.random.seed(9)
    a = region.sample_from_wrapping_ellipsoid(nsamples=300)
    state = np.random.get_state()[1][:5].copy()
    np.random.seed(9)
    b = region.sample_from_wrapping_ellipsoid(nsamples=300)
    assert (a == b).all()
    np.random.seed(9)
    region.device_rng = True
    region.current_sampling_method = region.sample_from_wrapping_ellipsoid
    c = region.sample(nsamples=300)
    # the device path did not consume the host stream
    assert not (np.random.get_state()[1][:5] == state).all() or len(c) > 0
    assert region.inside(c).all()
