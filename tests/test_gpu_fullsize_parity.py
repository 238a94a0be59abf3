"""Parity at the FULL sizes of BASELINE.json's configurations (VERDICT r1 "weak" 3): the round-1
tests ran configs[3] at N=1200 and d=100 at N<=300.  Here the region state has the stated size and
whole batches are compared with the oracle (not samples of them), the bootstrapped radius /
enlargement included.  Sizes are picked so that the CPU oracle finishes in seconds per case."""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ultranest_b200 import _native
    return _native.get_engine()


def _live(n, d, seed=1):
    import bench
    return bench.make_live(n, d, seed=seed)


def _draw(region, m, seed, inflate):
    rng = np.random.RandomState(seed)
    d = region.u.shape[1]
    z = rng.normal(size=(m, d))
    z /= ((z**2).sum(axis=1)**0.5).reshape((m, 1))
    uu = z * (region.enlarge * inflate)**0.5 * rng.uniform(size=(m, 1))**(1. / d)
    w = region.ellipsoid_center + np.dot(uu, region.ellipsoid_axes_T)
    return w[np.logical_and(w > 0, w < 1).all(axis=1)]


def test_config3_robust_ellipsoid_n8000_d50_rosenbrock(eng):
    """configs[3]: 50-D, N_live=8000, RobustEllipsoidRegion (Mahalanobis filter alone,
    mlfriends.pyx:1374-1390) + the Rosenbrock batch likelihood (examples/testrosenbrock.py:10-16):
    30-round enlargement, masks of 262 144 proposals and their likelihoods, all against the oracle."""
    from ultranest_b200 import mlfriends as ml
    from ultranest_b200.likelihoods import RosenbrockLogLike
    n, d = 8000, 50
    u = _live(n, d)
    layer = ml.AffineLayer()
    layer.optimize(u, u)
    region = ml.RobustEllipsoidRegion(u, layer)
    r, f = region.compute_enlargement(nbootstraps=30, rng=np.random.RandomState(2))
    # the oracle's version of RobustEllipsoidRegion.compute_enlargement (mlfriends.pyx:1392-1440)
    rng = np.random.RandomState(2)
    want_f = 0.0
    for _ in range(30):
        sel = cport.draw_selection(rng, n)
        if sel.all() or not sel.any():
            continue
        ctr, cov = cport.bounding_ellipsoid(u[sel])
        want_f = max(want_f, cport.enlargement_f(u, sel, ctr, np.linalg.inv(cov)))
    assert r == 1e300 and f == want_f
    region.maxradiussq, region.enlarge = r, f
    region.create_ellipsoid()
    pts = np.vstack([_draw(region, 140000, 3, 1.0), _draw(region, 140000, 4, 1.25)])[:262144]
    want = cport.inside_ellipsoid(pts, region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
    got = region.inside(pts)
    assert (got == want).all() and 0.3 < want.mean() < 0.95
    theta = pts[want][:50000] * 20 - 10
    assert (RosenbrockLogLike()(theta) == cport.loglike_rosenbrock(theta)).all()


def test_config4_mlfriends_n4000_d100(eng):
    """configs[4] at d=100: N_live=4000 MLFriends region; bootstrapped radius/enlargement
    (3 rounds: the CPU oracle needs ~1 s per round here) and the membership masks of 12 000
    proposals (accepted, ellipsoid-rejected and neighbour-rejected) against the oracle."""
    from ultranest_b200 import mlfriends as ml
    n, d = 4000, 100
    u = _live(n, d)
    layer = ml.AffineLayer()
    layer.optimize(u, u)
    region = ml.MLFriends(u, layer)
    # (d > 90: the enlargement's einsum follows NumPy's buffered reduction, chunks of 8192 // d rows)
    got = region.compute_enlargement(nbootstraps=3, rng=np.random.RandomState(2))
    want = cport.compute_enlargement(u, region.unormed, 3, np.random.RandomState(2))
    assert got == want
    region.maxradiussq, region.enlarge = got
    region.create_ellipsoid()
    pts = np.vstack([_draw(region, 7000, 3, 1.0), _draw(region, 7000, 4, 1.3)])[:12000]
    lay = region.transformLayer
    want_mask = cport.region_inside(pts, region.unormed,
                                    lambda p: cport.transform_affine(p, lay.ctr, lay.T),
                                    region.maxradiussq, region.ellipsoid_center,
                                    region.ellipsoid_invcov, region.enlarge)
    got_mask = region.inside(pts)
    assert (got_mask == want_mask).all() and 0.2 < want_mask.mean() < 0.9
    # first-neighbour indices at this size as well
    mask, idx = eng.region_inside(pts[:3000], want_index=True)
    ell = cport.inside_ellipsoid(pts[:3000], region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
    t = cport.transform_affine(pts[:3000], lay.ctr, lay.T)
    assert (idx == np.where(ell, cport.find_nearby(region.unormed, t, region.maxradiussq), -1)).all()


@pytest.mark.parametrize("sure", [1, 0])
def test_config1_n4000_d20_whole_batch_both_filter_levels(eng, sure):
    """configs[1] at full region size: every row of a 2^17-proposal batch (accepting + rejecting
    mix) against the oracle, with the certain-neighbour shortcut of the fp32 filter on and off
    (UNB_OPT_SURE_LEVEL; round 1 had this as a manual environment switch)."""
    from ultranest_b200 import _native
    from ultranest_b200 import mlfriends as ml
    n, d = 4000, 20
    u = _live(n, d)
    layer = ml.AffineLayer()
    layer.optimize(u, u)
    region = ml.MLFriends(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=30, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    pts = np.vstack([_draw(region, 70000, 3, 1.0), _draw(region, 70000, 4, 1.3)])[:1 << 17]
    lay = region.transformLayer
    want = cport.region_inside(pts, region.unormed,
                               lambda p: cport.transform_affine(p, lay.ctr, lay.T),
                               region.maxradiussq, region.ellipsoid_center,
                               region.ellipsoid_invcov, region.enlarge)
    eng.set_option(_native.OPT_SURE_LEVEL, sure)
    try:
        for block in (0, 1):
            eng.set_option(_native.OPT_BLOCK_KERNEL, block)
            assert (region.inside(pts) == want).all(), (sure, block)
            assert (region.inside(pts[:5000]) == want[:5000]).all(), (sure, block)
    finally:
        eng.set_option(_native.OPT_SURE_LEVEL, 1)
        eng.set_option(_native.OPT_BLOCK_KERNEL, 0)
    assert 0.3 < want.mean() < 0.95
