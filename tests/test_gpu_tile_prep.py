"""Parity of the tile prep kernel (k_prep_tile, 32 < d): ellipsoid membership through the fused
filter + exact einsum band, layer transform in the defined order, compaction and the fused
likelihood, against the oracle (mlfriends.pyx:882-912, 1186-1211) and against the library's own
exact-only mode (the plain per-row kernels)."""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu

DIMS = [33, 40, 50, 64, 65, 100, 128, 140, 190]   # 190: beyond the tile kernel, plain row kernel


@pytest.fixture(scope="module")
def eng():
    from ultranest_b200 import _native
    return _native.get_engine()


def _ball(rng, n, d):
    z = rng.normal(size=(n, d))
    z /= np.sqrt((z**2).sum(axis=1, keepdims=True))
    return z * rng.uniform(size=(n, 1))**(1.0 / d)


@pytest.mark.parametrize("d", DIMS)
@pytest.mark.parametrize("m", [1, 127, 128, 129, 1500])
def test_ellipsoid_mask_every_edge_radius(eng, d, m):
    """Radii taken from the einsum values themselves put rows exactly on `r <= radius`: those
    rows are inside the filter's band and must be decided in the reference's order."""
    rng = np.random.RandomState(1000 * d + m)
    pts = rng.uniform(size=(m, d))
    ctr = rng.uniform(0.4, 0.6, size=d)
    A = rng.normal(size=(d, d))
    invcov = A @ A.T / d + np.eye(d)
    _, r = cport.inside_ellipsoid(pts, ctr, invcov, 1.0, return_r=True)
    rs = np.sort(r)
    radii = [rs[0], np.nextafter(rs[0], 0), rs[-1], np.nextafter(rs[-1], np.inf), np.median(r)]
    radii += list(rs[rng.randint(m, size=4)])
    for radius in radii:
        want = cport.inside_ellipsoid(pts, ctr, invcov, radius)
        assert (eng.inside_ellipsoid(pts, ctr, invcov, radius) == want).all()
        assert want.sum() == (r <= radius).sum()


def _region(d, n, seed, layer_name):
    from ultranest_b200 import mlfriends as ml
    rng = np.random.RandomState(seed)
    L = np.linalg.cholesky(0.5 * np.ones((d, d)) + 0.5 * np.eye(d))
    u = 0.5 + 0.05 * _ball(rng, n, d) @ L.T
    layer = getattr(ml, layer_name)()
    layer.optimize(u, u)
    region = ml.MLFriends(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(
        nbootstraps=4, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    return region, rng


@pytest.mark.parametrize("d,layer", [(33, "AffineLayer"), (50, "AffineLayer"), (64, "ScalingLayer"),
                                     (100, "AffineLayer"), (128, "AffineLayer")])
def test_region_inside_and_fused_loglike(eng, d, layer):
    from ultranest_b200 import _native
    from ultranest_b200.likelihoods import GaussianLogLike
    region, rng = _region(d, 6 * d, 77 + d, layer)
    lay = region.transformLayer
    if layer == "AffineLayer":
        xf = lambda p: cport.transform_affine(p, lay.ctr, lay.T)   # noqa: E731
    else:
        xf = lambda p: cport.transform_scaling(p, lay.mean, lay.std)   # noqa: E731
    z = _ball(rng, 2000, d) * region.enlarge**0.5 * 1.02   # straddles the ellipsoid surface
    pts_a = region.ellipsoid_center + z @ region.ellipsoid_axes_T
    near = region.u[rng.randint(len(region.u), size=777)] + rng.normal(size=(777, d)) * 2e-3
    loglike = GaussianLogLike(0.5, 0.05)
    for pts in (pts_a, near, region.u.copy(), pts_a[:1], pts_a[:129]):
        want = cport.region_inside(pts, region.unormed, xf, region.maxradiussq,
                                   region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
        got = region.inside(pts)
        assert (got == want).all()
        mask, like = region.inside_and_loglike(pts, loglike)
        assert (mask == want).all()
        assert (like[mask] == cport.loglike_gauss(pts[mask], 0.5 * np.ones(d), 0.05)).all()
        assert np.isneginf(like[~mask]).all()
        # the plain per-row kernels (exact-only mode) agree
        eng.set_option(_native.OPT_EXACT_ONLY, 1)
        try:
            assert (region.inside(pts) == want).all()
        finally:
            eng.set_option(_native.OPT_EXACT_ONLY, 0)
    # first-neighbour indices go through the same prep kernel + the ordered scan
    mask, idx = eng.region_inside(near, want_index=True)
    ell = cport.inside_ellipsoid(near, region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
    want_idx = np.where(ell, cport.find_nearby(region.unormed, xf(near), region.maxradiussq), -1)
    assert (idx == want_idx).all()
    assert 0 < (want_idx >= 0).sum()


def test_ellipsoid_only_region_d50(eng):
    """BASELINE configs[3]: RobustEllipsoidRegion.inside is the Mahalanobis filter alone."""
    from ultranest_b200 import mlfriends as ml
    rng = np.random.RandomState(5)
    d, n = 50, 1200
    u = 0.5 + 0.04 * _ball(rng, n, d)
    layer = ml.AffineLayer()
    layer.optimize(u, u)
    region = ml.RobustEllipsoidRegion(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(
        nbootstraps=4, rng=np.random.RandomState(3))
    region.create_ellipsoid()
    pts = 0.5 + 0.04 * 1.3 * _ball(rng, 20001, d)
    want = cport.inside_ellipsoid(pts, region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
    got = region.inside(pts)
    assert (got == want).all() and 0 < want.sum() < len(want)
