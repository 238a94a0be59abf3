"""Clustered live tiles + proposals binned by nearest tile centroid (unb_cluster.cu): scheduling
only -- masks must equal the oracle's and the unbinned path's bit for bit, for every layout the
clustering can produce (few tiles, more than 64 tiles, ragged last tile, rows patched in place
after the clustering), and the point of it (fewer tile visits) must show."""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu


@pytest.fixture()
def eng():
    from ultranest_b200 import _native
    e = _native.get_engine()
    yield e
    e.set_option(_native.OPT_BIN_MIN_ROWS, 0)       # the default: off
    e.set_option(_native.OPT_SURE_LEVEL, 1)


def _region(n, d, seed=1, nboot=8):
    import bench
    from ultranest_b200 import mlfriends as ml
    u = bench.make_live(n, d, seed=seed)
    layer = ml.AffineLayer()
    layer.optimize(u, u)
    region = ml.MLFriends(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=nboot, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    return region


def _draw(region, m, seed, inflate):
    rng = np.random.RandomState(seed)
    d = region.u.shape[1]
    z = rng.normal(size=(m, d))
    z /= ((z**2).sum(axis=1)**0.5).reshape((m, 1))
    uu = z * (region.enlarge * inflate)**0.5 * rng.uniform(size=(m, 1))**(1. / d)
    w = region.ellipsoid_center + np.dot(uu, region.ellipsoid_axes_T)
    return w[np.logical_and(w > 0, w < 1).all(axis=1)]


def _oracle(region, pts):
    lay = region.transformLayer
    return cport.region_inside(pts, region.unormed, lambda p: cport.transform_affine(p, lay.ctr, lay.T),
                               region.maxradiussq, region.ellipsoid_center, region.ellipsoid_invcov,
                               region.enlarge)


@pytest.mark.parametrize("n,d", [(700, 20), (4000, 20), (300, 5), (257, 3), (5000, 8), (130, 4), (1000, 32)])
def test_binned_masks_equal_oracle_and_unbinned(eng, n, d):
    from ultranest_b200 import _native
    region = _region(n, d)
    pts = np.vstack([_draw(region, 9000, 3, 1.0), _draw(region, 6000, 4, 1.35)])
    want = _oracle(region, pts)
    assert 0.2 < want.mean() < 0.99
    for sure in (1, 0):
        eng.set_option(_native.OPT_SURE_LEVEL, sure)
        eng.set_option(_native.OPT_BIN_MIN_ROWS, 1)          # every launch is binned
        got = region.inside(pts)
        assert (got == want).all(), ("binned", n, d, sure, np.flatnonzero(got != want)[:5])
        assert (region.inside(pts[:777]) == want[:777]).all()
        eng.set_option(_native.OPT_BIN_MIN_ROWS, 0)          # never
        assert (region.inside(pts) == want).all(), ("unbinned on clustered tiles", n, d, sure)
    eng.set_option(_native.OPT_SURE_LEVEL, 1)
    # the integrator's in-place row patches (integrator.py:2749-2758) after the clustering
    eng.set_option(_native.OPT_BIN_MIN_ROWS, 1)
    rng = np.random.RandomState(5)
    for it in range(3):
        i = int(rng.randint(n))
        j = int(rng.randint(n))
        unew = region.u[j] + rng.normal(size=d) * 1e-4
        region.u[i] = unew
        region.unormed[i] = region.transformLayer.transform(unew)
        got = region.inside(pts)
        assert (got == _oracle(region, pts)).all(), ("after row patch", it)
    assert region.inside(region.u).all()


def test_binning_cuts_the_tile_visits_and_keeps_the_mask(eng):
    """BASELINE-size launch: same mask with and without bins; the binned launch streams fewer
    tiles per proposal (the reason it exists)."""
    import torch
    import ctypes
    from ultranest_b200 import _native
    region = _region(4000, 20, nboot=30)
    m = 1 << 19
    pts = np.vstack([_draw(region, m, 3, 1.0), _draw(region, m // 4, 4, 1.3)])[:m]
    region._bind()
    t = np.ascontiguousarray(region.transformLayer.transform(pts))
    t_dev = torch.from_numpy(t).cuda()
    masks, visits = [], []
    for thr in (0, 1 << 17):
        eng.set_option(_native.OPT_BIN_MIN_ROWS, thr)
        mk = torch.empty(m, dtype=torch.uint8, device="cuda")
        eng.call("unb_region_find_nearby_dev", t_dev.data_ptr(), m, None, mk.data_ptr(), None)
        eng.synchronize()
        masks.append(mk.cpu().numpy())
        visits.append(eng.stat(_native.STAT_TILE_VISITS))
    assert (masks[0] == masks[1]).all()
    assert visits[1] < 0.75 * visits[0], visits
    sel = np.random.RandomState(1).choice(m, 3000, replace=False)
    want = cport.find_nearby(region.unormed, t[sel], region.maxradiussq) >= 0
    assert (masks[1][sel].astype(bool) == want).all()
    # the fused host call (prep -> bins -> membership, chunked over two lanes) agrees as well
    eng.set_option(_native.OPT_BIN_MIN_ROWS, 1 << 17)
    full = region.inside(pts)
    eng.set_option(_native.OPT_BIN_MIN_ROWS, 0)
    assert (region.inside(pts) == full).all()
    assert (full[sel] == _oracle(region, pts[sel])).all()
