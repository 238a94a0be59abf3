"""Edge cases and size-independent properties of the CUDA path: empty and ragged inputs, tiny and
large dimensions (all three kernel families), launch sizes around the chunk / tile / block
boundaries, full-size (2^20 proposals) consistency properties, error reporting."""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ultranest_b200 import _native
    return _native.get_engine()


def _ball(rng, n, d):
    z = rng.normal(size=(n, d))
    z /= np.sqrt((z**2).sum(axis=1, keepdims=True))
    return np.ascontiguousarray(z * rng.uniform(size=(n, 1))**(1.0 / d))


def test_empty_and_single(eng):
    a = _ball(np.random.RandomState(0), 10, 3)
    assert eng.find_nearby(a, a[:0], 1.0).shape == (0,)
    assert (eng.find_nearby(a[:0], a, 1.0) == -1).all()
    assert (eng.count_nearby(a[:0], a, 1.0) == 0).all()
    assert eng.compute_maxradiussq(a, a[:0]) == 0.0
    assert eng.find_nearby(a[:1], a[:1], 0.0)[0] == 0
    assert eng.inside_ellipsoid(a[:0], np.zeros(3), np.eye(3), 1.0).shape == (0,)
    assert eng.loglike_rosenbrock(np.zeros((0, 4))).shape == (0,)
    # non-contiguous / fancy-indexed inputs are accepted like the Cython's callers pass them
    big = _ball(np.random.RandomState(1), 200, 6)
    view = big[::2, :]
    assert (eng.find_nearby(view, big[1::2], 0.4) == cport.find_nearby(view, big[1::2], 0.4)).all()


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 8, 9, 31, 32, 33, 64, 127, 260, 600])
def test_every_dimension_family(eng, d):
    """d <= 32: register kernels; 33..~110: tiled shared-memory kernel (tile 64/32/...); beyond:
    plain exact kernel.  All must agree with the oracle."""
    rng = np.random.RandomState(d)
    a = _ball(rng, 150, d)
    b = _ball(rng, 333, d) * 1.05
    dist = ((a[None, :40, :] - b[:40, None, :])**2).sum(axis=2).min(axis=1)
    r2 = float(np.median(dist))
    assert (eng.find_nearby(a, b, r2) == cport.find_nearby(a, b, r2)).all()
    assert (eng.count_nearby(a, b, r2) == cport.count_nearby(a, b, r2)).all()
    assert (eng.has_neighbour(a, b, r2) == (cport.find_nearby(a, b, r2) >= 0)).all()
    assert eng.compute_maxradiussq(a, b) == cport.maxradiussq(a, b)
    if d <= 260:
        assert (eng.subtract_nearby(a, r2) == cport.subtract_nearby(a, r2)).all()
        ctr = rng.uniform(-0.1, 0.1, size=d)
        A = np.eye(d) + 0.1 * np.ones((d, d)) / d
        _, r = cport.inside_ellipsoid(b, ctr, A, 1.0, return_r=True)
        assert (eng.inside_ellipsoid(b, ctr, A, np.median(r))
                == cport.inside_ellipsoid(b, ctr, A, np.median(r))).all()


@pytest.mark.parametrize("m", [1, 31, 32, 33, 127, 128, 129, 255, 256, 257, 1000, 4097])
def test_launch_size_boundaries(eng, m):
    rng = np.random.RandomState(m)
    a = _ball(rng, 777, 20)
    b = _ball(rng, m, 20) * 1.02
    r2 = 0.55
    want = cport.find_nearby(a, b, r2)
    assert (eng.find_nearby(a, b, r2) == want).all()
    assert (eng.count_nearby(a, b, r2) == cport.count_nearby(a, b, r2)).all()


def _region(n, d, seed):
    from ultranest_b200 import mlfriends as mm
    rng = np.random.RandomState(seed)
    L = np.linalg.cholesky(0.5 * np.ones((d, d)) + 0.5 * np.eye(d))
    u = 0.5 + 0.05 * _ball(rng, n, d) @ L.T
    layer = mm.AffineLayer()
    layer.optimize(u, u)
    region = mm.MLFriends(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(10, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    return region, rng


@pytest.mark.parametrize("n,d", [(1, 2), (2, 2), (63, 3), (64, 3), (65, 3), (129, 7)])
def test_tiny_live_sets(eng, n, d):
    """Live blocks around the tile size (64), including a single live point."""
    rng = np.random.RandomState(n * 10 + d)
    a = _ball(rng, n, d)
    b = _ball(rng, 500, d)
    for r2 in (0.05, 0.5, 4.0):
        assert (eng.find_nearby(a, b, r2) == cport.find_nearby(a, b, r2)).all()
    from ultranest_b200 import mlfriends as mm
    eng.region_sync_live(a)
    eng.region_set_radius(0.5)
    want = cport.find_nearby(a, b, 0.5)
    assert (eng.region_has_neighbour(b) == (want >= 0)).all()
    assert (eng.region_find_nearby(b) == want).all()


@pytest.mark.parametrize("n,d,m", [(70, 20, 3), (70, 20, 40000), (4000, 20, 300000), (130, 5, 70000), (100, 50, 20000)])
def test_membership_odd_candidates_and_drain(eng, n, d, m):
    """Proposals the fp32 pre-filter cannot bound (zero norm, norm beyond the fp32 range: they flag
    every pair, padded tile slots included) and launches whose tail runs through the cooperative
    drain: the mask is still the oracle's."""
    rng = np.random.RandomState(n + d + m)
    a = 0.3 + 0.2 * _ball(rng, n, d)            # the origin is far outside every ball
    b = 0.3 + 0.25 * _ball(rng, m, d)
    b[0] = 0.0                                   # |b|^2 = 0
    b[1] = 1e16                                  # |b|^2 > 1e30
    b[2] = a[n - 1]                              # a live point itself
    if m > 10:
        b[m // 2] = 0.0
        b[m - 1] = -3e15
    r2 = 0.02
    eng.region_sync_live(a)
    eng.region_set_radius(r2)
    got = eng.region_has_neighbour(b)
    sub = np.unique(np.concatenate([np.arange(min(m, 3000)), [m // 2, m - 1]]))
    want = cport.find_nearby(a, b[sub], r2) >= 0
    assert (got[sub] == want).all()
    assert not got[0] and not got[1] and got[2]
    # every proposal through the exact-only kernels as well
    from ultranest_b200 import _native
    eng.set_option(_native.OPT_EXACT_ONLY, 1)
    try:
        eng.region_sync_live(a)
        assert (eng.region_has_neighbour(b) == got).all()
    finally:
        eng.set_option(_native.OPT_EXACT_ONLY, 0)


def test_chunked_host_pipeline_boundaries(eng):
    """inside() through several chunk sizes (incl. a ragged last chunk, both lanes) and through
    pageable as well as pinned caller buffers."""
    import torch
    from ultranest_b200 import _native
    region, rng = _region(1500, 12, 3)
    z = _ball(rng, 10007, 12) * region.enlarge**0.5 * 1.15
    pts = region.ellipsoid_center + z @ region.ellipsoid_axes_T
    lay = region.transformLayer
    want = cport.region_inside(pts, region.unormed, lambda p: cport.transform_affine(p, lay.ctr, lay.T),
                               region.maxradiussq, region.ellipsoid_center, region.ellipsoid_invcov,
                               region.enlarge)
    assert 0 < want.sum() < len(want)
    for chunk in (0, 1000, 4096, 10007, 10006, 7):
        if chunk == 7 and len(pts) > 2000:
            sub, wsub = pts[:50], want[:50]      # tiny chunks on a short batch
        else:
            sub, wsub = pts, want
        eng.set_option(_native.OPT_CHUNK_ROWS, chunk)
        try:
            assert (region.inside(sub) == wsub).all(), chunk
            mask, idx = eng.region_inside(sub, want_index=True)
            assert (mask == wsub).all() and ((idx >= 0) == wsub).all()
        finally:
            eng.set_option(_native.OPT_CHUNK_ROWS, 0)
    pin = torch.empty(pts.shape, dtype=torch.float64).pin_memory()
    pin.numpy()[...] = pts
    assert (region.inside(pin.numpy()) == want).all()


def test_full_size_properties(eng):
    """BASELINE-size batch (2^20 x 20 against N_live=4000): properties that need no oracle --
    the any-neighbour and first-index kernels agree; masks are permutation-equivariant and
    idempotent across calls; live points are members; a point far outside is not; the count is
    positive exactly where a neighbour exists (sampled)."""
    region, rng = _region(4000, 20, 9)
    m = 1 << 20
    z = _ball(rng, m, 20) * region.enlarge**0.5 * 1.08
    pts = region.ellipsoid_center + z @ region.ellipsoid_axes_T
    mask = region.inside(pts)
    assert 0.2 < mask.mean() < 1.0
    mask_b, idx = eng.region_inside(pts, want_index=True)
    assert (mask_b == mask).all() and ((idx >= 0) == mask).all()
    assert (region.inside(pts) == mask).all()                       # idempotent
    perm = rng.permutation(m)
    assert (region.inside(pts[perm]) == mask[perm]).all()           # row order does not matter
    assert region.inside(region.u).all()
    assert not region.inside(np.full((4, 20), 0.999)).any()
    # spot check 4096 rows against the oracle, and counts against membership
    lay = region.transformLayer
    sel = rng.choice(m, 4096, replace=False)
    want = cport.region_inside(pts[sel], region.unormed,
                               lambda p: cport.transform_affine(p, lay.ctr, lay.T),
                               region.maxradiussq, region.ellipsoid_center,
                               region.ellipsoid_invcov, region.enlarge)
    assert (mask[sel] == want).all()
    t = cport.transform_affine(pts[sel], lay.ctr, lay.T)
    cnt = eng.region_count_nearby(t)
    assert ((cnt > 0) == (cport.find_nearby(region.unormed, t, region.maxradiussq) >= 0)).all()


def test_error_reporting(eng):
    from ultranest_b200 import _native
    fresh = _native.Engine(eng.device)
    try:
        with pytest.raises(RuntimeError):
            fresh.region_inside(np.zeros((3, 2)) + 0.5)             # region state not set
        with pytest.raises(ValueError):
            fresh.find_nearby(np.zeros((3, 2)), np.zeros((3, 3)), 1.0)   # ragged dimensionality
        with pytest.raises(ValueError):
            fresh.inside_ellipsoid(np.zeros((3, 2)), np.zeros(3), np.eye(2), 1.0)
        with pytest.raises(ValueError):
            fresh.loglike_rosenbrock(np.zeros((3, 1)))
    finally:
        fresh.close()
    # the shared engine is unaffected
    a = _ball(np.random.RandomState(0), 10, 3)
    assert (eng.find_nearby(a, a, 1e-90) == np.arange(10)).all()
