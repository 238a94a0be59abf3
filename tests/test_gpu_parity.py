"""Parity tests proper: the CUDA path, called through the C ABI (ctypes), against the CPU
oracle (oracle/mlfriends_oracle.c, itself pinned to the compiled reference) on identical seeded
inputs.  Bar: bit-exact for indices, counts, masks and every fp64 result whose operation order
the reference fixes; stated tolerances only where the reference itself is BLAS/libm ordered."""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ultranest_b200 import _native
    return _native.get_engine()


def _live(rng, n, d, scale=1.0):
    """Roughly whitened cloud (what MLFriends.unormed looks like): uniform in a ball."""
    z = rng.normal(size=(n, d))
    z /= np.sqrt((z**2).sum(axis=1, keepdims=True))
    z *= rng.uniform(size=(n, 1))**(1.0 / d)
    return np.ascontiguousarray(z * scale)


def _radius_for(a, b, frac=0.5):
    """A radius such that about `frac` of the candidates have a neighbour."""
    sub = a[: min(len(a), 200)]
    sb = b[: min(len(b), 200)]
    dist = ((sub[None, :, :] - sb[:, None, :])**2).sum(axis=2).min(axis=1)
    return float(np.quantile(dist, frac)) if len(dist) else 1.0


SHAPES = [
    # (na, nb, d)
    (1, 5, 1), (63, 127, 2), (64, 128, 3), (65, 129, 5), (300, 1000, 7), (257, 513, 8),
    (1000, 3000, 12), (400, 2000, 16), (4000, 3000, 20), (500, 700, 24), (300, 300, 31),
    (300, 300, 32), (300, 500, 33), (200, 300, 50), (150, 260, 100), (100, 100, 130),
]


@pytest.mark.parametrize("na,nb,d", SHAPES)
def test_find_count_bitexact(eng, na, nb, d):
    rng = np.random.RandomState(na * 7 + nb * 3 + d)
    a = _live(rng, na, d)
    b = _live(rng, nb, d, scale=1.1)
    for frac in (0.02, 0.5, 0.98):
        r2 = _radius_for(a, b, frac)
        want = cport.find_nearby(a, b, r2)
        got = eng.find_nearby(a, b, r2)
        assert (got == want).all(), (na, nb, d, frac, np.flatnonzero(got != want)[:5])
        assert (eng.count_nearby(a, b, r2) == cport.count_nearby(a, b, r2)).all()
    # degenerate radii
    assert (eng.find_nearby(a, b, 1e-90) == cport.find_nearby(a, b, 1e-90)).all()
    assert (eng.find_nearby(a, b, 1e3) == cport.find_nearby(a, b, 1e3)).all()


def test_find_nearby_out_param_and_empty(eng):
    from ultranest_b200 import mlfriends as m
    rng = np.random.RandomState(1)
    a = _live(rng, 100, 4)
    b = _live(rng, 50, 4)
    out = np.empty(50, dtype=np.int64)
    m.find_nearby(a, b, 0.3, out)
    assert (out == cport.find_nearby(a, b, 0.3)).all()
    # nb = 0 is a no-op, na = 0 writes -1
    m.find_nearby(a, b[:0], 0.3, np.empty(0, dtype=np.int64))
    out[:] = 7
    m.find_nearby(a[:0], b, 0.3, out)
    assert (out == -1).all()
    # identical point with a vanishing radius still matches (test_regionsampling.py:46-48)
    assert (eng.find_nearby(a, a, 1e-90) == np.arange(100)).all()


def test_find_nearby_le_edge_through_filter(eng):
    """`<=` edge: a radius equal to an exactly computed reference distance must hit, the next
    double below must miss -- for many pairs, so the filter's slack is exercised on both sides."""
    rng = np.random.RandomState(5)
    d = 20
    a = _live(rng, 256, d)
    b = _live(rng, 64, d)
    for j in range(0, 64, 7):
        i = int(rng.randint(256))
        D = 0.0
        for k in range(d):
            diff = a[i, k] - b[j, k]
            D = D + diff * diff
        only = a[i:i + 1]
        for r2, expect in ((D, 0), (np.nextafter(D, 0), -1), (np.nextafter(D, 10), 0)):
            got = eng.find_nearby(only, b[j:j + 1], r2)
            assert got[0] == expect == cport.find_nearby(only, b[j:j + 1], r2)[0]
        # inside a full scan too
        assert (eng.find_nearby(a, b, D) == cport.find_nearby(a, b, D)).all()
        assert (eng.find_nearby(a, b, np.nextafter(D, 0)) == cport.find_nearby(a, b, np.nextafter(D, 0))).all()


def test_filter_rechecks_are_rare_and_exact_only_agrees(eng):
    from ultranest_b200 import _native
    rng = np.random.RandomState(9)
    a = _live(rng, 4000, 20)
    b = _live(rng, 4096, 20, scale=1.3)
    r2 = _radius_for(a, b, 0.3)
    fast = eng.find_nearby(a, b, r2)
    rechecks = eng.stat(_native.STAT_RECHECKS)
    hits = int((fast >= 0).sum())
    # every true hit is rechecked once; false alarms of the filter must be a tiny minority
    assert hits <= rechecks <= hits + max(64, hits // 10)
    eng.set_option(_native.OPT_EXACT_ONLY, 1)
    try:
        slow = eng.find_nearby(a, b, r2)
        cnt_slow = eng.count_nearby(a, b[:512], r2)
    finally:
        eng.set_option(_native.OPT_EXACT_ONLY, 0)
    assert (fast == slow).all()
    assert (eng.count_nearby(a, b[:512], r2) == cnt_slow).all()
    assert (fast[:1024] == cport.find_nearby(a, b[:1024], r2)).all()


@pytest.mark.parametrize("n,d", [(1, 3), (50, 2), (400, 5), (1000, 20), (300, 33), (200, 64)])
def test_subtract_nearby_bitexact(eng, n, d):
    rng = np.random.RandomState(n + d)
    u = rng.uniform(size=(n, d))
    for r2 in (0.02 * d, 0.08 * d):
        assert (eng.subtract_nearby(u, r2) == cport.subtract_nearby(u, r2)).all()


@pytest.mark.parametrize("na,nb,d", [(1, 1, 1), (100, 60, 2), (253, 147, 5), (2528, 1472, 20),
                                     (700, 300, 24), (300, 200, 40), (100, 80, 100)])
def test_compute_maxradiussq_bitexact(eng, na, nb, d):
    rng = np.random.RandomState(na + nb + d)
    a = _live(rng, na, d)
    b = _live(rng, nb, d)
    got = eng.compute_maxradiussq(a, b)
    assert got == cport.maxradiussq(a, b)
    assert np.float32(got) == got


@pytest.mark.parametrize("d", [1, 2, 5, 20, 50, 100])
def test_inside_ellipsoid_bitexact(eng, d):
    rng = np.random.RandomState(19 + d)
    pts = rng.uniform(size=(3000, d))
    ctr = rng.uniform(0.4, 0.6, size=d)
    A = rng.normal(size=(d, d))
    invcov = A @ A.T / d + np.eye(d)
    _, r = cport.inside_ellipsoid(pts, ctr, invcov, 1.0, return_r=True)
    for radius in (np.median(r), r.min(), np.nextafter(r.min(), 0), r.max()):
        assert (eng.inside_ellipsoid(pts, ctr, invcov, radius)
                == cport.inside_ellipsoid(pts, ctr, invcov, radius)).all()


@pytest.mark.parametrize("d", [1, 2, 5, 20, 50, 100])
def test_transforms_bitexact_vs_defined_order(eng, d):
    from ultranest_b200 import _native
    rng = np.random.RandomState(23 + d)
    w = rng.uniform(size=(777, d))
    ctr = rng.uniform(0.4, 0.6, size=d)
    T = rng.normal(size=(d, d))
    t = eng.transform(_native.LAYER_AFFINE, False, w, ctr, T)
    assert (t == cport.transform_affine(w, ctr, T)).all()
    np.testing.assert_allclose(t, np.dot(w - ctr, T), rtol=0, atol=1e-13 * np.abs(t).max())
    assert (eng.transform(_native.LAYER_AFFINE, False, w[5], ctr, T) == t[5]).all()
    back = eng.transform(_native.LAYER_AFFINE, True, t, ctr, np.linalg.inv(T))
    assert (back == cport.untransform_affine(t, ctr, np.linalg.inv(T))).all()
    mean = rng.uniform(size=d)
    std = rng.uniform(0.1, 2, size=d)
    s = eng.transform(_native.LAYER_SCALING, False, w, mean, std)
    assert (s == (w - mean) / std).all()
    assert (eng.transform(_native.LAYER_SCALING, True, s, mean, std) == s * std + mean).all()


@pytest.mark.parametrize("d", [1, 5, 7, 8, 9, 20, 100, 150, 300])
def test_loglikes(eng, d):
    rng = np.random.RandomState(d)
    theta = rng.uniform(size=(1000, d))
    sigma = 0.01
    centers = np.ones(d) * 0.5
    like = -0.5 * (((theta - centers) / sigma)**2).sum(axis=1) - 0.5 * np.log(2 * np.pi * sigma**2) * d
    got = eng.loglike_gauss(theta, centers, sigma, 0.5 * np.log(2 * np.pi * sigma**2) * d)
    assert (got == like).all()
    assert (got == cport.loglike_gauss(theta, centers, sigma)).all()
    if d >= 2:
        th = theta * 20 - 10
        a = th[:, :-1]
        b = th[:, 1:]
        assert (eng.loglike_rosenbrock(th) == -2 * (100 * (b - a**2)**2 + (1 - a)**2).sum(axis=1)).all()
    z = theta * 10 * np.pi
    np.testing.assert_allclose(eng.loglike_eggbox(z), (2. + np.cos(z / 2.).prod(axis=1))**5, rtol=1e-13)


def test_mean_pair_distance(eng):
    rng = np.random.RandomState(17)
    pts = _live(rng, 500, 6)
    ids = rng.randint(0, 4, size=500).astype(np.int64)
    np.testing.assert_allclose(eng.mean_pair_distance(pts, ids), cport.mean_pair_distance(pts, ids),
                               rtol=1e-12)


def _make_region(n, d, seed, layer_cls_name="AffineLayer", nboot=10):
    from ultranest_b200 import mlfriends as m
    rng = np.random.RandomState(seed)
    z = _live(rng, n, d)
    L = np.linalg.cholesky(0.5 * np.ones((d, d)) + 0.5 * np.eye(d))
    u = 0.5 + 0.05 * z @ L.T
    layer = getattr(m, layer_cls_name)()
    layer.optimize(u, u)
    region = m.MLFriends(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(
        nbootstraps=nboot, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    return region, rng


@pytest.mark.parametrize("n,d,layer", [(400, 5, "AffineLayer"), (4000, 20, "AffineLayer"),
                                       (500, 3, "ScalingLayer"), (300, 40, "AffineLayer")])
def test_region_inside_bitexact(eng, n, d, layer):
    """MLFriends.inside (fused device pipeline) vs the oracle restatement of
    mlfriends.pyx:1186-1211 on the same region state; three candidate regimes."""
    region, rng = _make_region(n, d, 31 + d, layer)
    lay = region.transformLayer
    if layer == "AffineLayer":
        xf = lambda p: cport.transform_affine(p, lay.ctr, lay.T)   # noqa: E731
    else:
        xf = lambda p: cport.transform_scaling(p, lay.mean, lay.std)   # noqa: E731
    # (A) accepting: draws inside the wrapping ellipsoid
    z = _live(rng, 3000, d) * region.enlarge**0.5
    pts_a = region.ellipsoid_center + z @ region.ellipsoid_axes_T
    # (B) box around the live points; (C) the live points themselves
    lo, hi = region.u.min(axis=0), region.u.max(axis=0)
    pts_b = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo), size=(3000, d))
    for pts in (pts_a, pts_b, region.u.copy()):
        want = cport.region_inside(pts, region.unormed, xf, region.maxradiussq,
                                   region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
        got = region.inside(pts)
        assert got.dtype == bool and got.shape == (len(pts),)
        assert (got == want).all()
    assert region.inside(region.u).all()
    # first-neighbour index is exposed too (integrator.py:2041 tests `!= 0`)
    mask, idx = eng.region_inside(pts_a, want_index=True)
    t = xf(pts_a)
    ell = cport.inside_ellipsoid(pts_a, region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
    want_idx = np.where(ell, cport.find_nearby(region.unormed, t, region.maxradiussq), -1)
    assert (idx == want_idx).all() and (mask == (want_idx >= 0)).all()


def test_region_mirror_follows_inplace_mutation(eng):
    """integrator.py:2749-2758 patches region.u / region.unormed rows and the ellipsoid centre in
    place; the device mirror must follow without being told."""
    region, rng = _make_region(1000, 8, 77)
    lay = region.transformLayer
    xf = lambda p: cport.transform_affine(p, lay.ctr, lay.T)   # noqa: E731
    pts = region.u[rng.randint(1000, size=2000)] + rng.normal(size=(2000, 8)) * 0.004
    assert (region.inside(pts) == cport.region_inside(
        pts, region.unormed, xf, region.maxradiussq, region.ellipsoid_center,
        region.ellipsoid_invcov, region.enlarge)).all()
    for it in range(5):
        worst = int(rng.randint(1000))
        unew = region.u[int(rng.randint(1000))] + rng.normal(size=8) * 0.001
        region.u[worst] = unew
        region.unormed[worst] = region.transformLayer.transform(unew)
        region.ellipsoid_center = np.mean(region.u, axis=0)
        want = cport.region_inside(pts, region.unormed, xf, region.maxradiussq,
                                   region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
        assert (region.inside(pts) == want).all(), it
    region.maxradiussq = region.maxradiussq * 0.5
    want = cport.region_inside(pts, region.unormed, xf, region.maxradiussq,
                               region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
    assert (region.inside(pts) == want).all()


@pytest.mark.parametrize("n,d", [(400, 5), (2000, 10), (4000, 20), (300, 40)])
def test_bootstrap_matches_oracle(eng, n, d):
    """All rounds in one launch == the oracle's round-by-round loop, bit for bit
    (radius float32-rounded per round, enlargement in the einsum order)."""
    region, _ = _make_region(n, d, 101 + d, nboot=3)
    got = region.compute_enlargement(nbootstraps=12, rng=np.random.RandomState(4))
    want = cport.compute_enlargement(region.u, region.unormed, 12, np.random.RandomState(4))
    assert got == want
    np.random.seed(6)
    r = region.compute_maxradiussq(nbootstraps=5)
    np.random.seed(6)
    want_r = 0
    for _ in range(5):
        want_r = max(want_r, cport.maxradiussq_selected(region.unormed, cport.draw_selection(np.random, n)))
    assert r == want_r


def test_inside_and_loglike_fused(eng):
    from ultranest_b200.likelihoods import GaussianLogLike
    region, rng = _make_region(2000, 10, 55)
    z = _live(rng, 20000, 10) * region.enlarge**0.5 * 1.2
    pts = region.ellipsoid_center + z @ region.ellipsoid_axes_T
    loglike = GaussianLogLike(0.5, 0.05)
    mask, like = region.inside_and_loglike(pts, loglike)
    assert (mask == region.inside(pts)).all()
    assert 0 < mask.sum() < len(mask)
    want = -0.5 * (((pts[mask] - 0.5) / 0.05)**2).sum(axis=1) - 0.5 * np.log(2 * np.pi * 0.05**2) * 10
    assert (like[mask] == want).all()
    assert np.isneginf(like[~mask]).all()
    # chunked pipeline (several chunks through both lanes) gives the same answer
    from ultranest_b200 import _native
    eng.set_option(_native.OPT_CHUNK_ROWS, 3000)
    try:
        mask2, like2 = region.inside_and_loglike(pts, loglike)
    finally:
        eng.set_option(_native.OPT_CHUNK_ROWS, 0)
    assert (mask2 == mask).all() and (like2[mask] == like[mask]).all()


def test_membership_fp32_prefilter_agrees_and_edges(eng):
    """The any-neighbour kernel with the fp32 pre-filter, with the fp64 filter and the ordered
    first-index kernel must give the same masks -- including radii set exactly ON reference
    distances (the `<=` edge) and one ulp below, which the filters may never decide themselves."""
    from ultranest_b200 import _native
    rng = np.random.RandomState(77)
    d = 20
    a = _live(rng, 4000, d)
    b = _live(rng, 30000, d, scale=1.1)
    eng.region_sync_live(a)
    radii = [_radius_for(a, b, 0.3)]
    for j in (3, 1234, 20000):           # exact reference distances of real (nearest) pairs
        i = int(np.argmin(((a - b[j])**2).sum(axis=1)))
        D = 0.0
        for k in range(d):
            diff = a[i, k] - b[j, k]
            D = D + diff * diff
        radii += [D, np.nextafter(D, 0)]
    for r2 in radii:
        eng.region_set_radius(r2)
        want = cport.find_nearby(a, b, r2) >= 0
        assert want.any()
        got = {}
        for flag in (1, 0):
            eng.set_option(_native.OPT_FILTER_FP32, flag)
            try:
                got[flag] = eng.region_has_neighbour(b)
                rechecks = eng.stat(_native.STAT_RECHECKS)
            finally:
                eng.set_option(_native.OPT_FILTER_FP32, 1)
            assert (got[flag] == want).all(), (r2, flag)
            # the filter is tight: hardly more exact evaluations than hits
            assert rechecks <= want.sum() * 1.2 + 256, (rechecks, want.sum())
        assert ((eng.region_find_nearby(b) >= 0) == want).all()
    # magnitudes outside the fp32 comfort zone silently take the fp64 filter
    scale = 1e-20
    eng.region_sync_live(a * scale)
    eng.region_set_radius(radii[0] * scale * scale)
    assert (eng.region_has_neighbour(b * scale) == (cport.find_nearby(a * scale, b * scale, radii[0] * scale * scale) >= 0)).all()
    # a radius far below the data scale (false-alarm shell too thick for fp32) as well
    eng.region_sync_live(a)
    eng.region_set_radius(1e-9)
    assert (eng.region_has_neighbour(a[:500] + 1e-6) == (cport.find_nearby(a, a[:500] + 1e-6, 1e-9) >= 0)).all()


@pytest.mark.parametrize("na,nb,d", [(4000, 20000, 20), (3000, 9000, 5), (1200, 6000, 50), (700, 5000, 100)])
def test_find_nearby_two_phase_bitexact(eng, na, nb, d):
    """Launches of >= 4096 candidates take the two-phase path (membership kernel first, ordered
    first-index scan over the members only): indices must be the ordered exact kernel's, in the
    accepting, mixed and full-scan regimes, and the `<=` edge must survive both phases."""
    from ultranest_b200 import _native
    rng = np.random.RandomState(na + nb + d)
    a = _live(rng, na, d)
    b = _live(rng, nb, d, scale=1.15)
    for frac in (0.0, 0.3, 0.97):
        r2 = _radius_for(a, b, frac) if frac > 0 else 1e-6
        want = cport.find_nearby(a, b, r2)
        got = eng.find_nearby(a, b, r2)
        assert (got == want).all(), (na, nb, d, frac, np.flatnonzero(got != want)[:5])
    # the edge: radius = an exactly computed pair distance (hit) and the next double below (miss)
    j, i = 17, int(cport.find_nearby(a, b[17:18], 1e9)[0])
    D = 0.0
    for k in range(d):
        diff = a[i, k] - b[j, k]
        D = D + diff * diff
    for r2 in (D, np.nextafter(D, 0)):
        assert (eng.find_nearby(a, b, r2) == cport.find_nearby(a, b, r2)).all()
    # exact-only mode (plain kernels, no two-phase) agrees
    r2 = _radius_for(a, b, 0.3)
    fast = eng.find_nearby(a, b, r2)
    eng.set_option(_native.OPT_EXACT_ONLY, 1)
    try:
        assert (eng.find_nearby(a, b, r2) == fast).all()
    finally:
        eng.set_option(_native.OPT_EXACT_ONLY, 0)
