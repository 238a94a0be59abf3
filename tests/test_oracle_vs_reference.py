"""Pins the C oracle (oracle/mlfriends_oracle.c) to the compiled, unmodified reference
(oracle/_ref): every restated loop must be bit-identical on seeded inputs.
CPU-only; this is what makes the oracle trustworthy as the checker for the CUDA path."""
import numpy as np
import pytest

import oracle
from oracle import cport

pytestmark = pytest.mark.skipif(not oracle.reference_available(),
                                reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def ref():
    oracle.reference()
    import ultranest.mlfriends as m
    return m


def _points(rng, n, d, scale=1.0):
    return rng.normal(size=(n, d)) * scale


@pytest.mark.parametrize("d", [1, 2, 3, 5, 20, 50])
def test_find_nearby_bitexact(ref, d):
    rng = np.random.RandomState(d)
    a = _points(rng, 300, d)
    b = _points(rng, 500, d)
    # radius chosen so that a fair share of candidates has a neighbour
    dists = ((a[None, :50, :] - b[:, None, :]) ** 2).sum(axis=2)
    for r2 in [np.median(dists.min(axis=1)), dists.min(), 1e-90, 1e3]:
        out_ref = np.empty(len(b), dtype=np.int64)
        ref.find_nearby(a, b, r2, out_ref)
        out = cport.find_nearby(a, b, r2)
        assert (out == out_ref).all()


def test_find_nearby_edge_is_le(ref):
    """`d <= radiussq` edge: radius equal to an exactly computed distance must hit,
    the next double below must miss (mlfriends.pyx:181)."""
    rng = np.random.RandomState(7)
    a = _points(rng, 64, 20)
    b = _points(rng, 64, 20)
    d = 0.0
    for k in range(20):
        diff = a[17, k] - b[3, k]
        d = d + diff * diff
    for r2 in (d, np.nextafter(d, 0)):
        out_ref = np.empty(64, dtype=np.int64)
        ref.find_nearby(a[17:18], b[3:4], r2, out_ref[:1])
        out = cport.find_nearby(a[17:18], b[3:4], r2)
        assert out[0] == out_ref[0]
    assert cport.find_nearby(a[17:18], b[3:4], d)[0] == 0
    assert cport.find_nearby(a[17:18], b[3:4], np.nextafter(d, 0))[0] == -1


@pytest.mark.parametrize("n,d", [(50, 2), (400, 5), (300, 20)])
def test_subtract_nearby_bitexact(ref, n, d):
    rng = np.random.RandomState(n + d)
    u = rng.uniform(size=(n, d))
    r2 = 0.05 * d
    assert (cport.subtract_nearby(u, r2) == ref.subtract_nearby(u, r2)).all()


def test_count_nearby_consistent_with_model():
    rng = np.random.RandomState(3)
    a = _points(rng, 200, 7)
    b = _points(rng, 100, 7)
    r2 = 6.0
    cnt = cport.count_nearby(a, b, r2)
    model = np.zeros(len(b), dtype=np.int64)
    for j in range(len(b)):
        d = np.zeros(len(a))
        for k in range(7):
            diff = a[:, k] - b[j, k]
            d = d + diff * diff
        model[j] = (d <= r2).sum()
    assert (cnt == model).all()
    assert ((cnt > 0) == (cport.find_nearby(a, b, r2) >= 0)).all()


@pytest.mark.parametrize("n,d", [(400, 5), (1000, 20)])
def test_compute_maxradiussq_bootstrap_bitexact(ref, n, d):
    rng = np.random.RandomState(11)
    u = rng.uniform(0.3, 0.7, size=(n, d))
    layer = ref.AffineLayer()
    layer.optimize(u, u)
    region = ref.MLFriends(u, layer)
    np.random.seed(5)
    r_ref = region.compute_maxradiussq(nbootstraps=7)
    np.random.seed(5)
    r = 0
    for _ in range(7):
        sel = cport.draw_selection(np.random, n)
        r = max(r, cport.maxradiussq_selected(region.unormed, sel))
        # gathered form (what the Cython is actually handed) agrees with the masked form
        assert cport.maxradiussq(region.unormed[sel], region.unormed[~sel]) == \
            cport.maxradiussq_selected(region.unormed, sel)
    assert r == r_ref
    assert np.float32(r) == r  # float32-representable


@pytest.mark.parametrize("n,d", [(400, 5), (800, 20), (600, 100)])
def test_compute_enlargement_bitexact(ref, n, d):
    rng = np.random.RandomState(13)
    u = rng.uniform(0.3, 0.7, size=(n, d))
    layer = ref.AffineLayer()
    layer.optimize(u, u)
    region = ref.MLFriends(u, layer)
    r_ref, f_ref = region.compute_enlargement(nbootstraps=10, rng=np.random.RandomState(2))
    r, f = cport.compute_enlargement(u, region.unormed, 10, np.random.RandomState(2))
    assert r == r_ref
    assert f == f_ref


def test_mean_pair_distance_bitexact(ref):
    rng = np.random.RandomState(17)
    pts = _points(rng, 300, 6)
    ids = rng.randint(0, 4, size=300).astype(np.int64)
    assert cport.mean_pair_distance(pts, ids) == ref.compute_mean_pair_distance(pts, ids)


@pytest.mark.parametrize("d", [1, 2, 5, 20, 50, 90, 91, 100, 128, 150])
def test_inside_ellipsoid_bitexact(ref, d):
    rng = np.random.RandomState(19 + d)
    # d > 90: NumPy's buffered einsum restarts its partial sum every 8192 // d rows
    pts = rng.uniform(size=(2000 if d <= 50 else 400, d))
    ctr = rng.uniform(0.4, 0.6, size=d)
    A = rng.normal(size=(d, d))
    invcov = A @ A.T / d + np.eye(d)
    delta = pts - ctr
    r_np = np.einsum('ij,jk,ik->i', delta, invcov, delta)
    radius = np.median(r_np)
    mask, r = cport.inside_ellipsoid(pts, ctr, invcov, radius, return_r=True)
    assert (r == r_np).all()
    assert (mask == ref._inside_ellipsoid(pts, ctr, invcov, radius)).all()
    # `<=` edge
    j = int(np.argmin(np.abs(r_np - radius)))
    assert cport.inside_ellipsoid(pts[j:j + 1], ctr, invcov, r_np[j])[0]
    assert not cport.inside_ellipsoid(pts[j:j + 1], ctr, invcov, np.nextafter(r_np[j], 0))[0]


@pytest.mark.parametrize("n", list(range(0, 40)) + [127, 128, 129, 200, 257, 1000])
def test_numpy_pairwise_sum_model(n):
    rng = np.random.RandomState(n)
    a = rng.normal(size=(3, n)) * 10.0 ** rng.randint(-3, 4, size=(3, n))
    s = a.sum(axis=1)
    for i in range(3):
        assert cport.np_pairwise_sum(a[i]) == s[i]


@pytest.mark.parametrize("d", [1, 5, 7, 8, 9, 20, 100, 150])
def test_loglike_gauss_bitexact(d):
    rng = np.random.RandomState(d)
    theta = rng.uniform(size=(500, d))
    sigma = 0.01
    centers = np.ones(d) * 0.5
    like = -0.5 * (((theta - centers) / sigma)**2).sum(axis=1) \
        - 0.5 * np.log(2 * np.pi * sigma**2) * d
    assert (cport.loglike_gauss(theta, centers, sigma) == like).all()


@pytest.mark.parametrize("d", [2, 5, 10, 50])
def test_loglike_rosenbrock_bitexact(d):
    rng = np.random.RandomState(d)
    theta = rng.uniform(size=(500, d)) * 20 - 10
    a = theta[:, :-1]
    b = theta[:, 1:]
    like = -2 * (100 * (b - a**2)**2 + (1 - a)**2).sum(axis=1)
    assert (cport.loglike_rosenbrock(theta) == like).all()


def test_loglike_eggbox_close():
    rng = np.random.RandomState(1)
    z = rng.uniform(size=(500, 10)) * 10 * np.pi
    like = (2. + np.cos(z / 2.).prod(axis=1))**5
    np.testing.assert_allclose(cport.loglike_eggbox(z), like, rtol=1e-13)


def test_transforms(ref):
    rng = np.random.RandomState(23)
    u = rng.uniform(0.3, 0.7, size=(500, 20))
    layer = ref.AffineLayer()
    layer.optimize(u, u)
    t_ref = layer.transform(u)
    t = cport.transform_affine(u, layer.ctr, layer.T)
    # defined-order transform vs OpenBLAS dgemm: not bit-pinnable
    np.testing.assert_allclose(t, t_ref, rtol=0, atol=1e-13 * np.abs(t_ref).max())
    # single row == row of the batch, bit for bit (the property the product relies on)
    assert (cport.transform_affine(u[3], layer.ctr, layer.T) == t[3]).all()
    back = cport.untransform_affine(t, layer.ctr, layer.invT)
    np.testing.assert_allclose(back, u, rtol=1e-12)
    s = ref.ScalingLayer()
    s.optimize(u, u)
    assert (cport.transform_scaling(u, s.mean, s.std) == s.transform(u)).all()


def test_region_inside_pipeline(ref):
    """Oracle restatement of MLFriends.inside vs the reference, same np.dot transform."""
    rng = np.random.RandomState(29)
    u = rng.uniform(0.4, 0.6, size=(400, 5))
    layer = ref.AffineLayer()
    layer.optimize(u, u)
    region = ref.MLFriends(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(
        nbootstraps=10, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    pts = rng.uniform(0.35, 0.65, size=(5000, 5))
    mask_ref = region.inside(pts)
    mask = cport.region_inside(pts, region.unormed, layer.transform, region.maxradiussq,
                               region.ellipsoid_center, region.ellipsoid_invcov,
                               region.enlarge)
    assert (mask == mask_ref).all()
    assert 0 < mask.sum() < len(mask)
