"""The reference's region test scenarios (tests/test_regionsampling.py, test_clustering.py,
test_transforms.py) on the device-backed classes, plus a differential run of every sampling
method against the reference classes under the same np.random seed (same RNG order -> same
proposals)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ours():
    from ultranest_b200 import mlfriends
    return mlfriends


@pytest.fixture(scope="module")
def ref():
    if not oracle.reference_available():
        pytest.skip("oracle/_ref not built")
    oracle.reference()
    import ultranest.mlfriends as m
    return m


def _region(mod, upoints, layer_name, nboot=30, seed=5):
    layer = getattr(mod, layer_name)(wrapped_dims=[])
    layer.optimize(upoints, upoints)
    region = mod.MLFriends(upoints, layer)
    np.random.seed(seed)
    region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=nboot)
    region.create_ellipsoid()
    return region


@pytest.mark.parametrize("layer_name,scale", [("ScalingLayer", (0.3, 0.03)), ("AffineLayer", (1.0, 0.5))])
def test_region_sampling_scenarios(ours, layer_name, scale):
    np.random.seed(1)
    if layer_name == "ScalingLayer":
        upoints = np.random.uniform(0.2, 0.5, size=(1000, 2))
        upoints[:, 1] *= 0.1
    else:
        upoints = np.random.uniform(size=(1000, 2))
        upoints[:, 1] *= 0.5
    region = _region(ours, upoints, layer_name)
    assert region.transformLayer.nclusters == 1
    assert np.allclose(region.unormed, region.transformLayer.transform(upoints))
    assert region.inside(upoints).all(), "live points should lie near live points"
    for method in region.sampling_methods:
        newpoints = method(nsamples=4000)
        assert len(newpoints) > 0
        lo1, lo2 = newpoints.min(axis=0)
        hi1, hi2 = newpoints.max(axis=0)
        if layer_name == "ScalingLayer":    # windows of test_regionsampling.py:38-43
            assert 0.15 < lo1 < 0.25 and 0.015 < lo2 < 0.025, (method.__name__, lo1, lo2)
            assert 0.45 < hi1 < 0.55 and 0.045 < hi2 < 0.055, (method.__name__, hi1, hi2)
        else:                               # windows of test_regionsampling.py:78-83
            assert 0 <= lo1 < 0.1 and 0 <= lo2 < 0.1, (method.__name__, lo1, lo2)
            assert 0.95 < hi1 <= 1 and 0.45 <= hi2 < 0.55, (method.__name__, hi1, hi2)
        assert region.inside(newpoints).mean() > 0.99, method.__name__
    region.maxradiussq = 1e-90
    assert region.inside(upoints).all(), "live points should lie very near themselves"


@pytest.mark.parametrize("layer_name", ["ScalingLayer", "AffineLayer"])
def test_sampling_methods_match_reference_stream(ours, ref, layer_name):
    np.random.seed(1)
    upoints = np.random.uniform(0.3, 0.7, size=(800, 3))
    upoints[:, 1] = 0.5 + (upoints[:, 1] - 0.5) * 0.3 + (upoints[:, 0] - 0.5) * 0.4
    ro = _region(ours, upoints, layer_name)
    rr = _region(ref, upoints, layer_name)
    assert ro.maxradiussq == pytest.approx(rr.maxradiussq, rel=1e-6)
    assert ro.enlarge == pytest.approx(rr.enlarge, rel=1e-10)
    for name in ("sample_from_boundingbox", "sample_from_wrapping_ellipsoid",
                 "sample_from_transformed_boundingbox", "sample_from_points"):
        np.random.seed(11)
        a = getattr(ro, name)(nsamples=3000)
        state_o = np.random.get_state()[1][:5].copy()
        np.random.seed(11)
        b = getattr(rr, name)(nsamples=3000)
        state_r = np.random.get_state()[1][:5].copy()
        assert (state_o == state_r).all(), "%s consumed the RNG differently" % name
        assert a.shape == b.shape, name
        if name in ("sample_from_boundingbox", "sample_from_wrapping_ellipsoid"):
            assert (a == b).all(), name      # proposals pass through untouched
        else:
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-13, err_msg=name)
    # sample(): same switching behaviour on an impossible region
    ro.maxradiussq = rr.maxradiussq = 1e-90
    np.random.seed(3)
    so = [ro.sample(50).shape for _ in range(4)]
    mo = ro.current_sampling_method.__name__
    np.random.seed(3)
    sr = [rr.sample(50).shape for _ in range(4)]
    assert so == sr and mo == rr.current_sampling_method.__name__


def test_inside_ellipsoid_equals_einsum(ours):
    np.random.seed(1)
    points = np.random.uniform(0.4, 0.6, size=(1000, 2))
    points[:, 1] *= 0.5
    region = _region(ours, points, "AffineLayer")
    bpts = np.random.uniform(size=(100, 2))
    d = bpts - region.ellipsoid_center
    mask2 = np.einsum('ij,jk,ik->i', d, region.ellipsoid_invcov, d) <= region.enlarge
    assert (region.inside_ellipsoid(bpts) == mask2).all()


def test_mean_pair_distance_scenario(ours):
    np.random.seed(1)
    points = np.random.uniform(0.4, 0.6, size=(10000, 2))
    ring = np.abs((points[:, 0] - 0.5)**2 + (points[:, 1] - 0.5)**2 - 0.08**2) < 0.02**2
    points = points[ring]
    points = points[points[:, 0] < 0.5]
    region = _region(ours, points, "AffineLayer")
    t = region.transformLayer.transform(region.u)
    total, npairs = 0.0, 0
    for i in range(len(t)):
        total += (((t[i, :] - t[:i, :])**2).sum(axis=1)**0.5).sum()
        npairs += i
    assert np.isclose(region.compute_mean_pair_distance(), total / npairs)


def test_all_region_classes_contain_their_points(ours):
    np.random.seed(4)
    for umax in (0.6, 0.5):
        points = np.random.uniform(0.4, 0.6, size=(1000, 3))
        points = points[points[:, 0] < umax]
        tpoints = points * 10
        tpoints[:, 0] = np.floor(tpoints[:, 0])      # a constant ("categorical") dimension
        layer = ours.AffineLayer(wrapped_dims=[])
        layer.optimize(points, points)
        for cls in (ours.MLFriends, ours.RobustEllipsoidRegion, ours.SimpleRegion):
            region = cls(points, layer)
            region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=30)
            region.create_ellipsoid()
            inside = region.inside(points)
            assert inside.shape == (len(points),) and inside.all(), cls.__name__
        tregion = ours.WrappingEllipsoid(tpoints)
        tregion.enlarge = tregion.compute_enlargement(nbootstraps=30)
        tregion.create_ellipsoid()
        inside = tregion.inside(tpoints)
        assert inside.shape == (len(tpoints),) and inside.all()
        assert not tregion.inside(tpoints + np.array([1.0, 0, 0])).any() or umax == 0.6
    one_d = ours.WrappingEllipsoid(np.random.uniform(0.4, 0.6, size=(1000, 1)))
    one_d.enlarge = one_d.compute_enlargement(nbootstraps=30)
    one_d.create_ellipsoid()


def test_ellipsoid_regions_match_reference(ours, ref):
    np.random.seed(8)
    points = np.random.uniform(0.3, 0.7, size=(600, 4))
    for name in ("RobustEllipsoidRegion", "SimpleRegion"):
        out = []
        for mod in (ours, ref):
            layer = mod.AffineLayer()
            layer.optimize(points, points)
            region = getattr(mod, name)(points, layer)
            r, f = region.compute_enlargement(nbootstraps=20, rng=np.random.RandomState(3))
            region.maxradiussq, region.enlarge = r, f
            region.create_ellipsoid()
            np.random.seed(2)
            out.append((r, f, region.sample(2000), region.estimate_volume()))
        assert out[0][0] == out[1][0] == 1e300
        assert out[0][1] == pytest.approx(out[1][1], rel=1e-12)
        assert out[0][2].shape == out[1][2].shape
        assert out[0][3] == pytest.approx(out[1][3], rel=1e-12)
    w_o, w_r = ours.WrappingEllipsoid(points * 3), ref.WrappingEllipsoid(points * 3)
    f_o = w_o.compute_enlargement(nbootstraps=20, rng=np.random.RandomState(3))
    f_r = w_r.compute_enlargement(nbootstraps=20, rng=np.random.RandomState(3))
    assert f_o == pytest.approx(f_r, rel=1e-11)


def test_errors_match_reference_types(ours):
    with pytest.raises(ValueError):
        ours.MLFriends(np.array([[0.5, 1.5], [0.2, 0.3]]), ours.ScalingLayer())
    pts = np.random.RandomState(1).uniform(0.4, 0.6, size=(3, 5))
    layer = ours.ScalingLayer()
    layer.optimize(pts, pts)
    with pytest.raises(FloatingPointError):
        ours.RobustEllipsoidRegion(pts, layer).compute_enlargement(nbootstraps=5)
    # linearly dependent points -> singular covariance -> LinAlgError (test_run.py:62-72)
    line = np.linspace(0.2, 0.8, 50).reshape((-1, 1)) * np.ones((1, 3))
    lay = ours.ScalingLayer()
    lay.optimize(line, line)
    with pytest.raises(np.linalg.LinAlgError):
        ours.MLFriends(line, lay).compute_enlargement(nbootstraps=5, rng=np.random.RandomState(1))


def test_clustering_scenarios(ours, ref):
    for i in range(5):
        np.random.seed(i * 100)
        points = np.random.uniform(size=(100, 2))
        for r2, lo, hi in ((0.1**2, 1, 30), (0.2**2, 0, 2)):
            n_o, ids_o, over_o = ours.update_clusters(points, points, r2)
            n_r, ids_r, over_r = ref.update_clusters(points, points, r2)
            assert lo < n_o < hi or (lo == 0 and n_o == 1)
            assert n_o == n_r and (ids_o == ids_r).all() and (over_o == over_r).all()
    rng = np.random.RandomState(2)
    u = rng.uniform(size=(20, 2))
    u[:10, :] += 10
    assert np.all(np.abs(ours.subtract_nearby(u, 1.0)) < 0.5)
    # re-use of old cluster ids
    pts = np.vstack([rng.normal(0.3, 0.01, size=(50, 2)), rng.normal(0.7, 0.01, size=(60, 2))])
    old = np.concatenate([2 * np.ones(50), np.ones(60)]).astype(np.int64)
    n_o, ids_o, _ = ours.update_clusters(pts, pts, 0.05**2, old)
    n_r, ids_r, _ = ref.update_clusters(pts, pts, 0.05**2, old)
    assert n_o == n_r == 2 and (ids_o == ids_r).all() and (ids_o == old).all()


def test_layers_roundtrip_and_create_new(ours, ref):
    np.random.seed(1)
    for corr in (0, 0.6, 0.95):
        cov = np.array([[1., corr], [corr, 1.]])
        points = np.random.multivariate_normal(np.zeros(2), cov, size=400) * 0.01 + 0.5
        s = ours.ScalingLayer()
        s.optimize(points, points)
        assert (s.untransform(s.transform(points)) == points).all()
        assert (s.untransform(s.transform(points[0])) == points[0]).all()
        a = ours.AffineLayer()
        a.optimize(points, points)
        back = a.untransform(a.transform(points))
        np.testing.assert_allclose(back, points, rtol=1e-13)
        assert a.transform(points[0]).shape == (2,)
        np.testing.assert_allclose(a.transform(points[0]), a.transform(points)[0], rtol=1e-12)
    # wrapped (circular) dimensions take the host wrap + device scans path
    pts = np.random.normal(0.5, 0.01, size=(300, 2))
    pts[:, 0] = np.fmod(pts[:, 0] + 0.5, 1)
    for mod_name in ("ScalingLayer", "AffineLayer"):
        lo = getattr(ours, mod_name)(wrapped_dims=[0])
        lr = getattr(ref, mod_name)(wrapped_dims=[0])
        lo.optimize(pts, pts)
        lr.optimize(pts, pts)
        np.testing.assert_allclose(lo.transform(pts), lr.transform(pts), rtol=0, atol=1e-12)
        region = ours.MLFriends(pts, lo)
        region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=10)
        region.create_ellipsoid()
        assert region.inside(pts).all()
    # create_new for the three affine flavours agrees with the reference's clustering
    blob = np.vstack([np.random.normal(0.3, 0.02, size=(200, 4)), np.random.normal(0.7, 0.02, size=(200, 4))])
    for name in ("AffineLayer", "LocalAffineLayer", "MaxPrincipleGapAffineLayer", "ScalingLayer"):
        res = []
        for mod in (ours, ref):
            layer = getattr(mod, name)()
            layer.optimize(blob, blob)
            region = mod.MLFriends(blob, layer)
            r, f = region.compute_enlargement(nbootstraps=10, rng=np.random.RandomState(1))
            nxt = layer.create_new(blob, r)
            res.append((nxt.nclusters, nxt.clusterids.copy(), nxt.logvolscale))
        assert res[0][0] == res[1][0] and (res[0][1] == res[1][1]).all(), name
        assert res[0][2] == pytest.approx(res[1][2], rel=1e-9), name
