"""Randomised differential test: many random shapes, radii, layers and launch sizes through every
membership / scan variant (fp32 and fp64 filters, warp-independent and block-synchronous
kernels, ordered and any-order), all against the CPU oracle.  Seeds are fixed, so a failure is
reproducible from the printed case number."""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ultranest_b200 import _native
    return _native.get_engine()


def _cloud(rng, n, d, kind):
    z = rng.normal(size=(n, d))
    if kind == 0:      # uniform ball (whitened live points)
        z /= np.sqrt((z**2).sum(axis=1, keepdims=True))
        z *= rng.uniform(size=(n, 1))**(1.0 / d)
    elif kind == 1:    # two separated blobs
        z *= 0.2
        z[: n // 2] += 1.5
    elif kind == 2:    # anisotropic, offset from the origin (large norms vs radius)
        z *= rng.uniform(0.05, 2.0, size=(1, d))
        z += 3.0
    else:              # lattice-like with exact duplicates and ties
        z = np.round(z * 2) / 2
    return np.ascontiguousarray(z)


@pytest.mark.parametrize("case", range(60))
def test_random_membership_case(eng, case):
    from ultranest_b200 import _native
    rng = np.random.RandomState(1000 + case)
    d = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 20, 21, 24, 31, 32, 33, 40, 64]))
    n = int(rng.choice([1, 2, 63, 64, 65, 200, 700, 1500, 4000]))
    m = int(rng.choice([1, 33, 500, 3000, 17000, 40000]))
    if d > 32:
        m = min(m, 3000)
    kind = int(rng.randint(4))
    a = _cloud(rng, n, d, kind)
    if rng.rand() < 0.5:   # proposals near live points: many hits, early exits, refills
        b = a[rng.randint(n, size=m)] + rng.normal(size=(m, d)) * 0.3 * a.std()
    else:                  # proposals from the same cloud: mixed
        b = _cloud(rng, m, d, kind) * rng.uniform(0.8, 1.3)
    sub = ((a[None, :min(n, 100), :] - b[:min(m, 100), None, :])**2).sum(axis=2).min(axis=1)
    r2 = float(np.quantile(sub, rng.choice([0.05, 0.5, 0.9]))) * rng.choice([1.0, 1.0, 3.0])
    r2 = max(r2, 1e-12)
    want_idx = cport.find_nearby(a, b, r2)
    want = want_idx >= 0
    eng.region_sync_live(a)
    eng.region_set_radius(r2)
    for flag in (1, 0):
        eng.set_option(_native.OPT_FILTER_FP32, flag)
        try:
            got = eng.region_has_neighbour(b)
            got2 = eng.has_neighbour(a, b, r2)
        finally:
            eng.set_option(_native.OPT_FILTER_FP32, 1)
        assert (got == want).all(), (case, d, n, m, kind, flag, int((got != want).sum()))
        assert (got2 == want).all(), (case, "stateless", flag)
    assert (eng.region_find_nearby(b) == want_idx).all(), (case, "find")
    if m <= 3000:
        assert (eng.region_count_nearby(b) == cport.count_nearby(a, b, r2)).all(), (case, "count")


@pytest.mark.parametrize("case", range(20))
def test_random_region_inside_case(eng, case):
    """Fused inside(): random live sets, learned layers, mutated rows, random batch sizes."""
    from ultranest_b200 import mlfriends as mm
    rng = np.random.RandomState(5000 + case)
    d = int(rng.choice([1, 2, 3, 5, 8, 13, 20, 27, 32, 36]))
    n = int(rng.choice([4 * d + 30, 100 + 3 * d, 400, 1000]))   # enough points for non-singular bootstraps
    z = _cloud(rng, n, d, 0)
    A = rng.normal(size=(d, d)) * 0.3 + np.eye(d)
    u = 0.5 + 0.04 * z @ A.T
    u = np.clip(u, 1e-6, 1 - 1e-6)
    layer = (mm.ScalingLayer if (d == 1 or rng.rand() < 0.25) else mm.AffineLayer)()
    layer.optimize(u, u)
    region = mm.MLFriends(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(
        nbootstraps=8, rng=np.random.RandomState(case))
    region.create_ellipsoid()
    if isinstance(layer, mm.AffineLayer):
        xf = lambda p: cport.transform_affine(p, layer.ctr, layer.T)   # noqa: E731
    else:
        xf = lambda p: cport.transform_scaling(p, layer.mean, layer.std)   # noqa: E731
    for step in range(3):
        m = int(rng.choice([1, 100, 5000, 30000]))
        pts = u[rng.randint(n, size=m)] + rng.normal(size=(m, d)) * 0.02 * rng.choice([0.2, 1.0, 3.0])
        want = cport.region_inside(pts, region.unormed, xf, region.maxradiussq,
                                   region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
        assert (region.inside(pts) == want).all(), (case, step, d, n, m)
        # the integrator's in-place patching between calls
        worst = int(rng.randint(n))
        unew = np.clip(u[int(rng.randint(n))] + rng.normal(size=d) * 1e-3, 1e-6, 1 - 1e-6)
        region.u[worst] = unew
        region.unormed[worst] = region.transformLayer.transform(unew)
        region.ellipsoid_center = np.mean(region.u, axis=0)
