"""Shared scenario of the transform-tolerance tests (CPU stub and GPU): proposals whose nearest
live point sits EXACTLY on the radius when the layer transform is the reference's np.dot -- the one
place where the device's defined summation order could turn a decision around."""
import numpy as np

from oracle import cport


def seq_dist(a, b):
    D = 0.0
    for k in range(len(a)):
        diff = a[k] - b[k]
        D = D + diff * diff
    return D


def build(ml, n=600, d=9, seed=4):
    import bench
    u = 0.5 + (bench.make_live(n, d, seed=seed) - 0.5) * 3.0
    layer = ml.AffineLayer()
    layer.optimize(u, u)
    region = ml.MLFriends(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=5, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    return region


def reference_inside(region, pts):
    """The reference's decision: einsum ellipsoid, np.dot transform, exact first-neighbour scan."""
    lay = region.transformLayer
    return cport.region_inside(pts, region.unormed, lambda p: np.dot(p - lay.ctr, lay.T),
                               region.maxradiussq, region.ellipsoid_center, region.ellipsoid_invcov,
                               region.enlarge)


def defined_order_inside(region, pts):
    """What the device computes without the tolerance mechanism (defined-order transform)."""
    lay = region.transformLayer
    return cport.region_inside(pts, region.unormed, lambda p: cport.transform_affine(p, lay.ctr, lay.T),
                               region.maxradiussq, region.ellipsoid_center, region.ellipsoid_invcov,
                               region.enlarge)


def edge_cases(region, count, seed=7):
    """(proposal row, radius^2) pairs: radius^2 = the reference's own distance to the nearest live
    point (a hit by `<=`) and the next double below it (a miss)."""
    rng = np.random.RandomState(seed)
    lay = region.transformLayer
    d = region.u.shape[1]
    out = []
    while len(out) < 2 * count:
        i = int(rng.randint(len(region.u)))
        w = region.u[i] + rng.normal(size=d) * 0.02
        if not (np.logical_and(w > 0, w < 1).all() and region.inside_ellipsoid(w.reshape(1, -1))[0]):
            continue
        t = np.dot(w - lay.ctr, lay.T)
        dist = np.array([seq_dist(region.unormed[j], t) for j in range(len(region.unormed))])
        D = float(dist.min())
        out.append((w, D))
        out.append((w, float(np.nextafter(D, 0))))
    return out


def check_edges(region, eng, count=40):
    """Returns (cases, disagreements of the raw defined-order decision, fallbacks taken)."""
    saved = region.maxradiussq
    raw_diff = fallbacks = 0
    cases = edge_cases(region, count)
    try:
        for w, r2 in cases:
            region.maxradiussq = r2
            row = w.reshape(1, -1)
            want = reference_inside(region, row)[0]
            got = region.inside(row)[0]
            assert got == want, (r2, got, want)
            raw_diff += int(defined_order_inside(region, row)[0] != want)
            # the fused call itself must have reported the pair
            region._bind().region_inside(row)
            fallbacks += int(eng.uncertain() > 0)
            assert eng.uncertain() > 0, "a pair exactly on the radius was not reported"
    finally:
        region.maxradiussq = saved
    return len(cases), raw_diff, fallbacks
