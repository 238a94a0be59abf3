#!/usr/bin/env python
"""Small cases of every kernel family, sized for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tests/sanitize_cases.py
    compute-sanitizer --tool racecheck python tests/sanitize_cases.py

Each case is also checked against the oracle, so a sanitizer-clean run is a parity run too.
Covers: block membership kernel with refill + drain compaction + cooperative drain (forced with
UNB_OPT_BLOCK_KERNEL), the warp-independent kernel, the fp64-filter kernel, the ordered scans
(FIND/COUNT/SUBTRACT/MIN), the prep kernels (register d<=32, tile d=50), the bootstrap, the fused
inside+loglike call and the device proposal generator.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cport  # noqa: E402
from ultranest_b200 import _native  # noqa: E402
from ultranest_b200 import mlfriends as m  # noqa: E402
from ultranest_b200.likelihoods import GaussianLogLike  # noqa: E402
import bench  # noqa: E402


def region_for(n, d, cls=None, nboot=4):
    u = bench.make_live(n, d, seed=1)
    layer = m.AffineLayer()
    layer.optimize(u, u)
    region = (cls or m.MLFriends)(u, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=nboot, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    return region


def oracle_inside(region, pts):
    layer = region.transformLayer
    return cport.region_inside(pts, region.unormed, lambda p: cport.transform_affine(p, layer.ctr, layer.T),
                               region.maxradiussq, region.ellipsoid_center, region.ellipsoid_invcov,
                               region.enlarge)


def main():
    eng = _native.get_engine()
    rng = np.random.RandomState(3)
    done = []

    # membership: accepting + rejecting mix so that slots refill, drain, and the tail turns cooperative
    region = region_for(700, 20)
    pts = np.vstack([bench.make_candidates(region, 6000, 3),
                     rng.uniform(0.3, 0.7, size=(3000, 20))])
    want = oracle_inside(region, pts)
    for block, fp32, sure, coop in ((1, 1, 1, 24), (1, 1, 0, 0), (0, 1, 1, 24), (1, 0, 1, 24)):
        eng.set_option(_native.OPT_BLOCK_KERNEL, block)
        eng.set_option(_native.OPT_FILTER_FP32, fp32)
        eng.set_option(_native.OPT_SURE_LEVEL, sure)
        eng.set_option(_native.OPT_COOP_MAX, coop)
        got = region.inside(pts)
        assert (got == want).all(), ("membership", block, fp32, sure, coop)
        done.append("inside block=%d fp32=%d sure=%d coop=%d" % (block, fp32, sure, coop))
    eng.set_option(_native.OPT_BLOCK_KERNEL, 0)
    eng.set_option(_native.OPT_FILTER_FP32, 1)
    eng.set_option(_native.OPT_SURE_LEVEL, 1)
    eng.set_option(_native.OPT_COOP_MAX, 24)
    mask, like = region.inside_and_loglike(pts, GaussianLogLike(0.5, 0.05))
    assert (mask == want).all()
    done.append("inside_and_loglike")

    # ordered scans, d in the register kernel and in the shared-memory kernel
    for n, nb, d in ((300, 900, 7), (500, 600, 20), (200, 300, 50)):
        z = rng.normal(size=(n, d))
        a = z / np.sqrt((z**2).sum(axis=1, keepdims=True)) * rng.uniform(size=(n, 1))**(1.0 / d)
        z = rng.normal(size=(nb, d))
        b = z / np.sqrt((z**2).sum(axis=1, keepdims=True)) * rng.uniform(size=(nb, 1))**(1.0 / d) * 1.1
        r2 = float(np.quantile(((a[None, :100] - b[:100, None])**2).sum(axis=2).min(axis=1), 0.5))
        assert (eng.find_nearby(a, b, r2) == cport.find_nearby(a, b, r2)).all()
        assert (eng.count_nearby(a, b, r2) == cport.count_nearby(a, b, r2)).all()
        assert (eng.has_neighbour(a, b, r2) == (cport.find_nearby(a, b, r2) >= 0)).all()
        assert eng.compute_maxradiussq(a, b) == cport.maxradiussq(a, b)
        assert (eng.subtract_nearby(a, r2) == cport.subtract_nearby(a, r2)).all()
        done.append("scans n=%d nb=%d d=%d" % (n, nb, d))

    # tile prep kernel (d = 50) through an ellipsoid-only region, and the bootstrap
    reg50 = region_for(400, 50, cls=m.RobustEllipsoidRegion)
    p50 = bench.make_candidates(reg50, 2000, 5)
    p50[::3] += 0.01
    got = reg50.inside(p50)
    want50 = cport.inside_ellipsoid(p50, reg50.ellipsoid_center, reg50.ellipsoid_invcov, reg50.enlarge)
    assert (got == want50).all()
    done.append("ellipsoid-only d=50")
    r, f = region.compute_enlargement(nbootstraps=6, rng=np.random.RandomState(2))
    assert (r, f) == cport.compute_enlargement(region.u, region.unormed, 6, np.random.RandomState(2))
    done.append("bootstrap")

    if hasattr(region, "sample_device"):
        got = region.sample_device(4096, seed=11)
        assert oracle_inside(region, got).all()
        done.append("device proposals")
    print("sanitize_cases ok: " + "; ".join(done))


if __name__ == "__main__":
    main()
