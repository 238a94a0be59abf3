"""Repeats the integrator's per-iteration pattern a few times (host-side latency is noisy)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ultranest_b200 import mlfriends as m, _native

u = bench.make_live(4000, 20, seed=1)
layer = m.AffineLayer(); layer.optimize(u, u)
region = m.MLFriends(u, layer)
region.maxradiussq, region.enlarge = region.compute_enlargement(30, rng=np.random.RandomState(2))
region.create_ellipsoid()
u0 = region.u.copy()
eng = _native.get_engine()
for rep in range(4):
    t0 = time.perf_counter(); nit = 300
    tb = 0.0
    for i in range(nit):
        worst = i % len(u0)
        unew = u0[(i * 7 + 3) % len(u0)] + 1e-4
        region.u[worst] = unew
        region.unormed[worst] = region.transformLayer.transform(unew)
        region.ellipsoid_center = np.mean(region.u, axis=0)
        t1 = time.perf_counter()
        ok = region.inside(region.u)
        tb += time.perf_counter() - t1
    print("rep %d: %.0f us/iteration (inside() alone %.0f us)" % (rep, (time.perf_counter() - t0) / nit * 1e6, tb / nit * 1e6))
# inside() on unchanged state: pure call overhead + kernels
t0 = time.perf_counter()
for i in range(300):
    ok = region.inside(region.u)
print("unchanged state: %.0f us/call" % ((time.perf_counter() - t0) / 300 * 1e6))
