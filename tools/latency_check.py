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

# ---- _refill_samples: the reference's staged chain (region.sample -> transform -> loglike ->
# logl > Lmin, integrator.py:1773-1805) on this package's classes vs the fused device pipeline
import types
from ultranest_b200 import refill
from ultranest_b200.likelihoods import GaussianLogLike
from ultranest_b200.transforms import ScaleShiftTransform

region.u[...] = u0
region.unormed[...] = region.transformLayer.transform(u0)
region.ellipsoid_center = np.mean(region.u, axis=0)
region.current_sampling_method = region.sample_from_wrapping_ellipsoid
transform = ScaleShiftTransform(-1.0, 3.0)
loglike = GaussianLogLike(1.0, 0.4)
Lmin = float(np.median(loglike(transform(region.u))))
s = types.SimpleNamespace(region=region, tregion=None, loglike=loglike, transform=transform,
                          draw_multiple=True, x_dim=20, num_params=20,
                          sampling_slow_warned=False, ncall_region=0)


def staged(ndraw):
    uu = region.sample(nsamples=ndraw)
    v = transform(uu)
    logl = loglike(v) if len(uu) else np.empty(0)
    acc = logl > Lmin
    return uu[acc, :], v[acc, :], logl[acc]


for ndraw in (128, 4096, 65536):
    for name, fn in (("staged", lambda: staged(ndraw)),
                     ("fused", lambda: refill.refill_samples(s, Lmin, ndraw, 1))):
        np.random.seed(5)
        fn()
        t0 = time.perf_counter()
        reps = 200 if ndraw <= 4096 else 30
        for _ in range(reps):
            out = fn()
        dt = (time.perf_counter() - t0) / reps
        print("refill ndraw=%6d %-6s: %8.0f us/call  (%d accepted)" % (ndraw, name, dt * 1e6, len(out[0])))
