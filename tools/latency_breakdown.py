"""Where the per-iteration `region.inside(active_u)` call of the integrator (integrator.py:1855)
spends its time: host-side mirror sync, copies, kernels."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ultranest_b200 import mlfriends as m, _native  # noqa: E402

u = bench.make_live(4000, 20, seed=1)
layer = m.AffineLayer(); layer.optimize(u, u)
region = m.MLFriends(u, layer)
region.maxradiussq, region.enlarge = region.compute_enlargement(30, rng=np.random.RandomState(2))
region.create_ellipsoid()
eng = _native.get_engine()
region.inside(region.u)


def t(fn, n=300):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


kind, shift, mat = region.transformLayer._device_params(20)
print("region.inside(region.u) total          %7.1f us" % t(lambda: region.inside(region.u)))
print("  _bind() total                        %7.1f us" % t(lambda: region._bind()))
print("    region_sync_live (unchanged)       %7.1f us" % t(lambda: eng.region_sync_live(region.unormed)))
print("    region_set_radius                  %7.1f us" % t(lambda: eng.region_set_radius(region.maxradiussq)))
print("    region_set_layer                   %7.1f us" % t(lambda: eng.region_set_layer(kind, shift, mat, 20)))
print("    region_set_ellipsoid               %7.1f us" % t(lambda: eng.region_set_ellipsoid(region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)))
print("  eng.region_inside (pageable rows)    %7.1f us" % t(lambda: eng.region_inside(region.u)))
pin = torch.empty(region.u.shape, dtype=torch.float64).pin_memory(); pin.numpy()[...] = region.u
print("  eng.region_inside (pinned rows)      %7.1f us" % t(lambda: eng.region_inside(pin.numpy())))
dev = torch.from_numpy(region.u).cuda(); mk = torch.empty(4000, dtype=torch.uint8, device="cuda")


def dev_call():
    eng.call("unb_region_inside_dev", dev.data_ptr(), 4000, mk.data_ptr(), None)
    eng.synchronize()


print("  unb_region_inside_dev + sync         %7.1f us" % t(dev_call))
from ultranest_b200 import _native as N
eng.set_option(N.OPT_BLOCK_KERNEL, 1)
print("  ... block kernel                     %7.1f us" % t(dev_call))
eng.set_option(N.OPT_BLOCK_KERNEL, 0)
# one row patched per call (the integrator's pattern)
u0 = region.u.copy()
state = {"i": 0}


def patched():
    i = state["i"] = (state["i"] + 1) % 4000
    unew = u0[(i * 7 + 3) % 4000] + 1e-4
    region.u[i] = unew
    region.unormed[i] = region.transformLayer.transform(unew)
    region.ellipsoid_center = np.mean(region.u, axis=0)
    return region.inside(region.u)


print("patched row + new centre + inside()    %7.1f us" % t(patched))
print("  of which host numpy (transform+mean) %7.1f us" % t(lambda: (region.transformLayer.transform(u0[5]), np.mean(region.u, axis=0))))
