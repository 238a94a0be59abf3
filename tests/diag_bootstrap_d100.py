"""Diagnostic: per-round bootstrap results at (N=4000, d=100) against the oracle."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import cport
from ultranest_b200 import mlfriends as m
from ultranest_b200 import _native

for n, d in ((4000, 100), (1000, 100), (4000, 50), (4000, 36)):
    u = bench.make_live(n, d, seed=1)
    layer = m.AffineLayer(); layer.optimize(u, u)
    region = m.MLFriends(u, layer)
    R = 3
    rng = np.random.RandomState(2)
    sel = np.array([cport.draw_selection(rng, n) for _ in range(R)])
    ctrs = np.zeros((R, d)); inv = np.zeros((R, d, d))
    want_d = np.zeros(R); want_f = np.zeros(R)
    for r in range(R):
        ctr, cov = cport.bounding_ellipsoid(u[sel[r]])
        ctrs[r] = ctr; inv[r] = np.linalg.inv(cov)
        want_d[r] = cport.maxradiussq_selected(region.unormed, sel[r])
        want_f[r] = cport.enlargement_f(u, sel[r], ctr, inv[r])
    eng = _native.get_engine()
    got_d, got_f = eng.region_bootstrap(region.unormed, sel, u=u, ctrs=ctrs, invcovs=inv)
    print(n, d, "maxd equal", (got_d == want_d).tolist(), "f equal", (got_f == want_f).tolist())
    print("   maxd", got_d, want_d)
    print("   f", got_f, want_f, (got_f - want_f) / want_f)
    eng.set_option(_native.OPT_EXACT_ONLY, 1)
    got_d2, got_f2 = eng.region_bootstrap(region.unormed, sel, u=u, ctrs=ctrs, invcovs=inv)
    eng.set_option(_native.OPT_EXACT_ONLY, 0)
    print("   exact-only: maxd equal", (got_d2 == want_d).tolist(), "f equal", (got_f2 == want_f).tolist())
