"""CPU (oracle-backed stub engine): the fallback logic around the transform tolerance -- a pair
exactly on the radius is reported and `inside()` / `inside_and_loglike()` / the fused refill then
decide like the reference (np.dot transform), while ordinary calls never leave the fused path."""
import numpy as np
import pytest

import oracle
import tolerance_cases as tc

pytestmark = pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built")


@pytest.fixture()
def stub_engine(monkeypatch):
    from oracle_engine import OracleEngine
    from ultranest_b200 import _native
    eng = OracleEngine()
    monkeypatch.setattr(_native, "_engine", eng)
    monkeypatch.setattr(_native, "get_engine", lambda: eng)
    return eng


def test_tolerance_is_tiny_and_zero_for_exact_layers(stub_engine):
    from ultranest_b200 import mlfriends as ml
    region = tc.build(ml)
    tau = region._transform_tolerance()
    assert 0 < tau < 1e-9 * region.maxradiussq
    assert region._transform_tolerance() == tau          # cached
    u = region.u
    layer = ml.ScalingLayer()
    layer.optimize(u, u)
    reg2 = ml.MLFriends(u, layer)
    reg2.maxradiussq, reg2.enlarge = 0.1, 1.5
    reg2.create_ellipsoid()
    assert reg2._transform_tolerance() == 0.0


def test_edges_decide_like_the_reference(stub_engine):
    from ultranest_b200 import mlfriends as ml
    region = tc.build(ml)
    ncases, raw_diff, fallbacks = tc.check_edges(region, stub_engine, count=30)
    assert fallbacks == ncases
    print("edge cases %d, raw defined-order decision differs in %d" % (ncases, raw_diff))


def test_ordinary_calls_stay_fused(stub_engine):
    from ultranest_b200 import mlfriends as ml
    from ultranest_b200.likelihoods import GaussianLogLike
    region = tc.build(ml)
    rng = np.random.RandomState(3)
    pts = region.u[rng.randint(len(region.u), size=3000)] + rng.normal(size=(3000, region.u.shape[1])) * 0.05
    pts = pts[np.logical_and(pts > 0, pts < 1).all(axis=1)]
    calls = stub_engine.calls
    mask = region.inside(pts)
    assert stub_engine.uncertain() == 0 and stub_engine.calls == calls + 1
    assert (mask == tc.reference_inside(region, pts)).all()
    # a batch that contains ONE edge row: the whole call is re-decided, every row like the reference
    w, r2 = tc.edge_cases(region, 1)[0]
    saved = region.maxradiussq
    region.maxradiussq = r2
    try:
        batch = np.vstack([pts[:500], w.reshape(1, -1), pts[500:900]])
        want = tc.reference_inside(region, batch)
        assert (region.inside(batch) == want).all()
        loglike = GaussianLogLike(0.5, 0.1)
        m2, like = region.inside_and_loglike(batch, loglike)
        assert (m2 == want).all()
        assert (like[want] == loglike(batch[want])).all() and np.isneginf(like[~want]).all()
    finally:
        region.maxradiussq = saved
