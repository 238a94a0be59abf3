"""GPU: the transform tolerance through the real kernels.  A pair exactly on the
radius (under the reference's np.dot transform) must be counted by the membership kernels (fp32 and
fp64 filter, block and warp kernels), and `inside()` must then decide like the reference; ordinary
large batches must never take the fallback."""
import numpy as np
import pytest

import tolerance_cases as tc

pytestmark = pytest.mark.gpu


@pytest.fixture()
def eng():
    from ultranest_b200 import _native
    e = _native.get_engine()
    yield e
    e.set_option(_native.OPT_FILTER_FP32, 1)
    e.set_option(_native.OPT_BLOCK_KERNEL, 0)
    e.set_option(_native.OPT_SURE_LEVEL, 1)


@pytest.mark.parametrize("fp32,block", [(1, 0), (1, 1), (0, 0)])
def test_edges_decide_like_the_reference(eng, fp32, block):
    from ultranest_b200 import _native
    from ultranest_b200 import mlfriends as ml
    eng.set_option(_native.OPT_FILTER_FP32, fp32)
    eng.set_option(_native.OPT_BLOCK_KERNEL, block)
    region = tc.build(ml)
    ncases, raw_diff, fallbacks = tc.check_edges(region, eng, count=40)
    assert fallbacks == ncases
    print("fp32=%d block=%d: %d edge cases, raw defined-order decision differs in %d"
          % (fp32, block, ncases, raw_diff))


def test_edge_row_inside_a_large_batch_and_fused_calls(eng):
    from ultranest_b200 import mlfriends as ml
    from ultranest_b200.likelihoods import GaussianLogLike
    region = tc.build(ml, n=2000, d=12)
    rng = np.random.RandomState(3)
    d = region.u.shape[1]
    pts = region.u[rng.randint(len(region.u), size=60000)] + rng.normal(size=(60000, d)) * 0.05
    pts = pts[np.logical_and(pts > 0, pts < 1).all(axis=1)]
    mask = region.inside(pts)
    assert eng.uncertain() == 0, "an ordinary batch must not leave the fused path"
    sel = rng.choice(len(pts), 4000, replace=False)
    assert (mask[sel] == tc.reference_inside(region, pts[sel])).all()
    w, r2 = tc.edge_cases(region, 1)[0]
    saved = region.maxradiussq
    region.maxradiussq = r2
    try:
        batch = np.vstack([pts[:30000], w.reshape(1, -1), pts[30000:]])
        got = region.inside(batch)
        assert got[30000] == tc.reference_inside(region, w.reshape(1, -1))[0]
        want = tc.reference_inside(region, batch[29000:31000])
        assert (got[29000:31000] == want).all()
        loglike = GaussianLogLike(0.5, 0.1)
        m2, like = region.inside_and_loglike(batch, loglike)
        assert (m2 == got).all()
        assert (like[m2] == loglike(batch[m2])).all() and np.isneginf(like[~m2]).all()
    finally:
        region.maxradiussq = saved
