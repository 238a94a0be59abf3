"""GPU tier of the population step-sampler helpers: the CUDA path, through
the C ABI, against the golden vectors of the reference, against the C oracle on random cases,
and end to end: the reference's vectorised slice samplers driven by the device helpers return
the same points as the unmodified reference with NumPy callables."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import oracle  # noqa: E402
from oracle import stepport  # noqa: E402
import stepfuncs_cases as cases  # noqa: E402
import stepfuncs_checks as checks  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sf():
    from ultranest_b200 import stepfuncs
    return stepfuncs


def _device_callables(centre=0.5, sigma=0.1, lo=None, hi=None):
    from ultranest_b200.likelihoods import GaussianLogLike
    from ultranest_b200.transforms import IdentityTransform, ScaleShiftTransform
    xf = IdentityTransform() if lo is None else ScaleShiftTransform(lo, hi)
    return xf, GaussianLogLike(centre, sigma)


@pytest.mark.parametrize("check", [c for c in checks.ALL_CHECKS if c is not checks.check_evolve],
                         ids=lambda f: f.__name__)
def test_device_matches_golden(sf, check):
    check(sf, checks.golden())


def test_evolve_fused_and_staged_match_golden(sf):
    xf, ll = _device_callables(0.5, 0.1)
    checks.check_evolve(sf, checks.golden(), xf, ll)                    # one fused kernel
    checks.check_evolve(sf, checks.golden(), cases.identity, cases.gauss_loglike(0.5, 0.1))   # host callables


@pytest.mark.parametrize("seed", range(8))
def test_differential_vs_oracle(sf, seed):
    rng = np.random.RandomState(200 + seed)
    n, d = int(rng.choice([1, 31, 32, 33, 500, 1024, 1025, 5000])), int(rng.randint(1, 40))
    u = cases.cube_case(seed, n, d)
    assert (sf.within_unit_cube(u) == stepport.within_unit_cube(u)).all()
    a, b = cases.evolve_update_case(seed, n), cases.evolve_update_case(seed, n)
    sa, sb = np.zeros(n, dtype=bool), np.zeros(n, dtype=bool)
    pa = sf.evolve_prepare(a["searching_left"], a["searching_right"])
    pb = stepport.evolve_prepare(b["searching_left"], b["searching_right"])
    assert (pa[0] == pb[0]).all() and (pa[1] == pb[1]).all()
    sf.evolve_update(a["acceptable"], a["Lnew"], a["Lmin"], pa[0], pa[1], a["currentt"], a["current_left"],
                     a["current_right"], a["searching_left"], a["searching_right"], sa)
    stepport.evolve_update(b["acceptable"], b["Lnew"], b["Lmin"], pb[0], pb[1], b["currentt"], b["current_left"],
                           b["current_right"], b["searching_left"], b["searching_right"], sb)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    assert (sa == sb).all()
    a, b = cases.step_back_case(seed, n, 70), cases.step_back_case(seed, n, 70)
    sf.step_back(a["Lmin"], a["allL"], a["generation"], a["currentt"])
    stepport.step_back(b["Lmin"], b["allL"], b["generation"], b["currentt"])
    for k in ("allL", "generation", "currentt"):
        np.testing.assert_array_equal(a[k], b[k])
    # fused evolve with a scale-shift prior against the oracle's NumPy chain
    lo, hi = -np.arange(1, d + 1.0), np.arange(2, d + 2.0)
    xf, ll = _device_callables(0.3, 2.5, lo, hi)
    sa, sb = cases.evolve_state(seed, n, d), cases.evolve_state(seed, n, d)
    np.random.seed(seed)
    ra = sf.evolve(xf, ll, -30.0 * d, **sa)
    np.random.seed(seed)
    rb = stepport.evolve(lambda x: x * (hi - lo) + lo, cases.gauss_loglike(0.3, 2.5), -30.0 * d, **sb)
    for k in sa:
        np.testing.assert_array_equal(sa[k], sb[k])
    for x, y in zip(ra[1], rb[1]):
        np.testing.assert_array_equal(x, y)
    assert ra[2] == rb[2]


@pytest.mark.parametrize("popsize,d,shrink", [(64, 3, 1.0), (1000, 5, 1.0), (1500, 20, 1.5), (4097, 8, 1.0)])
def test_device_resident_slice_loop_vs_oracle(popsize, d, shrink):
    """unb_popslice_begin/iterate/end == the oracle's restatement of popstepsampler.py:940-965,
    pass by pass (running count, discarded count) and in the final state."""
    from ultranest_b200 import popstepsampler as pp
    c = cases.slice_sampler_case(popsize + d, popsize, d, 40, shrink=shrink)
    xf, ll = _device_callables(0.5, c["sigma"])
    loop = pp.SliceLoop(xf, ll, d)
    loop.begin(c["allu"], c["allL"], c["v"], c["tleft"], c["tright"], c["Lmin"], shrink)
    allu, allL = c["allu"].copy(), c["allL"].copy()
    allp = np.full_like(allu, np.nan)
    tleft, tright = c["tleft"].copy(), c["tright"].copy()
    tlw, trw = tleft.copy(), tright.copy()
    worker = np.arange(popsize, dtype=np.int64)
    status = np.zeros(popsize, dtype=np.int64)
    host_ll = cases.gauss_loglike(0.5, c["sigma"])
    for it in range(len(c["draws"])):
        n_running, disc = loop.iterate(c["draws"][it])
        tlw, trw, want_disc = stepport.popslice_iteration(
            c["draws"][it], tlw, trw, tleft, tright, worker, status, allu, allL, allp, c["v"],
            cases.identity, host_ll, c["Lmin"], shrink)
        assert n_running == int((status == 0).sum())
        assert disc == want_disc
        if n_running == 0:
            break
    got = loop.end()
    for x, y in zip(got, (allu, allp, allL, tleft, tright, status)):
        np.testing.assert_array_equal(x, y)
    assert (status == 1).mean() > 0.5


# ---- end to end: the reference's samplers on top of the device helpers --------------------------

def _ref_modules():
    if not oracle.reference_available():
        pytest.skip("oracle/_ref not built")
    oracle.reference()
    import ultranest.popstepsampler as rp
    import ultranest.stepfuncs as rs
    return rs, rp


def _harvest(sampler, region, us, Ls, transform, loglike, nsamples, seed, lmin_of=np.median):
    np.random.seed(seed)
    Lmin = float(lmin_of(Ls))
    out, calls = [], 0
    while len(out) < nsamples and calls < 4000:
        u, p, L, nc = sampler.__next__(region, Lmin, us, Ls, transform, loglike)
        calls += 1
        if u is not None:
            out.append((u, p, L, nc))
    assert len(out) == nsamples
    return out


def _problem(d, nlive, seed):
    from ultranest_b200 import mlfriends as ml
    rng = np.random.RandomState(seed)
    us = rng.uniform(0.35, 0.65, size=(nlive, d))
    layer = ml.AffineLayer()
    layer.optimize(us, us)
    region = ml.MLFriends(us, layer)
    region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=4, rng=np.random.RandomState(2))
    region.create_ellipsoid()
    return us, region


def test_population_slice_sampler_run_is_the_reference_run(sf):
    rs, rp = _ref_modules()
    d = 4
    us, region = _problem(d, 300, 8)
    host_ll = cases.gauss_loglike(0.5, 0.08)
    Ls = host_ll(us)
    make = lambda: rp.PopulationSliceSampler(popsize=40, nsteps=6, generate_direction=rs.generate_mixture_random_direction)  # noqa: E731
    want = _harvest(make(), region, us, Ls, cases.identity, host_ll, 25, 21)
    xf, ll = _device_callables(0.5, 0.08)
    undo = sf.install()
    try:
        assert rp.evolve is sf.evolve
        got = _harvest(make(), region, us, Ls, xf, ll, 25, 21)          # fused evolve
        got_staged = _harvest(make(), region, us, Ls, cases.identity, host_ll, 25, 21)   # host callables
    finally:
        sf.uninstall(undo)
    assert rp.evolve is not sf.evolve
    for run in (got, got_staged):
        for (u, p, L, nc), (u2, p2, L2, nc2) in zip(run, want):
            np.testing.assert_array_equal(u, u2)
            np.testing.assert_array_equal(p, p2)
            assert L == L2 and nc == nc2


@pytest.mark.parametrize("shrink,scale", [(1.0, 1.0), (1.3, 0.4)])
def test_simple_slice_sampler_device_loop_is_the_reference_run(shrink, scale):
    from ultranest_b200 import popstepsampler as pp
    rs, rp = _ref_modules()
    d = 6
    us, region = _problem(d, 400, 9)
    lo, hi = np.full(d, -2.0), np.full(d, 3.0)
    host_xf = lambda x: x * (hi - lo) + lo   # noqa: E731
    host_ll = cases.gauss_loglike(0.5, 0.4)
    Ls = host_ll(host_xf(us))
    kw = dict(popsize=96, nsteps=5, generate_direction=rs.generate_region_random_direction,
              shrink_factor=shrink, scale=scale,
              slice_limit=rp.slice_limit_to_scale if scale != 1.0 else rp.slice_limit_to_unitcube)
    want = _harvest(rp.PopulationSimpleSliceSampler(**kw), region, us, Ls, host_xf, host_ll, 200, 31, np.min)
    xf, ll = _device_callables(0.5, 0.4, lo, hi)
    sampler = rp.PopulationSimpleSliceSampler(**kw)
    stats = pp.attach(sampler)
    got = _harvest(sampler, region, us, Ls, xf, ll, 200, 31, np.min)
    assert stats['fused_calls'] > 0 and stats['delegated_calls'] == 0
    for (u, p, L, nc), (u2, p2, L2, nc2) in zip(got, want):
        np.testing.assert_array_equal(u, u2)
        np.testing.assert_array_equal(p, p2)
        assert L == L2 and nc == nc2
    # host callables: delegated to the reference method, still the same run
    sampler = rp.PopulationSimpleSliceSampler(**kw)
    stats = pp.attach(sampler)
    got = _harvest(sampler, region, us, Ls, host_xf, host_ll, 50, 31, np.min)
    assert stats['fused_calls'] == 0
    for (u, p, L, nc), (u2, p2, L2, nc2) in zip(got, want):
        np.testing.assert_array_equal(u, u2)


def test_evolve_in_chunks(sf):
    """Large populations pass through unb_evolve in chunks of walkers; forced here with a tiny chunk."""
    from ultranest_b200 import _native
    xf, ll = _device_callables(0.5, 0.1)
    a, b = cases.evolve_state(9, 2000, 7), cases.evolve_state(9, 2000, 7)
    np.random.seed(3)
    ra = sf.evolve(xf, ll, -8.0, **a)
    eng = _native.get_engine()
    eng.set_option(_native.OPT_CHUNK_ROWS, 333)
    try:
        np.random.seed(3)
        rb = sf.evolve(xf, ll, -8.0, **b)
    finally:
        eng.set_option(_native.OPT_CHUNK_ROWS, 0)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    for x, y in zip(ra[1], rb[1]):
        np.testing.assert_array_equal(x, y)
    assert ra[2] == rb[2]


def test_argument_errors(sf):
    with pytest.raises(ValueError):
        sf.step_back(0.0, np.zeros((3, 4000)), np.zeros(3, dtype=np.int64), np.zeros(3))
    with pytest.raises(ValueError):
        sf.update_vectorised_slice_sampler(
            np.zeros(2), np.zeros(2), np.zeros(2), np.zeros(2), np.zeros((2, 1)), np.zeros((2, 1)),
            np.array([0, 5], dtype=np.int64), np.zeros(2, dtype=np.int64), 0.0, 1.0,
            np.zeros((2, 1)), np.zeros(2), np.zeros((2, 1)), 2)
    with pytest.raises(ValueError):
        sf.evolve_update(np.ones(2, dtype=bool), np.zeros(2), 0.0, np.ones(2, dtype=bool), np.ones(2, dtype=bool),
                         np.zeros(2, dtype=np.float32), np.zeros(2), np.zeros(2), np.zeros(2, dtype=bool),
                         np.zeros(2, dtype=bool), np.zeros(2, dtype=bool))
