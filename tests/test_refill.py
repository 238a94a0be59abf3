"""Fused ``_refill_samples`` (ultranest_b200/refill.py, ``unb_region_refill``; SURVEY 8-f rank 1,
integrator.py:1773-1837).

* GPU tier: the device pipeline's flags / likelihoods / counters against the stage-by-stage oracle
  chain (tests/oracle_engine.py) on the same proposals, for every region mode, with and without
  a prior transform and a transformed-space ellipsoid; and seeded ``ReactiveNestedSampler`` runs
  with the fused refill attached against the reference's own run.
* CPU tier: the same integrator runs with the kernels answered by the oracle engine -- pins the
  host logic of refill.py (RNG order, counters, method switching, delegation).
"""
import sys
import types

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built")

NDIM = 5
SIGMA = 0.01
LO, HI = -3.0, 5.0


def numpy_gauss(theta, center=0.5, sigma=SIGMA):
    ndim = theta.shape[1]
    return -0.5 * (((theta - center) / sigma)**2).sum(axis=1) - 0.5 * np.log(2 * np.pi * sigma**2) * ndim


def _run(loglike, transform, fused, nlive=200, max_ncalls=6000, seed=11, ndim=NDIM):
    from ultranest import ReactiveNestedSampler
    np.random.seed(seed)
    sampler = ReactiveNestedSampler(["p%d" % i for i in range(ndim)], loglike, transform=transform,
                                    log_dir=None, vectorized=True)
    stats = None
    if fused:
        from ultranest_b200 import refill
        stats = refill.attach(sampler)
    res = sampler.run(min_num_live_points=nlive, max_ncalls=max_ncalls, viz_callback=False,
                      show_status=False)
    return dict(logz=res["logz"], ncall=res["ncall"], niter=res["niter"],
                ncall_region=sampler.ncall_region, region=type(sampler.region).__module__,
                stats=stats)


def _restore_integrator():
    """Context: remember the integrator's bindings, restore them afterwards."""
    oracle.reference()
    import ultranest.integrator as integ
    import ultranest.mlfriends as refmod
    names = ("AffineLayer", "LocalAffineLayer", "MLFriends", "RobustEllipsoidRegion",
             "ScalingLayer", "WrappingEllipsoid", "find_nearby")
    saved = {n: getattr(integ, n) for n in names}
    fns = [fn for cls in vars(integ).values() if isinstance(cls, type)
           for fn in vars(cls).values() if getattr(fn, "__defaults__", None)]
    saved_defaults = [(fn, fn.__defaults__) for fn in fns]

    def restore():
        for n, v in saved.items():
            setattr(integ, n, v)
        for fn, d in saved_defaults:
            fn.__defaults__ = d
        sys.modules["ultranest.mlfriends"] = refmod
        sys.modules["ultranest"].mlfriends = refmod
    return restore


def _identity_case():
    from ultranest_b200.likelihoods import GaussianLogLike
    from ultranest_b200.transforms import IdentityTransform
    want = _run(numpy_gauss, lambda u: u, fused=False)
    assert want["region"] == "ultranest.mlfriends"
    import ultranest_b200
    ultranest_b200.install(force=True)
    got = _run(GaussianLogLike(0.5, SIGMA), IdentityTransform(), fused=True)
    return want, got


def _scaleshift_case():
    from ultranest_b200.likelihoods import GaussianLogLike
    from ultranest_b200.transforms import ScaleShiftTransform
    center, sigma = 1.25, 0.08

    def ref_transform(u):
        return u * (HI - LO) + LO

    want = _run(lambda t: numpy_gauss(t, center, sigma), ref_transform, fused=False)
    import ultranest_b200
    ultranest_b200.install(force=True)
    got = _run(GaussianLogLike(center, sigma), ScaleShiftTransform(LO, HI), fused=True)
    return want, got


def _check_same_run(want, got):
    assert got["region"] == "ultranest_b200.mlfriends"
    for key in ("niter", "ncall", "ncall_region"):
        assert got[key] == want[key], key
    assert abs(got["logz"] - want["logz"]) <= 1e-10 * abs(want["logz"])
    st = got["stats"]
    assert st["fused_calls"] > 50, st
    assert st["fused_calls"] > 5 * st["delegated_calls"], st


# ---- CPU tier: host logic of refill.py over the oracle engine -----------------------------------
@pytest.fixture()
def stub_engine(monkeypatch):
    from oracle_engine import OracleEngine
    from ultranest_b200 import _native
    eng = OracleEngine()
    monkeypatch.setattr(_native, "_engine", eng)
    monkeypatch.setattr(_native, "get_engine", lambda: eng)
    return eng


@needs_ref
@pytest.mark.parametrize("case", ["identity", "scaleshift"])
def test_fused_refill_host_logic_reproduces_reference_run(stub_engine, case):
    restore = _restore_integrator()
    try:
        want, got = _identity_case() if case == "identity" else _scaleshift_case()
        _check_same_run(want, got)
    finally:
        restore()


@needs_ref
def test_unfusable_configurations_delegate(stub_engine):
    """A NumPy likelihood (no device_spec) or a plain-function transform cannot be fused: every
    call goes to the reference method and the run is still the reference's."""
    restore = _restore_integrator()
    try:
        want = _run(numpy_gauss, lambda u: u, fused=False, max_ncalls=3000)
        import ultranest_b200
        ultranest_b200.install(force=True)
        got = _run(numpy_gauss, lambda u: u, fused=True, max_ncalls=3000)
        assert got["stats"]["fused_calls"] == 0 and got["stats"]["delegated_calls"] > 0
        assert (got["niter"], got["ncall"], got["logz"]) == (want["niter"], want["ncall"], want["logz"])
    finally:
        restore()


def _fake_sampler(region, tregion, loglike, transform, ndim):
    s = types.SimpleNamespace()
    s.region, s.tregion, s.loglike, s.transform = region, tregion, loglike, transform
    s.draw_multiple, s.x_dim, s.num_params = True, ndim, ndim
    s.sampling_slow_warned, s.ncall_region = False, 0
    return s


def _staged_refill(s, Lmin, ndraw):
    """integrator.py:1773-1805, 1836-1837 restated with the sampler's own callables."""
    u = s.region.sample(nsamples=ndraw)
    nu = len(u)
    if nu == 0:
        return np.empty((0, s.x_dim)), np.empty((0, s.x_dim)), np.empty(0), 0
    v = s.transform(u)
    logl = np.ones(nu) * -np.inf
    accepted = s.tregion.inside(v) if s.tregion is not None else np.ones(nu, dtype=bool)
    nt = accepted.sum()
    if nt > 0:
        logl[accepted] = s.loglike(v[accepted, :])
    accepted = logl > Lmin
    return u[accepted, :], v[accepted, :], logl[accepted], nt


def _region_fixture(ml, ndim, nlive, seed):
    rng = np.random.RandomState(seed)
    u = 0.5 + 0.08 * rng.normal(size=(nlive, ndim)) * np.linspace(0.3, 1.0, ndim)
    u = u[np.logical_and(u > 0, u < 1).all(axis=1)]
    layer = ml.AffineLayer()
    layer.optimize(u, u)
    region = ml.MLFriends(u, layer)
    np.random.seed(seed)
    region.maxradiussq, region.enlarge = region.compute_enlargement(nbootstraps=10)
    region.create_ellipsoid()
    return region


def _check_refill_against_staged(ml, ndim=4, with_tregion=True):
    from ultranest_b200 import refill
    from ultranest_b200.likelihoods import GaussianLogLike
    from ultranest_b200.transforms import ScaleShiftTransform
    region = _region_fixture(ml, ndim, 300, 5)
    transform = ScaleShiftTransform(LO, HI)
    loglike = GaussianLogLike(1.0, 0.7)
    tregion = None
    if with_tregion:
        tregion = ml.WrappingEllipsoid(transform(region.u))
        np.random.seed(2)
        tregion.enlarge = 0.6 * tregion.compute_enlargement(nbootstraps=5)   # cut some members off
        tregion.create_ellipsoid()
    Lmin = np.median(loglike(transform(region.u)))
    for name in region._method_names:
        for ndraw in (1, 40, 3000):
            outs = []
            for fused in (False, True):
                region.current_sampling_method = getattr(region, name)
                s = _fake_sampler(region, tregion, loglike, transform, ndim)
                np.random.seed(77)
                if fused:
                    out = refill.refill_samples(s, Lmin, ndraw, 1)
                    assert out is not None
                    outs.append(out[:4] + (np.random.uniform(),))
                else:
                    outs.append(_staged_refill(s, Lmin, ndraw) + (np.random.uniform(),))
            a, b = outs
            assert a[3] == b[3], (name, ndraw, "nc")
            assert a[4] == b[4], (name, ndraw, "RNG stream position")
            for x, y in zip(a[:3], b[:3]):
                np.testing.assert_array_equal(x, y, err_msg="%s ndraw=%d" % (name, ndraw))


@pytest.mark.parametrize("with_tregion", [False, True])
def test_refill_equals_staged_chain_on_host_mirror(stub_engine, with_tregion):
    from ultranest_b200 import mlfriends as ml
    _check_refill_against_staged(ml, with_tregion=with_tregion)


# ---- GPU tier ---------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("ndim,nlive", [(2, 120), (5, 400), (20, 1500), (40, 900)])
def test_device_refill_matches_oracle_chain(ndim, nlive):
    from oracle_engine import OracleEngine
    from ultranest_b200 import _native, mlfriends as ml
    region = _region_fixture(ml, ndim, nlive, 100 + ndim)
    eng = region._bind()
    orc = OracleEngine()
    orc.region_sync_live(region.unormed)
    orc.region_set_radius(region.maxradiussq)
    kind, shift, mat = region.transformLayer._device_params(ndim)
    orc.region_set_layer(kind, shift, mat, ndim)
    orc.region_set_ellipsoid(region.ellipsoid_center, region.ellipsoid_invcov, region.enlarge)
    rng = np.random.RandomState(ndim)
    n = 20000
    # proposals around the live points: a mix of members / non-members, some outside the cube
    base = region.u[rng.randint(len(region.u), size=n)]
    u = base + rng.normal(size=(n, ndim)) * 0.02 * rng.uniform(0, 3, size=(n, 1))
    u[::97] += 1.0
    scale, lo = np.linspace(1.0, 4.0, ndim), np.linspace(-2.0, 1.0, ndim)

    def ellipsoid_of(pts):     # transformed-space ellipsoid that cuts some members off
        cov = np.cov(pts, rowvar=False) + 1e-12 * np.eye(ndim)
        return pts.mean(axis=0), np.linalg.inv(cov), float(ndim) * 1.2

    tregs = {False: ellipsoid_of(region.u), True: ellipsoid_of(region.u * scale + lo)}
    lps = {False: np.concatenate([region.u.mean(axis=0), [0.05, 0.5 * np.log(2 * np.pi * 0.05**2) * ndim]]),
           True: np.concatenate([(region.u * scale + lo).mean(axis=0),
                                 [0.3, 0.5 * np.log(2 * np.pi * 0.3**2) * ndim]])}
    for mode in (0, 1, 2):
        for check_cube in (False, True):
            for xform in (None, (scale, lo)):
                for tregion in (None, tregs[xform is not None]):
                    cases = [(_native.LOGLIKE_GAUSS, lps[xform is not None]),
                             (_native.LOGLIKE_ROSENBROCK, None)]
                    like_kind, lp = cases[(mode + (tregion is None)) % 2]
                    ref_like = orc.region_refill(u, mode, check_cube, xform, tregion, like_kind, lp, -np.inf)[1]
                    finite = ref_like[np.isfinite(ref_like)]
                    Lmin = np.median(finite) if len(finite) else 0.0
                    want = orc.region_refill(u, mode, check_cube, xform, tregion, like_kind, lp, Lmin)
                    got = eng.region_refill(u, mode, check_cube, xform, tregion, like_kind, lp, Lmin)
                    tag = "mode=%d cube=%s xform=%s treg=%s" % (mode, check_cube, xform is not None,
                                                                 tregion is not None)
                    np.testing.assert_array_equal(got[0], want[0], err_msg=tag)
                    np.testing.assert_array_equal(got[1], want[1], err_msg=tag)
                    assert got[2] == want[2], tag
                    if mode == 2:
                        assert 0 < want[2][2] < want[2][1] <= want[2][0] < n, (tag, want[2])
                        if tregion is not None:
                            assert want[2][1] < want[2][0], (tag, want[2])


@pytest.mark.gpu
def test_device_refill_chunked_and_empty():
    from ultranest_b200 import _native, mlfriends as ml
    region = _region_fixture(ml, 6, 500, 9)
    eng = region._bind()
    rng = np.random.RandomState(0)
    u = region.u[rng.randint(len(region.u), size=70001)] + rng.normal(size=(70001, 6)) * 0.01
    lp = np.concatenate([np.full(6, 0.5), [0.1, 0.0]])
    whole = eng.region_refill(u, 2, True, None, None, _native.LOGLIKE_GAUSS, lp, -20.0)
    eng.set_option(_native.OPT_CHUNK_ROWS, 4096)
    try:
        parts = eng.region_refill(u, 2, True, None, None, _native.LOGLIKE_GAUSS, lp, -20.0)
    finally:
        eng.set_option(_native.OPT_CHUNK_ROWS, 0)
    np.testing.assert_array_equal(whole[0], parts[0])
    np.testing.assert_array_equal(whole[1], parts[1])
    assert whole[2] == parts[2]
    assert whole[2][0] == int((whole[0] & 1).sum()) and whole[2][2] == int(((whole[0] & 4) != 0).sum())
    empty = eng.region_refill(np.empty((0, 6)), 2, True, None, None, _native.LOGLIKE_GAUSS, lp, 0.0)
    assert len(empty[0]) == 0 and empty[2] == (0, 0, 0)
    with pytest.raises(ValueError):
        eng.region_refill(u[:10], 2, True, None, None, _native.LOGLIKE_NONE, None, 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("with_tregion", [False, True])
def test_refill_equals_staged_chain_on_device(with_tregion):
    from ultranest_b200 import mlfriends as ml
    _check_refill_against_staged(ml, with_tregion=with_tregion)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("case", ["identity", "scaleshift"])
def test_fused_refill_run_is_the_reference_run(case):
    restore = _restore_integrator()
    try:
        want, got = _identity_case() if case == "identity" else _scaleshift_case()
        _check_same_run(want, got)
    finally:
        restore()
