"""CPU tier: the host-side pieces of the step-sampler mirror (direction proposals, unit-cube line
intersection, move diagnostics) consume the RNG and round exactly like the reference's
(ultranest/stepfuncs.pyx:348-533, popstepsampler.py:26-94).  Needs oracle/_ref."""
import types

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def ref():
    oracle.reference()
    import ultranest.popstepsampler as rp
    import ultranest.stepfuncs as rs
    return rs, rp


def _region(seed, n, d):
    rng = np.random.RandomState(seed)
    u = rng.uniform(0.2, 0.8, size=(n, d))
    A = rng.normal(size=(d, d))
    layer = types.SimpleNamespace(axes=A / np.linalg.norm(A, axis=1, keepdims=True),
                                  transform=lambda x: np.dot(x - 0.5, A))
    return types.SimpleNamespace(u=u, transformLayer=layer, maxradiussq=0.3)


GENERATORS = ["generate_cube_oriented_direction", "generate_cube_oriented_direction_scaled",
              "generate_random_direction", "generate_region_oriented_direction",
              "generate_region_random_direction", "generate_differential_direction",
              "generate_mixture_random_direction"]


@pytest.mark.parametrize("name", GENERATORS)
@pytest.mark.parametrize("scale", [1, 0.3, 1.7])
def test_direction_generators_bitexact_and_rng_aligned(ref, name, scale):
    from ultranest_b200 import stepfuncs as ours
    rs, _ = ref
    region = _region(3, 50, 7)
    ui = region.u[:33]
    np.random.seed(12)
    want = getattr(rs, name)(ui, region, scale=scale)
    tail_want = np.random.uniform()
    np.random.seed(12)
    got = getattr(ours, name)(ui, region, scale=scale)
    tail_got = np.random.uniform()
    np.testing.assert_array_equal(got, want)
    assert tail_got == tail_want   # same number of draws consumed


def test_line_intersection_and_diagnostics(ref):
    from ultranest_b200 import popstepsampler as ours
    _, rp = ref
    rng = np.random.RandomState(4)
    origin = rng.uniform(size=(200, 6))
    direction = rng.normal(size=(200, 6))
    direction[::7, 2] = 0.0    # axis-parallel slabs give NaN/inf intermediates
    a = ours.unitcube_line_intersection(origin, direction)
    b = rp.unitcube_line_intersection(origin, direction)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    region = _region(5, 40, 6)
    fa, (ma, ra) = ours.diagnose_move_distances(region, origin, origin + 0.01 * direction)
    fb, (mb, rb) = rp.diagnose_move_distances(region, origin, origin + 0.01 * direction)
    np.testing.assert_array_equal(fa, fb)
    np.testing.assert_array_equal(ma, mb)
    assert ra == rb


# ---- host logic of the device mirror with the kernels answered by the oracle -------------------

@pytest.fixture()
def stub_engine(monkeypatch):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from ultranest_b200 import _native
    eng = OracleEngine()
    monkeypatch.setattr(_native, "_engine", eng)
    monkeypatch.setattr(_native, "get_engine", lambda: eng)
    return eng


def _device_callables(centre, sigma, lo=None, hi=None):
    from ultranest_b200.likelihoods import GaussianLogLike
    from ultranest_b200.transforms import IdentityTransform, ScaleShiftTransform
    return (IdentityTransform() if lo is None else ScaleShiftTransform(lo, hi)), GaussianLogLike(centre, sigma)


def test_mirror_matches_golden_through_the_stub(stub_engine):
    """Argument marshalling, in-place contracts, RNG draw of the fused evolve, return shapes."""
    import stepfuncs_cases as cases
    import stepfuncs_checks as checks
    from ultranest_b200 import stepfuncs as sf
    g = checks.golden()
    for check in checks.ALL_CHECKS:
        if check is checks.check_evolve:
            xf, ll = _device_callables(0.5, 0.1)
            check(sf, g, xf, ll)                                             # fused entry point
            check(sf, g, cases.identity, cases.gauss_loglike(0.5, 0.1))      # staged around host callables
        else:
            check(sf, g)
    assert stub_engine.calls > 10


def test_sampler_runs_are_the_reference_runs_through_the_stub(stub_engine, ref):
    """PopulationSliceSampler on the installed helpers and PopulationSimpleSliceSampler on the
    device loop return the reference's points (host logic only; the GPU tier repeats this with
    the real kernels)."""
    import test_gpu_stepfuncs as G
    import stepfuncs_cases as cases
    from ultranest_b200 import popstepsampler as pp
    from ultranest_b200 import stepfuncs as sf
    rs, rp = ref
    d = 4
    region = _region(8, 300, d)
    region.transformLayer.transform = lambda x: x * 3.0
    us = region.u
    host_ll = cases.gauss_loglike(0.5, 0.15)
    Ls = host_ll(us)
    make = lambda: rp.PopulationSliceSampler(popsize=30, nsteps=5, generate_direction=rs.generate_mixture_random_direction)  # noqa: E731
    want = G._harvest(make(), region, us, Ls, cases.identity, host_ll, 20, 21)
    xf, ll = _device_callables(0.5, 0.15)
    undo = sf.install()
    try:
        got = G._harvest(make(), region, us, Ls, xf, ll, 20, 21)
    finally:
        sf.uninstall(undo)
    for (u, p, L, nc), (u2, p2, L2, nc2) in zip(got, want):
        np.testing.assert_array_equal(u, u2)
        np.testing.assert_array_equal(p, p2)
        assert L == L2 and nc == nc2
    lo, hi = np.full(d, -1.0), np.full(d, 2.0)
    host_xf = lambda x: x * (hi - lo) + lo   # noqa: E731
    host_ll = cases.gauss_loglike(0.5, 0.5)
    Ls = host_ll(host_xf(us))
    kw = dict(popsize=64, nsteps=4, generate_direction=rs.generate_region_random_direction, shrink_factor=1.2)
    want = G._harvest(rp.PopulationSimpleSliceSampler(**kw), region, us, Ls, host_xf, host_ll, 100, 31, np.min)
    xf, ll = _device_callables(0.5, 0.5, lo, hi)
    sampler = rp.PopulationSimpleSliceSampler(**kw)
    stats = pp.attach(sampler)
    got = G._harvest(sampler, region, us, Ls, xf, ll, 100, 31, np.min)
    assert stats["fused_calls"] > 0 and stats["delegated_calls"] == 0
    for (u, p, L, nc), (u2, p2, L2, nc2) in zip(got, want):
        np.testing.assert_array_equal(u, u2)
        np.testing.assert_array_equal(p, p2)
        assert L == L2 and nc == nc2
    assert sampler.ncalls > 0


def test_package_install_rebinds_step_helpers(stub_engine, ref):
    """ultranest_b200.install(stepfuncs=True) swaps the names popstepsampler imported;
    uninstall_stepfuncs() restores them.  (The region half of install() is exercised, with full
    save/restore of the integrator's bindings, by tests/test_host_logic_cpu.py.)"""
    import sys
    import ultranest_b200
    from ultranest_b200 import stepfuncs as sf
    rs, rp = ref
    region_too = "ultranest.integrator" not in sys.modules   # else install() would need force=True
    saved_mod = sys.modules.get("ultranest.mlfriends")
    saved_attr = getattr(sys.modules["ultranest"], "mlfriends", None)
    before = {n: getattr(rp, n) for n in sf._NAMES if hasattr(rp, n)}
    try:
        if region_too:
            ultranest_b200.install(stepfuncs=True)
        else:
            ultranest_b200._step_undo.extend(sf.install())
        assert rp.evolve is sf.evolve and rp.step_back is sf.step_back
        assert rs.update_vectorised_slice_sampler is sf.update_vectorised_slice_sampler
    finally:
        ultranest_b200.uninstall_stepfuncs()
        if saved_mod is not None:
            sys.modules["ultranest.mlfriends"] = saved_mod
            sys.modules["ultranest"].mlfriends = saved_attr
    for n, fn in before.items():
        assert getattr(rp, n) is fn


def test_integrator_run_with_attached_step_sampler(stub_engine, ref):
    """The unmodified ReactiveNestedSampler driving a PopulationSimpleSliceSampler whose inner
    loop is attached to the device path (here: the oracle stub) makes the reference's run --
    attach() has to survive the integrator's keyword-argument call (integrator.py:1896-1900)."""
    from ultranest import ReactiveNestedSampler
    from ultranest_b200 import popstepsampler as pp
    from ultranest_b200.likelihoods import GaussianLogLike
    from ultranest_b200.transforms import IdentityTransform
    rs, rp = ref
    d, sigma = 3, 0.1

    def numpy_loglike(theta):
        return -0.5 * (((theta - 0.5) / sigma)**2).sum(axis=1) - 0.5 * np.log(2 * np.pi * sigma**2) * d

    def run(loglike, transform, attach):
        np.random.seed(5)
        sampler = ReactiveNestedSampler(["a", "b", "c"], loglike, transform=transform, log_dir=None,
                                        vectorized=True)
        sampler.stepsampler = rp.PopulationSimpleSliceSampler(
            popsize=40, nsteps=4, generate_direction=rs.generate_mixture_random_direction)
        stats = pp.attach(sampler.stepsampler) if attach else None
        res = sampler.run(min_num_live_points=100, max_ncalls=6000, viz_callback=False, show_status=False)
        return res["logz"], res["ncall"], res["niter"], stats

    want = run(numpy_loglike, lambda x: x, False)
    got = run(GaussianLogLike(0.5, sigma), IdentityTransform(), True)
    assert got[3]["fused_calls"] > 0 and got[3]["delegated_calls"] == 0
    assert got[:3] == want[:3]


def test_mirror_rejects_unusable_buffers_before_any_device_call(stub_engine):
    """The reference writes through typed memoryviews and raises on a wrong dtype / layout; the
    mirror must refuse the same inputs instead of silently working on a copy."""
    from ultranest_b200 import stepfuncs as sf
    calls = stub_engine.calls
    f = np.zeros(4)
    b = np.zeros(4, dtype=bool)
    with pytest.raises(ValueError):    # float32 state
        sf.evolve_update(b, f[:0], 0.0, b, b, np.zeros(4, dtype=np.float32), f.copy(), f.copy(),
                         b.copy(), b.copy(), b.copy())
    with pytest.raises(ValueError):    # non-contiguous in/out array
        sf.evolve_update(b, f[:0], 0.0, b, b, np.zeros(8)[::2], f.copy(), f.copy(), b.copy(), b.copy(), b.copy())
    with pytest.raises(ValueError):    # int flags
        sf.evolve_prepare(np.zeros(4, dtype=np.int64), b)
    with pytest.raises(ValueError):    # generation must be int64
        sf.step_back(0.0, np.zeros((4, 3)), np.zeros(4, dtype=np.int32), f.copy())
    with pytest.raises(ValueError):    # arrays shorter than popsize
        sf.update_vectorised_slice_sampler(f, f.copy(), f.copy(), f, np.zeros((4, 2)), np.zeros((4, 2)),
                                           np.arange(4, dtype=np.int64), np.zeros(4, dtype=np.int64), 0.0, 1.0,
                                           np.zeros((4, 2)), f.copy(), np.zeros((4, 2)), 5)
    assert stub_engine.calls == calls
    # empty populations are fine and touch nothing
    e = np.empty(0)
    eb = np.empty(0, dtype=bool)
    assert sf.within_unit_cube(np.empty((0, 3))).shape == (0,)
    sf.evolve_update(eb, e, 0.0, eb, eb, e.copy(), e.copy(), e.copy(), eb.copy(), eb.copy(), eb.copy())
    sf.step_back(0.0, np.empty((0, 4)), np.empty(0, dtype=np.int64), e.copy())


def test_mirror_rejects_mismatched_shapes_and_generations(stub_engine, ref):
    """the mirrors hand raw pointers to the C ABI, so every per-walker array must agree
    in length BEFORE the call (the reference's typed memoryviews raise there), and `step_back`
    must raise IndexError for a generation outside the chain like NumPy does in the reference."""
    from ultranest_b200 import stepfuncs as sf
    rs, _ = ref
    calls = stub_engine.calls
    n, d = 6, 3
    f = np.zeros(n)
    b = np.zeros(n, dtype=bool)
    with pytest.raises(ValueError):    # current_left one short
        sf.evolve_update(b, f[:0], 0.0, b, b, f.copy(), np.zeros(n - 1), f.copy(), b.copy(), b.copy(), b.copy())
    with pytest.raises(ValueError):    # fewer likelihoods than acceptable walkers
        sf.evolve_update(np.ones(n, dtype=bool), f[:2], 0.0, b, b, f.copy(), f.copy(), f.copy(),
                         b.copy(), b.copy(), b.copy())
    with pytest.raises(ValueError):    # allu / proposed_u disagree in ndim
        sf.update_vectorised_slice_sampler(f, f.copy(), f.copy(), f, np.zeros((n, d)), np.zeros((n, d)),
                                           np.arange(n, dtype=np.int64), np.zeros(n, dtype=np.int64), 0.0, 1.0,
                                           np.zeros((n, d + 1)), f.copy(), np.zeros((n, d)), n)
    with pytest.raises(ValueError):    # allp / proposed_p disagree
        sf.update_vectorised_slice_sampler(f, f.copy(), f.copy(), f, np.zeros((n, d)), np.zeros((n, d)),
                                           np.arange(n, dtype=np.int64), np.zeros(n, dtype=np.int64), 0.0, 1.0,
                                           np.zeros((n, d)), f.copy(), np.zeros((n, d + 2)), n)
    with pytest.raises(ValueError):    # generation / currentt shorter than the chains
        sf.step_back(0.0, np.zeros((n, 4)), np.zeros(n - 1, dtype=np.int64), f.copy())
    # a walker that has to step back from a generation outside its chain: IndexError in both
    allL = np.full((n, 4), 1.0)
    allL[2, 1] = -5.0                  # below the threshold -> walker 2 is "problematic"
    gen = np.array([1, 2, 7, 1, 0, 3], dtype=np.int64)
    with pytest.raises(IndexError):
        rs.step_back(0.0, allL.copy(), gen.copy(), f.copy())
    with pytest.raises(IndexError):
        sf.step_back(0.0, allL.copy(), gen.copy(), f.copy())
    assert stub_engine.calls == calls
    # an out-of-range generation of a walker that does NOT step back is harmless in both
    allL2 = np.full((n, 4), 1.0)
    allL2[0, 0] = -5.0
    for impl in (rs, sf):
        a, g, t = allL2.copy(), gen.copy(), f.copy()
        impl.step_back(0.0, a, g, t)
        assert g[2] == 7
