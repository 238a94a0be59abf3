"""CPU tier for the HOST logic of ultranest_b200/mlfriends.py: the behavioural / differential
scenarios of tests/test_gpu_region_behaviour.py and the integrator drop-in run again with the
kernels answered by the CPU oracle (tests/oracle_engine.py injected as the engine).  What this
pins without a GPU: RNG consumption order of every sampling method against the reference, the
clustering growth loop and id re-use, layer learning, bootstrap orchestration and error types,
`ultranest_b200.install()`, and that a seeded ReactiveNestedSampler run on top of the mirror is
the reference's run."""
import sys

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built")

import test_gpu_region_behaviour as B   # noqa: E402  (test bodies are reused; their gpu mark is per module)
import test_integrator_e2e as E          # noqa: E402


@pytest.fixture()
def stub_engine(monkeypatch):
    from oracle_engine import OracleEngine
    from ultranest_b200 import _native
    eng = OracleEngine()
    monkeypatch.setattr(_native, "_engine", eng)
    monkeypatch.setattr(_native, "get_engine", lambda: eng)
    return eng


@pytest.fixture()
def ours(stub_engine):
    from ultranest_b200 import mlfriends
    return mlfriends


@pytest.fixture()
def ref():
    oracle.reference()
    import ultranest.mlfriends as m
    return m


@pytest.mark.parametrize("layer_name,scale", [("ScalingLayer", None), ("AffineLayer", None)])
def test_region_sampling_scenarios(ours, layer_name, scale):
    B.test_region_sampling_scenarios(ours, layer_name, scale)


@pytest.mark.parametrize("layer_name", ["ScalingLayer", "AffineLayer"])
def test_sampling_methods_match_reference_stream(ours, ref, layer_name):
    B.test_sampling_methods_match_reference_stream(ours, ref, layer_name)


def test_inside_ellipsoid_equals_einsum(ours):
    B.test_inside_ellipsoid_equals_einsum(ours)


def test_all_region_classes_contain_their_points(ours):
    B.test_all_region_classes_contain_their_points(ours)


def test_ellipsoid_regions_match_reference(ours, ref):
    B.test_ellipsoid_regions_match_reference(ours, ref)


def test_errors_match_reference_types(ours):
    B.test_errors_match_reference_types(ours)


def test_clustering_scenarios(ours, ref):
    B.test_clustering_scenarios(ours, ref)


def test_layers_roundtrip_and_create_new(ours, ref):
    B.test_layers_roundtrip_and_create_new(ours, ref)


def test_bootstrap_rewinds_rng_on_failure(ours, ref):
    """A failing round must leave the caller's RNG where the reference's early exit leaves it."""
    line = np.linspace(0.2, 0.8, 50).reshape((-1, 1)) * np.ones((1, 3))
    states = []
    for mod in (ours, ref):
        lay = mod.ScalingLayer()
        lay.optimize(line, line)
        rng = np.random.RandomState(1)
        with pytest.raises(np.linalg.LinAlgError):
            mod.MLFriends(line, lay).compute_enlargement(nbootstraps=5, rng=rng)
        states.append(rng.randint(1 << 30, size=4))
    assert (states[0] == states[1]).all()


def _with_restored_integrator(fn):
    oracle.reference()
    import ultranest.integrator as integ
    import ultranest.mlfriends as refmod
    names = ("AffineLayer", "LocalAffineLayer", "MLFriends", "RobustEllipsoidRegion",
             "ScalingLayer", "WrappingEllipsoid", "find_nearby")
    saved = {n: getattr(integ, n) for n in names}
    fns = [f for cls in vars(integ).values() if isinstance(cls, type)
           for f in vars(cls).values() if getattr(f, "__defaults__", None)]
    saved_defaults = [(f, f.__defaults__) for f in fns]
    try:
        return fn()
    finally:
        for n, v in saved.items():
            setattr(integ, n, v)
        for f, d in saved_defaults:
            f.__defaults__ = d
        sys.modules["ultranest.mlfriends"] = refmod
        sys.modules["ultranest"].mlfriends = refmod


def test_eggbox_run_on_host_mirror_is_the_reference_run(stub_engine):
    """Multimodal problem: clustering, cluster-centred LocalAffineLayer, tregion."""
    def body():
        want = E.run_eggbox(max_ncalls=5000)
        assert want["nclusters"] > 1
        import ultranest_b200
        ultranest_b200.install(force=True)
        got = E.run_eggbox(max_ncalls=5000)
        assert got["region"] == "ultranest_b200.mlfriends"
        for key in ("niter", "ncall", "ncall_region", "nclusters"):
            assert got[key] == want[key], key
        assert abs(got["logz"] - want["logz"]) <= 1e-10 * abs(want["logz"])
    _with_restored_integrator(body)


def test_integrator_run_on_host_mirror_is_the_reference_run(stub_engine):
    """The unmodified integrator over ultranest_b200.mlfriends (kernels = oracle): identical run."""
    oracle.reference()
    import ultranest.integrator as integ
    import ultranest.mlfriends as refmod
    names = ("AffineLayer", "LocalAffineLayer", "MLFriends", "RobustEllipsoidRegion",
             "ScalingLayer", "WrappingEllipsoid", "find_nearby")
    saved = {n: getattr(integ, n) for n in names}
    fns = [fn for cls in vars(integ).values() if isinstance(cls, type)
           for fn in vars(cls).values() if getattr(fn, "__defaults__", None)]
    saved_defaults = [(fn, fn.__defaults__) for fn in fns]
    try:
        want = E.run_once(E.numpy_loglike, nlive=100, max_ncalls=4000)
        assert want["region"] == "ultranest.mlfriends"
        import ultranest_b200
        ultranest_b200.install(force=True)
        got = E.run_once(E.numpy_loglike, nlive=100, max_ncalls=4000)
        assert got["region"] == "ultranest_b200.mlfriends"
        assert stub_engine.calls > 100
        assert (got["niter"], got["ncall"], got["ncall_region"]) == \
            (want["niter"], want["ncall"], want["ncall_region"])
        assert got["logz"] == want["logz"]
    finally:
        for n, v in saved.items():
            setattr(integ, n, v)
        for fn, d in saved_defaults:
            fn.__defaults__ = d
        sys.modules["ultranest.mlfriends"] = refmod
        sys.modules["ultranest"].mlfriends = refmod
