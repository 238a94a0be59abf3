"""End-to-end drop-in test: the reference's UNMODIFIED ReactiveNestedSampler (oracle/_ref) runs
BASELINE config #1 (5-D Gaussian, N_live=400, vectorized) twice with the same seed -- once on its
own Cython region module, once with ultranest_b200 installed behind the same names -- and must
produce the same run: logZ within 1e-10 relative, identical ncall / niter."""
import sys

import numpy as np
import pytest

import oracle

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built")]

NDIM = 5
SIGMA = 0.01


def numpy_loglike(theta):
    centers = np.ones(NDIM) * 0.5
    return -0.5 * (((theta - centers) / SIGMA)**2).sum(axis=1) - 0.5 * np.log(2 * np.pi * SIGMA**2) * NDIM


def transform(x):
    return x


def run_once(loglike, nlive=400, **runargs):
    from ultranest import ReactiveNestedSampler
    np.random.seed(42)
    names = ["p%d" % i for i in range(NDIM)]
    sampler = ReactiveNestedSampler(names, loglike, transform=transform, log_dir=None,
                                    vectorized=True)
    res = sampler.run(min_num_live_points=nlive, viz_callback=False, show_status=False, **runargs)
    return dict(logz=res["logz"], logzerr=res["logzerr"], ncall=res["ncall"],
                niter=res["niter"], ncall_region=sampler.ncall_region,
                region=type(sampler.region).__module__)


@pytest.fixture()
def swapped_modules():
    """Install ours for the duration of one test, then put the reference's bindings back."""
    oracle.reference()
    import ultranest.integrator as integ
    import ultranest.mlfriends as refmod
    names = ("AffineLayer", "LocalAffineLayer", "MLFriends", "RobustEllipsoidRegion",
             "ScalingLayer", "WrappingEllipsoid", "find_nearby")
    saved = {n: getattr(integ, n) for n in names}
    fns = [fn for cls in vars(integ).values() if isinstance(cls, type)
           for fn in vars(cls).values() if getattr(fn, "__defaults__", None)]
    saved_defaults = [(fn, fn.__defaults__) for fn in fns]
    yield integ, refmod
    for n, v in saved.items():
        setattr(integ, n, v)
    for fn, d in saved_defaults:
        fn.__defaults__ = d
    sys.modules["ultranest.mlfriends"] = refmod
    sys.modules["ultranest"].mlfriends = refmod


def test_gaussian_run_is_identical(swapped_modules):
    want = run_once(numpy_loglike)
    assert want["region"] == "ultranest.mlfriends"
    import ultranest_b200
    ultranest_b200.install(force=True)
    got = run_once(numpy_loglike)
    assert got["region"] == "ultranest_b200.mlfriends"
    assert got["niter"] == want["niter"]
    assert got["ncall"] == want["ncall"]
    assert got["ncall_region"] == want["ncall_region"]
    assert abs(got["logz"] - want["logz"]) <= 1e-10 * abs(want["logz"])
    assert got["logz"] == want["logz"], "same host: expected the identical run, bit for bit"
    # and with the likelihood batch call on the device as well
    from ultranest_b200.likelihoods import GaussianLogLike
    got2 = run_once(GaussianLogLike(0.5, SIGMA))
    assert (got2["niter"], got2["ncall"]) == (want["niter"], want["ncall"])
    assert abs(got2["logz"] - want["logz"]) <= 1e-10 * abs(want["logz"])


def test_gaussian20d_run_prefix_is_identical(swapped_modules):
    """Headline-scale geometry (20-D, 2000 live points, default LocalAffineLayer, clustering,
    repeated bootstrapped rebuilds, every sampling method the run picks): the first 12 000
    likelihood calls of a seeded run must be the same run on both region implementations."""
    from ultranest import ReactiveNestedSampler
    d, sigma = 20, 0.05

    def loglike(t):
        return -0.5 * (((t - 0.5) / sigma)**2).sum(axis=1)

    def once():
        np.random.seed(7)
        sampler = ReactiveNestedSampler(["p%d" % i for i in range(d)], loglike, transform=transform,
                                        log_dir=None, vectorized=True)
        res = sampler.run(min_num_live_points=2000, max_ncalls=12000, viz_callback=False,
                          show_status=False)
        return (res["logz"], res["ncall"], res["niter"], sampler.ncall_region,
                type(sampler.region).__module__, sampler.region.maxradiussq)

    want = once()
    assert want[4] == "ultranest.mlfriends"
    import ultranest_b200
    ultranest_b200.install(force=True)
    got = once()
    assert got[4] == "ultranest_b200.mlfriends"
    assert got[1:4] == want[1:4]
    assert got[5] == want[5]                       # last bootstrapped radius, bit for bit
    assert abs(got[0] - want[0]) <= 1e-10 * abs(want[0])
    assert got[0] == want[0]


def eggbox_loglike(z):
    chi = (np.cos(z / 2.)).prod(axis=1)
    return (2. + chi)**5


def eggbox_transform(x):
    return x * 10 * np.pi


def run_eggbox(ndim=2, nlive=400, max_ncalls=8000, loglike=None):
    """examples/testeggbox.py: many separated modes -> clustering, cluster-centred layers,
    id re-use across rebuilds, a transformed-space wrapping ellipsoid (tregion)."""
    from ultranest import ReactiveNestedSampler
    np.random.seed(3)
    sampler = ReactiveNestedSampler(["a", "b", "c", "d"][:ndim], loglike or eggbox_loglike,
                                    transform=eggbox_transform, log_dir=None, vectorized=True)
    res = sampler.run(min_num_live_points=nlive, max_ncalls=max_ncalls, viz_callback=False,
                      show_status=False)
    return dict(logz=res["logz"], ncall=res["ncall"], niter=res["niter"],
                ncall_region=sampler.ncall_region, nclusters=sampler.transformLayer.nclusters,
                region=type(sampler.region).__module__,
                tregion=type(sampler.tregion).__module__ if sampler.tregion is not None else None)


def test_eggbox_run_is_identical(swapped_modules):
    want = run_eggbox()
    assert want["region"] == "ultranest.mlfriends" and want["nclusters"] > 1
    import ultranest_b200
    ultranest_b200.install(force=True)
    got = run_eggbox()
    assert got["region"] == "ultranest_b200.mlfriends"
    assert got["tregion"] in (None, "ultranest_b200.mlfriends")
    for key in ("niter", "ncall", "ncall_region", "nclusters"):
        assert got[key] == want[key], key
    assert abs(got["logz"] - want["logz"]) <= 1e-10 * abs(want["logz"])


def test_eggbox_run_with_device_likelihood_logz(swapped_modules):
    """The same seeded eggbox run with the DEVICE likelihood batch call (`EggboxLogLike`: CUDA
    cos/pow, <= 2 ulp from libm, so not bit-identical to NumPy): logZ must stay within the 1e-10
    relative bar of BASELINE.json, and -- ulp-level likelihood noise never reorders the dead
    points here -- the run must be the same run (niter, ncall, cluster count)."""
    want = run_eggbox()
    import ultranest_b200
    from ultranest_b200.likelihoods import EggboxLogLike
    ultranest_b200.install(force=True)
    got = run_eggbox(loglike=EggboxLogLike())
    assert got["region"] == "ultranest_b200.mlfriends"
    for key in ("niter", "ncall", "ncall_region", "nclusters"):
        assert got[key] == want[key], key
    assert abs(got["logz"] - want["logz"]) <= 1e-10 * abs(want["logz"])


def test_region_class_argument(swapped_modules):
    """The per-run plug-in point (no module swap): run(region_class=...) +
    sampler.transform_layer_class (integrator.py:2298, 1137)."""
    from ultranest import ReactiveNestedSampler
    from ultranest_b200 import mlfriends as ours
    np.random.seed(1)
    sampler = ReactiveNestedSampler(["a", "b", "c"], lambda t: -0.5 * (((t - 0.5) / 0.05)**2).sum(axis=1),
                                    transform=transform, log_dir=None, vectorized=True)
    sampler.transform_layer_class = ours.AffineLayer
    res = sampler.run(min_num_live_points=100, region_class=ours.RobustEllipsoidRegion,
                      viz_callback=False, show_status=False, max_ncalls=20000)
    assert isinstance(sampler.region, ours.RobustEllipsoidRegion)
    assert np.isfinite(res["logz"])
    # analytic: integral of an (unnormalised) 3-D Gaussian of sigma 0.05 inside the unit cube
    assert abs(res["logz"] - 1.5 * np.log(2 * np.pi * 0.05**2)) < 0.5
