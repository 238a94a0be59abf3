"""Device-backed mirror of ``ultranest.stepfuncs``.

Same names, argument meaning, in-place contracts and RNG consumption as the compiled helpers of
the reference (``ultranest/stepfuncs.pyx``); the per-walker loops run as CUDA kernels through the
C ABI (``unb_within_unit_cube``, ``unb_evolve_prepare``, ``unb_evolve_update``, ``unb_evolve``,
``unb_step_back``, ``unb_update_vectorised_slice_sampler``).  Random draws stay on the host's
``np.random`` stream in the reference's order, so a seeded sampler run sees the same numbers.

Drop-in: ``ultranest.popstepsampler`` binds these names at import
(``from .stepfuncs import evolve, step_back, ...``, popstepsampler.py:14-22), so
:func:`install` rebinds them there (and in ``ultranest.stepfuncs``).

:func:`evolve` runs as ONE kernel (proposal, unit-cube test, prior transform, likelihood,
slice-state update) when ``transform`` and ``loglike`` carry a ``device_spec``
(:mod:`ultranest_b200.transforms`, :mod:`ultranest_b200.likelihoods`); any other callables are
called on the host between the device stages, exactly where the reference calls them.
"""
import ctypes

import numpy as np

from . import _native

int_dtype = np.int64

_pnew_empty = np.empty((0, 1))
_Lnew_empty = np.empty(0)


def _inplace(a, dtype, name):
    """The reference writes through typed memoryviews: the caller's array itself must be usable."""
    if not isinstance(a, np.ndarray) or not a.flags.c_contiguous or not a.flags.writeable:
        raise ValueError("%s must be a writeable C-contiguous ndarray" % name)
    if dtype is bool:
        if a.dtype not in (np.bool_, np.uint8):
            raise ValueError("%s: buffer dtype mismatch, expected bool" % name)
    elif a.dtype != dtype:
        raise ValueError("%s: buffer dtype mismatch, expected %s" % (name, np.dtype(dtype)))
    return a


def _same_length(n, **arrays):
    """The reference's typed memoryviews raise on a length mismatch; here it would be an
    out-of-bounds host copy, so the mirrors check before the engine call."""
    for name, a in arrays.items():
        if np.ndim(a) != 1 or len(a) != n:
            raise ValueError("%s: expected a 1-d array of length %d, got shape %s"
                             % (name, n, np.shape(a)))


def _flags(a):
    a = np.asarray(a)
    if a.dtype not in (np.bool_, np.uint8):
        raise ValueError("buffer dtype mismatch, expected bool")
    return np.ascontiguousarray(a)


def within_unit_cube(u):
    """Whether all fields are strictly between 0 and 1, for each row (stepfuncs.pyx:37-52)."""
    u = _native.as_f64(u, 2)
    acceptable = np.ones(u.shape[0], dtype=bool)
    if len(u):
        _native.get_engine().call("unb_within_unit_cube", _native._ptr(u), u.shape[0], u.shape[1],
                                  _native._ptr(acceptable))
    return acceptable


def evolve_prepare(searching_left, searching_right):
    """Auxiliary slice sampler state selectors (stepfuncs.pyx:72-94)."""
    sl, sr = _flags(searching_left), _flags(searching_right)
    search_right = np.empty_like(searching_left)
    bisecting = np.empty_like(searching_left)
    if len(sl):
        _native.get_engine().call("unb_evolve_prepare", _native._ptr(sl), _native._ptr(sr), len(sl),
                                  _native._ptr(search_right), _native._ptr(bisecting))
    return search_right, bisecting


def evolve_update(acceptable, Lnew, Lmin, search_right, bisecting, currentt, current_left,
                  current_right, searching_left, searching_right, success):
    """Update the state of each walker (stepfuncs.pyx:99-183).  Writes to ``currentt``,
    ``current_left``, ``current_right``, ``searching_left``, ``searching_right``, ``success``."""
    acc, srm, bis = _flags(acceptable), _flags(search_right), _flags(bisecting)
    Lnew = _native.as_f64(Lnew, 1)
    n = len(acc)
    _inplace(currentt, np.float64, "currentt")
    _inplace(current_left, np.float64, "current_left")
    _inplace(current_right, np.float64, "current_right")
    _inplace(searching_left, bool, "searching_left")
    _inplace(searching_right, bool, "searching_right")
    _inplace(success, bool, "success")
    _same_length(n, search_right=srm, bisecting=bis, currentt=currentt, current_left=current_left,
                 current_right=current_right, searching_left=searching_left,
                 searching_right=searching_right, success=success)
    if len(Lnew) < int(np.count_nonzero(acc)):
        raise ValueError("Lnew has %d entries for %d acceptable walkers" % (len(Lnew), np.count_nonzero(acc)))
    if n == 0:
        return
    p = _native._ptr
    _native.get_engine().call("unb_evolve_update", p(acc), p(Lnew), len(Lnew), float(Lmin), p(srm),
                              p(bis), p(currentt), p(current_left), p(current_right),
                              p(searching_left), p(searching_right), p(success), n)


def _device_specs(transform, loglike, ndim):
    """``(xform, kind, lparams)`` when both callables can run inside the kernel, else ``None``."""
    xs = getattr(transform, 'device_spec', None)
    ls = getattr(loglike, 'device_spec', None)
    if xs is None or ls is None:
        return None
    kind, lparams = ls(ndim)
    return xs(ndim), kind, lparams


def evolve(transform, loglike, Lmin, currentu, currentL, currentt, currentv, current_left,
           current_right, searching_left, searching_right):
    """Evolve each slice sampling walker (stepfuncs.pyx:189-282); same return value, same
    in-place writes (``currentu`` ends up holding the proposals, like the reference's alias)."""
    n, ndim = currentu.shape
    specs = _device_specs(transform, loglike, ndim)
    if specs is None or n == 0:
        return _evolve_staged(transform, loglike, Lmin, currentu, currentL, currentt, currentv,
                              current_left, current_right, searching_left, searching_right)
    _inplace(currentu, np.float64, "currentu")
    _inplace(currentt, np.float64, "currentt")
    _inplace(current_left, np.float64, "current_left")
    _inplace(current_right, np.float64, "current_right")
    _inplace(searching_left, bool, "searching_left")
    _inplace(searching_right, bool, "searching_right")
    v = _native.as_f64(currentv, 2)
    _same_length(n, currentt=currentt, current_left=current_left, current_right=current_right,
                 searching_left=searching_left, searching_right=searching_right)
    if v.shape != currentu.shape:
        raise ValueError("currentv %s does not match currentu %s" % (v.shape, currentu.shape))
    # the only random numbers of a step: the bisecting walkers' slice coordinates (:255)
    bisecting = ~(searching_left.astype(bool) | searching_right.astype(bool))
    currentt[bisecting] = np.random.uniform(current_left[bisecting], current_right[bisecting])
    xform, kind, lparams = specs
    desc, keep = _native.make_step_desc(ndim, xform, kind, lparams)
    acceptable = np.empty(n, dtype=bool)
    success = np.empty(n, dtype=bool)
    like = np.empty(n)
    p = _native._ptr
    _native.get_engine().call("unb_evolve", ctypes.addressof(desc), float(Lmin), p(currentu), p(v),
                              p(currentt), p(current_left), p(current_right), p(searching_left),
                              p(searching_right), n, ndim, p(acceptable), p(success), p(like))
    del keep
    nc = int(acceptable.sum())
    unew = currentu[success, :]
    if nc:
        pnew = np.asarray(transform(unew)) if len(unew) else np.empty((0, ndim))
        if pnew is unew:
            pnew = unew.copy()
        Lnew = like[success]
    else:
        pnew, Lnew = _pnew_empty[:0], _Lnew_empty
    return ((currentt, currentv, current_left, current_right, searching_left, searching_right),
            (success, unew, pnew, Lnew), nc)


def _evolve_staged(transform, loglike, Lmin, currentu, currentL, currentt, currentv, current_left,
                   current_right, searching_left, searching_right):
    """Host callables: device helpers around the two user calls, stage by stage as the reference."""
    search_right, bisecting = evolve_prepare(searching_left, searching_right)
    unew = currentu
    for sel, coef in ((searching_left, current_left), (search_right, current_right)):
        unew[sel, :] = currentu[sel, :] + currentv[sel, :] * coef[sel].reshape((-1, 1))
    currentt[bisecting] = np.random.uniform(current_left[bisecting], current_right[bisecting])
    unew[bisecting, :] = currentu[bisecting, :] + currentv[bisecting, :] * currentt[bisecting].reshape((-1, 1))
    acceptable = within_unit_cube(unew)
    nc = 0
    if acceptable.any():
        pnew = transform(unew[acceptable, :])
        Lnew = loglike(pnew)
        nc += len(pnew)
    else:
        pnew, Lnew = _pnew_empty, _Lnew_empty
    success = np.zeros_like(searching_left)
    evolve_update(acceptable, Lnew, Lmin, search_right, bisecting, currentt, current_left,
                  current_right, searching_left, searching_right, success)
    return ((currentt, currentv, current_left, current_right, searching_left, searching_right),
            (success, unew[success, :], pnew[success[acceptable], :], Lnew[success[acceptable]]), nc)


def step_back(Lmin, allL, generation, currentt, log=False):
    """Revert walkers whose chain holds a likelihood below ``Lmin`` (stepfuncs.pyx:285-334).
    Updates ``currentt``, ``generation`` and ``allL`` in place."""
    _inplace(allL, np.float64, "allL")
    _inplace(generation, np.int64, "generation")
    _inplace(currentt, np.float64, "currentt")
    if allL.ndim != 2 or generation.shape != (allL.shape[0],) or currentt.shape != (allL.shape[0],):
        raise ValueError("step_back: allL %s, generation %s, currentt %s do not agree"
                         % (allL.shape, generation.shape, currentt.shape))
    if allL.size == 0:
        return
    # allL[i, g] of a walker that has to step back (stepfuncs.pyx:322-327): NumPy raises
    # IndexError for g outside [-ncols, ncols); check the (rare) offenders like the reference
    ncols = allL.shape[1]
    offenders = np.flatnonzero((generation >= ncols) | (generation < -ncols))
    if len(offenders):
        max_width = generation.max() + 1
        for i in offenders:
            if (allL[i, :max_width] < Lmin).any():
                raise IndexError("index %d is out of bounds for axis 1 with size %d"
                                 % (generation[i], ncols))
    p = _native._ptr
    _native.get_engine().call("unb_step_back", float(Lmin), p(allL), allL.shape[0], allL.shape[1],
                              p(generation), p(currentt))


def update_vectorised_slice_sampler(t, tleft, tright, proposed_L, proposed_u, proposed_p,
                                    worker_running, status, Likelihood_threshold, shrink_factor,
                                    allu, allL, allp, popsize):
    """Update the slice sampler state of each walker in the population (stepfuncs.pyx:537-630);
    in place on the state arrays, which are also returned together with ``discarded``."""
    popsize = int(popsize)
    t = _native.as_f64(t, 1)
    pL, pu, pp = _native.as_f64(proposed_L, 1), _native.as_f64(proposed_u, 2), _native.as_f64(proposed_p, 2)
    _inplace(tleft, np.float64, "tleft")
    _inplace(tright, np.float64, "tright")
    _inplace(worker_running, np.int64, "worker_running")
    _inplace(status, np.int64, "status")
    _inplace(allu, np.float64, "allu")
    _inplace(allL, np.float64, "allL")
    _inplace(allp, np.float64, "allp")
    if min(len(t), len(tleft), len(tright), len(pL), len(pu), len(pp), len(worker_running),
           len(status), len(allu), len(allL), len(allp)) < popsize:
        raise ValueError("arrays shorter than popsize=%d" % popsize)
    if pu.ndim != 2 or allu.ndim != 2 or allu.shape[1] != pu.shape[1]:
        raise ValueError("allu %s does not match proposed_u %s" % (allu.shape, pu.shape))
    if pp.ndim != 2 or allp.ndim != 2 or allp.shape[1] != pp.shape[1]:
        raise ValueError("allp %s does not match proposed_p %s" % (allp.shape, pp.shape))
    _same_length(len(tleft), tright=tright, status=status)
    discarded = np.zeros(1, dtype=np.int64)
    p = _native._ptr
    if popsize:
        _native.get_engine().call(
            "unb_update_vectorised_slice_sampler", p(t), p(tleft), p(tright), p(pL), p(pu), p(pp),
            p(worker_running), p(status), float(Likelihood_threshold), float(shrink_factor), p(allu),
            p(allL), p(allp), popsize, pu.shape[1], pp.shape[1], p(discarded))
    return (tleft, tright, worker_running, status, allu, allL, allp, int(discarded[0]))


# -- slice direction proposals (stepfuncs.pyx:348-533): host RNG in the reference's order ----------

def _axis_vectors(nsamples, ndim, scale):
    """Rows with one non-zero entry on a random axis.  The reference routes ``scale`` through a C
    ``float`` argument (stepfuncs.pyx:337-345), hence the float32 round trip."""
    v = np.zeros((nsamples, ndim))
    j = np.random.randint(ndim, size=nsamples, dtype=int_dtype)
    v[np.arange(nsamples), j] = float(np.float32(scale))
    return v, j


def generate_cube_oriented_direction(ui, region, scale=1):
    """Direction along a random unit cube axis, length ``scale`` (stepfuncs.pyx:348-370)."""
    nsamples, ndim = ui.shape
    return _axis_vectors(nsamples, ndim, scale)[0]


def generate_cube_oriented_direction_scaled(ui, region, scale=1):
    """Random cube axis, scaled by the live points' spread along it (stepfuncs.pyx:373-398)."""
    nsamples, ndim = ui.shape
    scales = region.u.std(axis=0)
    v, j = _axis_vectors(nsamples, ndim, scale)
    v *= scales[j].reshape((-1, 1))
    return v


def generate_random_direction(ui, region, scale=1):
    """Isotropic direction of length ``scale`` in cube space (stepfuncs.pyx:400-421)."""
    nsamples, ndim = ui.shape
    v = np.random.normal(size=(nsamples, ndim))
    v *= scale / np.linalg.norm(v, axis=1).reshape((nsamples, 1))
    return v


def generate_region_oriented_direction(ui, region, scale=1):
    """Direction along a random axis of the region's layer (stepfuncs.pyx:424-448)."""
    nsamples, ndim = ui.shape
    j = np.random.randint(ndim, size=nsamples, dtype=int_dtype)
    return region.transformLayer.axes[j] * scale


def generate_region_random_direction(ui, region, scale=1):
    """Isotropic t-space direction mapped through the layer axes (stepfuncs.pyx:451-475)."""
    nsamples, ndim = ui.shape
    w = np.random.normal(size=(nsamples, ndim))
    w *= scale / np.linalg.norm(w, axis=1).reshape((nsamples, 1))
    return np.einsum('ij,kj->ki', region.transformLayer.axes, w)


def generate_differential_direction(ui, region, scale=1):
    """Difference of two distinct random live points (stepfuncs.pyx:477-503)."""
    nsamples = ui.shape[0]
    nlive = region.u.shape[0]
    first = np.random.randint(nlive, size=nsamples, dtype=int_dtype)
    second = np.random.randint(nlive - 1, size=nsamples, dtype=int_dtype)
    second[second >= first] += 1
    return (region.u[first, :] - region.u[second, :]) * scale


def generate_mixture_random_direction(ui, region, scale=1):
    """Per walker, differential or region-oriented proposal with equal odds
    (stepfuncs.pyx:507-533)."""
    nsamples = ui.shape[0]
    v_de = generate_differential_direction(ui, region, scale=scale)
    v_axis = generate_region_oriented_direction(ui, region, scale=scale)
    coin = np.random.uniform(size=nsamples).reshape((-1, 1))
    return np.where(coin < 0.5, v_de, v_axis)


_NAMES = ["within_unit_cube", "evolve_prepare", "evolve_update", "evolve", "step_back",
          "update_vectorised_slice_sampler", "generate_cube_oriented_direction",
          "generate_cube_oriented_direction_scaled", "generate_random_direction",
          "generate_region_oriented_direction", "generate_region_random_direction",
          "generate_differential_direction", "generate_mixture_random_direction"]
__all__ = _NAMES + ["int_dtype", "install"]


def install(modules=None):
    """Rebind the helper names in the reference's modules (``ultranest.stepfuncs`` and
    ``ultranest.popstepsampler``, which imported them by name).  Returns what was replaced so the
    caller can restore it."""
    import importlib
    if modules is None:
        modules = [importlib.import_module("ultranest.stepfuncs"),
                   importlib.import_module("ultranest.popstepsampler")]
    here = globals()
    undo = []
    for mod in modules:
        for name in _NAMES:
            if hasattr(mod, name):
                undo.append((mod, name, getattr(mod, name)))
                setattr(mod, name, here[name])
    return undo


def uninstall(undo):
    for mod, name, fn in reversed(undo):
        setattr(mod, name, fn)
