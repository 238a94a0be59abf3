"""Device-resident inner loop for ``ultranest.popstepsampler.PopulationSimpleSliceSampler``
(SURVEY 8-f rank 2).

The reference's ``__next__`` (popstepsampler.py:863-1001) refills its pool of ``popsize`` chains by
``nsteps`` slice moves; each move loops up to ``max_it`` times over

    draw ``popsize`` uniforms -> ``t`` on every worker's slice -> proposals ``allu[w] + t * v[w]``
    -> ``transform`` -> ``loglike`` -> ``update_vectorised_slice_sampler``          (:940-965)

with seven ``popsize``-sized temporaries per pass.  :func:`attach` rebinds ``__next__`` on a
sampler instance so that this loop runs on the device with the population resident there
(``unb_popslice_begin / _iterate / _end``): per pass only the uniforms go up and two counters come
back.  Everything outside the loop -- live-point choice, direction proposals, unit-cube
intersections, scale adaptation, diagnostics -- stays the reference's host NumPy in the
reference's RNG order, so a seeded run returns the same points.

Needs a prior transform and a likelihood with a ``device_spec``
(:mod:`ultranest_b200.transforms`, :mod:`ultranest_b200.likelihoods`); otherwise the call is
delegated to the reference method.
"""
import ctypes
import types

import numpy as np

from . import _native


def unitcube_line_intersection(ray_origin, ray_direction):
    """Where the line ``origin + t * direction`` leaves the unit cube, as ``(t_neg, t_pos)``
    (popstepsampler.py:26-61; slab method, NaN for axes the line does not move along)."""
    assert (ray_origin >= 0).all(), ray_origin
    assert (ray_origin <= 1).all(), ray_origin
    assert ((ray_direction**2).sum()**0.5 > 1e-200).all(), ray_direction
    with np.errstate(divide='ignore', invalid='ignore'):
        inv = 1. / ray_direction
        mid = inv * (ray_origin - 0.5)
        half = np.abs(inv) * 0.5
        return np.nanmax(-mid - half, axis=1), np.nanmin(-mid + half, axis=1)


def diagnose_move_distances(region, ustart, ufinal):
    """Whitened travel distance against the MLFriends radius (popstepsampler.py:64-94)."""
    assert ustart.shape == ufinal.shape, (ustart.shape, ufinal.shape)
    tstart = region.transformLayer.transform(ustart)
    tfinal = region.transformLayer.transform(ufinal)
    d2 = ((tstart - tfinal)**2).sum(axis=1)
    return d2 > region.maxradiussq, [d2**0.5, region.maxradiussq**0.5]


class SliceLoop(object):
    """One slice move of the whole population on the device (popstepsampler.py:916-965)."""

    def __init__(self, transform, loglike, ndim, engine=None):
        xs, ls = transform.device_spec, loglike.device_spec
        kind, lparams = ls(ndim)
        self.desc, self._keep = _native.make_step_desc(ndim, xs(ndim), kind, lparams)
        self.eng = engine or _native.get_engine()
        self.ndim = ndim

    def begin(self, allu, allL, v, tleft, tright, Lmin, shrink_factor):
        p = _native._ptr
        self.arrays = [_native.as_f64(a) for a in (allu, allL, v, tleft, tright)]
        a = self.arrays
        self.popsize = len(a[0])
        self.eng.call("unb_popslice_begin", ctypes.addressof(self.desc), p(a[0]), p(a[1]), p(a[2]),
                      p(a[3]), p(a[4]), self.popsize, self.ndim, float(Lmin), float(shrink_factor))

    def iterate(self, slice_position):
        pos = _native.as_f64(slice_position, 1)
        if len(pos) != self.popsize:
            raise ValueError("need one uniform draw per worker")
        n_running, discarded = ctypes.c_int64(0), ctypes.c_int64(0)
        self.eng.call("unb_popslice_iterate", _native._ptr(pos), ctypes.byref(n_running),
                      ctypes.byref(discarded))
        return int(n_running.value), int(discarded.value)

    def end(self):
        n, d = self.popsize, self.ndim
        allu, allp = np.empty((n, d)), np.empty((n, d))
        allL, tleft, tright = np.empty(n), np.empty(n), np.empty(n)
        status = np.empty(n, dtype=np.int64)
        p = _native._ptr
        self.eng.call("unb_popslice_end", p(allu), p(allp), p(allL), p(tleft), p(tright), p(status))
        return allu, allp, allL, tleft, tright, status


def _slice_move(sampler, loop, region, Lmin, allu, allL, allp):
    """One of the ``nsteps`` slice moves of the whole population (popstepsampler.py:913-968):
    direction and limits on the host, the pass loop on the device.  Returns the moved population,
    the likelihood calls spent, the discarded evaluations and the median final slice width."""
    jitter = sampler.scale_jitter_func()
    v = sampler.generate_direction(allu, region, scale=1.0) * sampler.scale * jitter
    cube_lo, cube_hi = unitcube_line_intersection(allu, v)
    # the reference derives the per-worker and the per-point limits by two identical calls
    # (:926-929); workers start on their own point, so the second pair serves both
    sampler.slice_limit(cube_lo, cube_hi)
    tleft, tright = sampler.slice_limit(cube_lo, cube_hi)
    loop.begin(allu, allL, v, tleft, tright, Lmin, sampler.shrink_factor)
    ncalls = ndiscarded = 0
    for _ in range(sampler.max_it):
        n_running, n_disc = loop.iterate(np.random.uniform(size=(sampler.popsize,)))
        ncalls += sampler.popsize
        ndiscarded += n_disc
        if n_running == 0:
            break
    allu, moved_p, allL, tleft, tright, status = loop.end()
    moved = status == 1
    allp[moved, :] = moved_p[moved, :]
    return allu, allL, allp, ncalls, ndiscarded, np.median(tright - tleft)


def _refill_pool(sampler, region, Lmin, us, Ls, transform, loglike, test):
    """Refill ``prepared_samples`` (popstepsampler.py:901-1000): same RNG consumption, counters,
    diagnostics and scale adaptation as the reference; returns the likelihood calls spent."""
    nlive, ndim = us.shape
    ilive = np.random.randint(0, nlive, size=sampler.popsize)
    allu = np.array(us) if test else np.array(us[ilive, :])
    allp = np.full((sampler.popsize, ndim), np.nan)
    allL = np.array(Ls[ilive])
    loop = SliceLoop(transform, loglike, ndim)
    nc = n_discarded = 0
    width_sum = 0.
    for _ in range(sampler.nsteps):
        allu, allL, allp, dc, dd, width = _slice_move(sampler, loop, region, Lmin, allu, allL, allp)
        nc += dc
        n_discarded += dd
        width_sum += width
    mean_width = width_sum / sampler.nsteps
    sampler.discarded += n_discarded
    sampler.ncalls += nc
    assert np.isfinite(allp).all(), 'some walkers never moved! Double nsteps of PopulationSimpleSliceSampler.'
    far_enough, (moved_by, radius) = diagnose_move_distances(region, us[ilive, :], allu)
    sampler.prepared_samples = list(zip(allu, allp, allL))
    have = len(far_enough) > 0
    sampler.logstat.append([
        sampler.popsize / nc, sampler.scale, sampler.nsteps,
        np.mean(far_enough) if have else 0,
        np.exp(np.mean(np.log(moved_by / radius + 1e-10))) if have else 0])
    # widen the slice when the final intervals stay long, shrink it otherwise (:994-998)
    if mean_width >= 1. / sampler.adapt_slice_scale_target:
        sampler.scale *= 1. / sampler.scale_adapt_factor
    else:
        sampler.scale *= sampler.scale_adapt_factor
    return nc


def fused_next(self, region, Lmin, us, Ls, transform, loglike, ndraw=10, plot=False, tregion=None,
               log=False, test=False):
    """``PopulationSimpleSliceSampler.__next__`` (popstepsampler.py:863-1001) with the pass loop
    on the device.  Same return value, RNG consumption, counters and adaptation."""
    nc = 0
    if len(self.prepared_samples) == 0:
        nc = _refill_pool(self, region, Lmin, us, Ls, transform, loglike, test)
    u, p, L = self.prepared_samples.pop(0)
    return u, p, L, nc


def attach(stepsampler, stats=None):
    """Install the device loop on a ``PopulationSimpleSliceSampler`` instance (the integrator calls
    ``stepsampler.__next__(...)`` as an attribute, integrator.py:1896).  Calls whose transform or
    likelihood cannot run on the device go to the reference method."""
    original = stepsampler.__next__
    stats = {'fused_calls': 0, 'delegated_calls': 0} if stats is None else stats

    def __next__(self, region, Lmin, us, Ls, transform, loglike, *args, **kwargs):
        if hasattr(transform, 'device_spec') and hasattr(loglike, 'device_spec'):
            stats['fused_calls'] += 1
            return fused_next(self, region, Lmin, us, Ls, transform, loglike, *args, **kwargs)
        stats['delegated_calls'] += 1
        return original(region, Lmin, us, Ls, transform, loglike, *args, **kwargs)

    stepsampler.__next__ = types.MethodType(__next__, stepsampler)
    stepsampler._unb_stats = stats
    return stats
