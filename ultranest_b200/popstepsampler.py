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


def fused_next(self, region, Lmin, us, Ls, transform, loglike, ndraw=10, plot=False, tregion=None,
               log=False, test=False):
    """``PopulationSimpleSliceSampler.__next__`` (popstepsampler.py:863-1001) with the pass loop
    on the device.  Same return value, RNG consumption, counters and adaptation."""
    nlive, ndim = us.shape
    if len(self.prepared_samples) == 0:
        ilive = np.random.randint(0, nlive, size=self.popsize)
        allu = np.array(us[ilive, :]) if not test else np.array(us)
        allp = np.zeros((self.popsize, ndim)) * np.nan
        allL = np.array(Ls[ilive])
        nc = 0
        n_discarded = 0
        interval_final = 0.
        loop = SliceLoop(transform, loglike, ndim)
        for _k in range(self.nsteps):
            factor_scale = self.scale_jitter_func()
            v = self.generate_direction(allu, region, scale=1.0) * self.scale * factor_scale
            tleft_unitcube, tright_unitcube = unitcube_line_intersection(allu, v)
            # the reference derives the per-worker and the per-point limits by two identical
            # calls (:926-929); workers start on their own point, so one pair serves both
            self.slice_limit(tleft_unitcube, tright_unitcube)
            tleft, tright = self.slice_limit(tleft_unitcube, tright_unitcube)
            loop.begin(allu, allL, v, tleft, tright, Lmin, self.shrink_factor)
            for _it in range(self.max_it):
                slice_position = np.random.uniform(size=(self.popsize,))
                n_running, n_discarded_it = loop.iterate(slice_position)
                nc += self.popsize
                n_discarded += n_discarded_it
                if n_running == 0:
                    break
            allu, allp_step, allL, tleft, tright, status = loop.end()
            moved = status == 1
            allp[moved, :] = allp_step[moved, :]
            interval_final += np.median(tright - tleft)

        interval_final = interval_final / self.nsteps
        self.discarded += n_discarded
        self.ncalls += nc
        assert np.isfinite(allp).all(), 'some walkers never moved! Double nsteps of PopulationSimpleSliceSampler.'
        far_enough, (move_distance, reference_distance) = diagnose_move_distances(region, us[ilive, :], allu)
        self.prepared_samples = list(zip(allu, allp, allL))
        self.logstat.append([
            self.popsize / nc,
            self.scale,
            self.nsteps,
            np.mean(far_enough) if len(far_enough) > 0 else 0,
            np.exp(np.mean(np.log(move_distance / reference_distance + 1e-10))) if len(far_enough) > 0 else 0
        ])
        if interval_final >= 1. / self.adapt_slice_scale_target:
            self.scale *= 1. / self.scale_adapt_factor
        else:
            self.scale *= self.scale_adapt_factor
    else:
        nc = 0
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
