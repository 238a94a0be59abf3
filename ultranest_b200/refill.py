"""Fused ``_refill_samples``.

``ReactiveNestedSampler._refill_samples`` (integrator.py:1773-1837) chains five host stages per
batch of proposals -- ``region.sample`` (draw, transform, neighbour scan, compaction),
``transform``, ``tregion.inside``, ``loglike`` on a compacted copy, ``logl > Lmin`` -- each with
its own temporary arrays.  :func:`attach` replaces that bound method by one with the same
signature, return value, RNG consumption and counters whose middle part is ONE device pipeline
(``unb_region_refill``): the draws cross PCIe once, one flag byte and one double come back per row.

The fused path needs

* a region of this package's :class:`~ultranest_b200.mlfriends.MLFriends` whose layer can be
  applied on the device (``_fused_ok``); other regions / sampling methods still fuse everything
  after ``region.sample``;
* a device likelihood (``device_spec``; :mod:`ultranest_b200.likelihoods`);
* an identity or :class:`~ultranest_b200.transforms.ScaleShiftTransform` prior transform;
* ``tregion`` absent or a wrapping ellipsoid over all dimensions.

Anything else is delegated to the reference method, call by call, so attaching is always safe.
"""
import types

import numpy as np

from . import _native
from . import mlfriends as _ml


def _device_transform(transform, ndim):
    """``(ok, xform)`` for the sampler's prior transform."""
    spec = getattr(transform, 'device_spec', None)
    if spec is None:
        return False, None
    return True, spec(ndim)


def _device_tregion(tregion, ndim):
    """``(ok, (center, invcov, enlarge) | None)``: WrappingEllipsoid (ours or the reference's,
    mlfriends.pyx:1540-1649) with every dimension variable."""
    if tregion is None:
        return True, None
    if getattr(tregion, 'variable_dims', None) is not Ellipsis:
        return False, None
    try:
        ctr = np.asarray(tregion.ellipsoid_center, dtype=float)
        inv = np.asarray(tregion.ellipsoid_invcov, dtype=float)
        enlarge = float(tregion.enlarge)
    except (AttributeError, TypeError):
        return False, None
    if ctr.shape != (ndim,) or inv.shape != (ndim, ndim):
        return False, None
    return True, (ctr, inv, enlarge)


def refill_samples(sampler, Lmin, ndraw, nit, loglike=None, transform=None, stats=None):
    """One fused ``_refill_samples(Lmin, ndraw, nit)`` for ``sampler``; returns ``None`` when the
    configuration cannot be fused (nothing has been consumed from the RNG in that case)."""
    region = sampler.region
    loglike = sampler.loglike if loglike is None else loglike
    transform = sampler.transform if transform is None else transform
    like_spec = getattr(loglike, 'device_spec', None)
    if like_spec is None or not sampler.draw_multiple:
        return None
    if sampler.x_dim != sampler.num_params:
        return None
    # the reference's slow-sampling diagnostics (integrator.py:1809-1834) need the full arrays
    if not sampler.sampling_slow_warned and nit * ndraw >= 100000 and nit > 20:
        return None
    ndim = sampler.x_dim
    ok_x, xform = _device_transform(transform, ndim)
    ok_t, treg = _device_tregion(sampler.tregion, ndim)
    if not (ok_x and ok_t):
        return None
    kind, lparams = like_spec(ndim)

    if isinstance(region, _ml.MLFriends):
        rows, region_mode, check_cube = region._propose(ndraw)
    else:
        rows, region_mode, check_cube = region.sample(nsamples=ndraw), 0, False
    if region_mode != 0:
        eng = region._bind()
    else:
        eng = _native.get_engine()
    if len(rows) == 0:
        flags = np.empty(0, dtype=np.uint8)
        logl_all = np.empty(0)
        nu = nt = 0
    else:
        flags, logl_all, (nu, nt, _) = eng.region_refill(rows, region_mode, check_cube, xform,
                                                         treg, kind, lparams, Lmin)
        if region_mode != 0 and eng.uncertain():
            # a pair distance within the transform tolerance of the radius (mlfriends.py,
            # `_transform_tolerance`): membership with the reference's own np.dot transform, then
            # the same device tail on the members
            member = region._members_host_transform(rows, use_ellipsoid=(region_mode == 2),
                                                    check_cube=check_cube)
            flags = np.zeros(len(rows), dtype=np.uint8)
            logl_all = np.full(len(rows), -np.inf)
            nu, nt = int(member.sum()), 0
            if nu:
                f2, l2, (_, nt, _) = eng.region_refill(rows[member, :], 0, False, xform, treg, kind,
                                                       lparams, Lmin)
                flags[member] = f2 | _native.REFILL_MEMBER
                logl_all[member] = l2
    if region_mode != 0:
        region._after_sample(nu)
    if stats is not None:
        stats['fused_calls'] = stats.get('fused_calls', 0) + 1
        stats['fused_rows'] = stats.get('fused_rows', 0) + len(rows)
    sampler.ncall_region += ndraw
    if nu == 0:
        return (np.empty((0, ndim)), np.empty((0, sampler.num_params)), np.empty((0,)), 0, 0)
    member = (flags & _native.REFILL_MEMBER) != 0
    # integrator.py:1777 -- every region sample must lie strictly inside the unit cube
    if region_mode == 2:
        um = rows[member, :]
        assert np.logical_and(um > 0, um < 1).all(), (um)
    elif region_mode == 0:
        assert np.logical_and(rows > 0, rows < 1).all(), (rows)
    accepted = (flags & _native.REFILL_ACCEPTED) != 0
    u_acc = rows[accepted, :]
    v_acc = transform(u_acc) if len(u_acc) else np.empty((0, sampler.num_params))
    if v_acc is u_acc:
        v_acc = u_acc.copy()
    return u_acc, v_acc, logl_all[accepted], nt, 0


def attach(sampler, loglike=None, transform=None):
    """Install the fused refill on a ``ReactiveNestedSampler`` (or ``NestedSampler``) instance.

    ``loglike`` / ``transform`` default to the callables the sampler already holds; pass them
    when the sampler wraps yours (``vectorized=False`` or ``make_safe``).  Returns a dict that
    counts fused and delegated calls."""
    original = sampler._refill_samples
    stats = {'fused_calls': 0, 'fused_rows': 0, 'delegated_calls': 0}

    def _refill_samples(self, Lmin, ndraw, nit):
        out = refill_samples(self, Lmin, ndraw, nit, loglike=loglike, transform=transform,
                             stats=stats)
        if out is None:
            stats['delegated_calls'] += 1
            return original(Lmin, ndraw, nit)
        return out

    sampler._refill_samples = types.MethodType(_refill_samples, sampler)
    sampler._unb_refill_stats = stats
    return stats


def detach(sampler):
    """Restore the reference's ``_refill_samples``."""
    try:
        del sampler._refill_samples
    except AttributeError:
        pass
