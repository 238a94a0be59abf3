"""Prior transforms the fused refill pipeline can evaluate on the device.

``ReactiveNestedSampler(param_names, loglike, transform, vectorized=True)`` calls
``transform(u)`` on every batch of region samples (integrator.py:1790).  The classes here are
plain NumPy callables with that contract; ``device_spec`` tells
:func:`ultranest_b200.refill.attach` how to repeat exactly the same arithmetic inside the device
pipeline (two roundings per element, no fused multiply-add), so the likelihood the device
evaluates is the likelihood of the very ``v`` the host would have produced.
"""
import numpy as np


class IdentityTransform(object):
    """``v = u`` (what the integrator installs for ``transform=None``, integrator.py:1337-1338)."""

    def __call__(self, u):
        return u

    def device_spec(self, ndim):
        return None


class ScaleShiftTransform(object):
    """Independent uniform priors: ``v = u * (hi - lo) + lo`` -- the transform of the reference's
    examples (e.g. examples/testsine.py, docs ``cube * (hi - lo) + lo``)."""

    def __init__(self, lo, hi):
        self.lo = np.atleast_1d(np.asarray(lo, dtype=float))
        self.hi = np.atleast_1d(np.asarray(hi, dtype=float))
        self.scale = self.hi - self.lo

    def __call__(self, u):
        return np.asarray(u, dtype=float) * self.scale + self.lo

    def device_spec(self, ndim):
        return (np.ascontiguousarray(np.broadcast_to(self.scale, (ndim,)), dtype=float),
                np.ascontiguousarray(np.broadcast_to(self.lo, (ndim,)), dtype=float))
