"""Vectorised likelihood batch calls on the device.

Each class is a plain callable ``f(ndarray[n, d]) -> ndarray[n]`` -- the contract
``ReactiveNestedSampler(..., vectorized=True)`` expects (integrator.py:1789, 1802) -- backed by a
CUDA kernel with the ``(params, d, n, like)`` shape of the reference's C convention
(``languages/c/mylib.c:33``).  ``device_spec`` lets :meth:`MLFriends.inside_and_loglike` fuse the
evaluation with the membership test so the proposals cross PCIe once.
"""
import numpy as np

from . import _native


class GaussianLogLike(object):
    """``-0.5 * (((theta - centers)/sigma)**2).sum(axis=1) - 0.5*log(2*pi*sigma**2)*ndim``
    (docs/gauss.py:25-27, examples/testgauss.py:13-15), with NumPy's pairwise summation order
    reproduced: bit-identical to the NumPy expression."""

    def __init__(self, centers, sigma):
        self.centers = np.atleast_1d(np.asarray(centers, dtype=float))
        self.sigma = float(sigma)

    def norm_const(self, ndim):
        return 0.5 * np.log(2 * np.pi * self.sigma**2) * ndim

    def _centers(self, ndim):
        return np.ascontiguousarray(np.broadcast_to(self.centers, (ndim,)), dtype=float)

    def __call__(self, theta):
        theta = np.asarray(theta, dtype=float)
        ndim = theta.shape[1]
        return _native.get_engine().loglike_gauss(theta, self._centers(ndim), self.sigma,
                                                  self.norm_const(ndim))

    def device_spec(self, ndim):
        lparams = np.concatenate([self._centers(ndim), [self.sigma, self.norm_const(ndim)]])
        return _native.LOGLIKE_GAUSS, lparams


class RosenbrockLogLike(object):
    """``-2 * (100 * (b - a**2)**2 + (1 - a)**2).sum(axis=1)`` with ``a = theta[:, :-1]``,
    ``b = theta[:, 1:]`` (examples/testrosenbrock.py:10-13).  Bit-identical to NumPy."""

    def __call__(self, theta):
        return _native.get_engine().loglike_rosenbrock(np.asarray(theta, dtype=float))

    def device_spec(self, ndim):
        return _native.LOGLIKE_ROSENBROCK, None


class EggboxLogLike(object):
    """``(2 + cos(z / 2).prod(axis=1))**5`` (examples/testeggbox.py:9-11).  ``cos``/``pow`` are
    the CUDA math library's (<= 2 ulp), so parity with NumPy is ~1e-15 relative, not bitwise."""

    def __call__(self, z):
        return _native.get_engine().loglike_eggbox(np.asarray(z, dtype=float))

    def device_spec(self, ndim):
        return _native.LOGLIKE_EGGBOX, None
