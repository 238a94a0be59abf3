"""ctypes front-end of ``liboracle.so`` (``mlfriends_oracle.c``) -- TEST INFRASTRUCTURE ONLY.

NumPy-in / NumPy-out wrappers with the argument meaning of the reference functions
they restate (``ultranest/mlfriends.pyx``), plus NumPy restatements of the host-side
orchestration (bootstrap rounds, ``MLFriends.inside``) built from them.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)
_bp = ctypes.POINTER(ctypes.c_uint8)
_sz = ctypes.c_size_t
_dbl = ctypes.c_double


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "mlfriends_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_count_nearby.argtypes = [_dp, _sz, _dp, _sz, _sz, _dbl, _ip]
        L.orc_find_nearby.argtypes = [_dp, _sz, _dp, _sz, _sz, _dbl, _ip]
        L.orc_subtract_nearby.argtypes = [_dp, _sz, _sz, _dbl, _dp]
        L.orc_maxradiussq.argtypes = [_dp, _sz, _dp, _sz, _sz]
        L.orc_maxradiussq.restype = _dbl
        L.orc_maxradiussq_double.argtypes = [_dp, _sz, _dp, _sz, _sz]
        L.orc_maxradiussq_double.restype = _dbl
        L.orc_maxradiussq_selected.argtypes = [_dp, _sz, _sz, _bp]
        L.orc_maxradiussq_selected.restype = _dbl
        L.orc_mean_pair_distance.argtypes = [_dp, _ip, _sz, _sz]
        L.orc_mean_pair_distance.restype = _dbl
        L.orc_inside_ellipsoid.argtypes = [_dp, _sz, _sz, _dp, _dp, _dbl, _bp, _dp]
        L.orc_transform_scaling.argtypes = [_dp, _sz, _sz, _dp, _dp, _dp]
        L.orc_transform_affine.argtypes = [_dp, _sz, _sz, _dp, _dp, _dp]
        L.orc_untransform_affine.argtypes = [_dp, _sz, _sz, _dp, _dp, _dp]
        L.orc_np_pairwise_sum.argtypes = [_dp, _sz]
        L.orc_np_pairwise_sum.restype = _dbl
        L.orc_loglike_gauss.argtypes = [_dp, _sz, _sz, _dp, _dp, _dbl, _dbl]
        L.orc_loglike_rosenbrock.argtypes = [_dp, _sz, _sz, _dp]
        L.orc_loglike_eggbox.argtypes = [_dp, _sz, _sz, _dp]
        L.orc_enlargement_f.argtypes = [_dp, _sz, _sz, _bp, _dp, _dp]
        L.orc_enlargement_f.restype = _dbl
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def find_nearby(apts, bpts, radiussq, nnearby=None):
    a, pa = _d(apts)
    b, pb = _d(bpts)
    if nnearby is None:
        nnearby = np.empty(len(b), dtype=np.int64)
    lib().orc_find_nearby(pa, len(a), pb, len(b), a.shape[1], radiussq,
                          nnearby.ctypes.data_as(_ip))
    return nnearby


def count_nearby(apts, bpts, radiussq):
    a, pa = _d(apts)
    b, pb = _d(bpts)
    out = np.empty(len(b), dtype=np.int64)
    lib().orc_count_nearby(pa, len(a), pb, len(b), a.shape[1], radiussq,
                           out.ctypes.data_as(_ip))
    return out


def subtract_nearby(upoints, maxradiussq):
    a, pa = _d(upoints)
    out = np.zeros_like(a)
    lib().orc_subtract_nearby(pa, len(a), a.shape[1], maxradiussq, out.ctypes.data_as(_dp))
    return out


def maxradiussq(apts, bpts, as_float32=True):
    a, pa = _d(apts)
    b, pb = _d(bpts)
    f = lib().orc_maxradiussq if as_float32 else lib().orc_maxradiussq_double
    return f(pa, len(a), pb, len(b), a.shape[1])


def maxradiussq_selected(pts, selected):
    a, pa = _d(pts)
    s = np.ascontiguousarray(selected, dtype=np.uint8)
    return lib().orc_maxradiussq_selected(pa, len(a), a.shape[1], s.ctypes.data_as(_bp))


def mean_pair_distance(pts, clusterids):
    a, pa = _d(pts)
    c = np.ascontiguousarray(clusterids, dtype=np.int64)
    return lib().orc_mean_pair_distance(pa, c.ctypes.data_as(_ip), len(a), a.shape[1])


def inside_ellipsoid(points, center, invcov, square_radius, return_r=False):
    p, pp = _d(points)
    c, pc = _d(center)
    A, pA = _d(invcov)
    mask = np.empty(len(p), dtype=np.uint8)
    r = np.empty(len(p)) if return_r else None
    lib().orc_inside_ellipsoid(pp, len(p), p.shape[1], pc, pA, square_radius,
                               mask.ctypes.data_as(_bp),
                               r.ctypes.data_as(_dp) if return_r else None)
    return (mask.astype(bool), r) if return_r else mask.astype(bool)


def transform_scaling(w, mean, std):
    p, pp = _d(np.atleast_2d(w))
    m, pm = _d(np.ravel(mean))
    s, ps = _d(np.ravel(std))
    out = np.empty_like(p)
    lib().orc_transform_scaling(pp, len(p), p.shape[1], pm, ps, out.ctypes.data_as(_dp))
    return out.reshape(np.shape(w))


def transform_affine(w, ctr, T):
    p, pp = _d(np.atleast_2d(w))
    c, pc = _d(ctr)
    t, pt = _d(T)
    out = np.empty_like(p)
    lib().orc_transform_affine(pp, len(p), p.shape[1], pc, pt, out.ctypes.data_as(_dp))
    return out.reshape(np.shape(w))


def untransform_affine(ww, ctr, invT):
    p, pp = _d(np.atleast_2d(ww))
    c, pc = _d(ctr)
    t, pt = _d(invT)
    out = np.empty_like(p)
    lib().orc_untransform_affine(pp, len(p), p.shape[1], pc, pt, out.ctypes.data_as(_dp))
    return out.reshape(np.shape(ww))


def np_pairwise_sum(a):
    a, pa = _d(a)
    return lib().orc_np_pairwise_sum(pa, a.size)


def gauss_norm_const(sigma, ndim):
    """``0.5 * np.log(2 * np.pi * sigma**2) * ndim`` exactly as docs/gauss.py:26 evaluates it."""
    return 0.5 * np.log(2 * np.pi * sigma**2) * ndim


def loglike_gauss(theta, centers, sigma):
    p, pp = _d(theta)
    c, pc = _d(np.broadcast_to(centers, (p.shape[1],)))
    out = np.empty(len(p))
    lib().orc_loglike_gauss(pp, p.shape[1], len(p), out.ctypes.data_as(_dp), pc, sigma,
                            gauss_norm_const(sigma, p.shape[1]))
    return out


def loglike_rosenbrock(theta):
    p, pp = _d(theta)
    out = np.empty(len(p))
    lib().orc_loglike_rosenbrock(pp, p.shape[1], len(p), out.ctypes.data_as(_dp))
    return out


def loglike_eggbox(z):
    p, pp = _d(z)
    out = np.empty(len(p))
    lib().orc_loglike_eggbox(pp, p.shape[1], len(p), out.ctypes.data_as(_dp))
    return out


def enlargement_f(u, selected, ctr, a):
    p, pp = _d(u)
    c, pc = _d(ctr)
    A, pA = _d(a)
    s = np.ascontiguousarray(selected, dtype=np.uint8)
    return lib().orc_enlargement_f(pp, len(p), p.shape[1], s.ctypes.data_as(_bp), pc, pA)


# --------------------------------------------------------------------------------------
# NumPy restatements of the host-side orchestration around the C loops
# --------------------------------------------------------------------------------------

def bounding_ellipsoid(x, minvol=0.):
    """mlfriends.pyx:426-476 (minvol == 0 branch only; the eigenvalue lift is LAPACK
    host code the product re-uses unchanged in spirit)."""
    ndim = x.shape[1]
    ctr = np.mean(x, axis=0)
    cov = np.atleast_2d(np.cov(x - ctr, rowvar=0)) * (ndim + 2)
    assert minvol == 0.
    return ctr, cov


def draw_selection(rng, n):
    """One bootstrap round's selection mask, mlfriends.pyx:1045-1047."""
    idx = rng.randint(n, size=n)
    sel = np.zeros(n, dtype=bool)
    sel[idx] = True
    return sel


def compute_enlargement(u, unormed, nbootstraps, rng):
    """MLFriends.compute_enlargement, mlfriends.pyx:1017-1070 (minvol=0)."""
    n = len(u)
    maxd = 0.0
    maxf = 0.0
    for _ in range(nbootstraps):
        sel = draw_selection(rng, n)
        if sel.all() or not sel.any():
            continue
        maxd = max(maxd, maxradiussq_selected(unormed, sel))
        ctr, cov = bounding_ellipsoid(u[sel])
        a = np.linalg.inv(cov)
        f = enlargement_f(u, sel, ctr, a)
        if not f > 0:
            raise np.linalg.LinAlgError("Distances are not positive")
        maxf = max(maxf, f)
    assert maxd > 0 and maxf > 0
    return maxd, maxf


def region_inside(pts, unormed, transform, maxradiussq_, center, invcov, enlarge):
    """MLFriends.inside, mlfriends.pyx:1186-1211, given a ``transform`` callable."""
    mask = inside_ellipsoid(pts, center, invcov, enlarge)
    if mask.any():
        bpts = transform(pts[mask, :])
        idnearby = find_nearby(unormed, bpts, maxradiussq_)
        mask[mask] = idnearby >= 0
    return mask
