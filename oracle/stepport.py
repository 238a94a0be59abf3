"""ctypes front-end of ``libstepfuncs_oracle.so`` (``stepfuncs_oracle.c``) -- TEST INFRASTRUCTURE ONLY.

NumPy-in / NumPy-out wrappers with the argument meaning of the reference functions they restate
(``ultranest/stepfuncs.pyx``), plus NumPy restatements of ``evolve`` (stepfuncs.pyx:189-282) and
of one inner iteration of ``PopulationSimpleSliceSampler.__next__`` (popstepsampler.py:940-965)
built from them.
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
_i64 = ctypes.c_int64


def build(force=False):
    so = os.path.join(HERE, "libstepfuncs_oracle.so")
    src = os.path.join(HERE, "stepfuncs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.sfo_within_unit_cube.argtypes = [_dp, _sz, _sz, _bp]
        L.sfo_evolve_prepare.argtypes = [_bp, _bp, _sz, _bp, _bp]
        L.sfo_evolve_update.argtypes = [_bp, _dp, _dbl, _bp, _bp, _dp, _dp, _dp, _bp, _bp, _bp, _sz]
        L.sfo_step_back.argtypes = [_dbl, _dp, _sz, _sz, _ip, _dp]
        L.sfo_update_vectorised_slice_sampler.argtypes = [
            _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _dbl, _dbl, _dp, _dp, _dp, _i64, _sz, _sz]
        L.sfo_update_vectorised_slice_sampler.restype = _i64
        _LIB = L
    return _LIB


def _f(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _b(a):
    assert a.dtype in (np.bool_, np.uint8) and a.flags.c_contiguous
    return a.ctypes.data_as(_bp)


def _i(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


def within_unit_cube(u):
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty(u.shape[0], dtype=bool)
    lib().sfo_within_unit_cube(_f(u), u.shape[0], u.shape[1], _b(out))
    return out


def evolve_prepare(searching_left, searching_right):
    search_right = np.empty_like(searching_left)
    bisecting = np.empty_like(searching_left)
    lib().sfo_evolve_prepare(_b(searching_left), _b(searching_right), len(searching_left),
                             _b(search_right), _b(bisecting))
    return search_right, bisecting


def evolve_update(acceptable, Lnew, Lmin, search_right, bisecting, currentt, current_left,
                  current_right, searching_left, searching_right, success):
    """In place, like the reference (stepfuncs.pyx:144-145)."""
    Lnew = np.ascontiguousarray(Lnew, dtype=np.float64)
    lib().sfo_evolve_update(_b(acceptable), _f(Lnew), Lmin, _b(search_right), _b(bisecting),
                            _f(currentt), _f(current_left), _f(current_right),
                            _b(searching_left), _b(searching_right), _b(success), len(acceptable))


def step_back(Lmin, allL, generation, currentt):
    lib().sfo_step_back(Lmin, _f(allL), allL.shape[0], allL.shape[1], _i(generation), _f(currentt))


def update_vectorised_slice_sampler(t, tleft, tright, proposed_L, proposed_u, proposed_p,
                                    worker_running, status, Likelihood_threshold, shrink_factor,
                                    allu, allL, allp, popsize):
    discarded = lib().sfo_update_vectorised_slice_sampler(
        _f(t), _f(tleft), _f(tright), _f(proposed_L), _f(proposed_u), _f(proposed_p),
        _i(worker_running), _i(status), Likelihood_threshold, shrink_factor, _f(allu), _f(allL),
        _f(allp), popsize, proposed_u.shape[1], proposed_p.shape[1])
    return tleft, tright, worker_running, status, allu, allL, allp, int(discarded)


_pnew_empty = np.empty((0, 1))
_Lnew_empty = np.empty(0)


def evolve(transform, loglike, Lmin, currentu, currentL, currentt, currentv, current_left,
           current_right, searching_left, searching_right, rng=np.random):
    """stepfuncs.pyx:189-282 (the reference draws from the global ``np.random``)."""
    search_right, bisecting = evolve_prepare(searching_left, searching_right)
    unew = currentu   # alias, like the reference (:252): the proposals overwrite currentu
    unew[searching_left, :] = currentu[searching_left, :] + currentv[searching_left, :] * current_left[searching_left].reshape((-1, 1))
    unew[search_right, :] = currentu[search_right, :] + currentv[search_right, :] * current_right[search_right].reshape((-1, 1))
    currentt[bisecting] = rng.uniform(current_left[bisecting], current_right[bisecting])
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


def popslice_iteration(slice_position, tleft_worker, tright_worker, tleft, tright, worker_running,
                       status, allu, allL, allp, v, transform, loglike, Lmin, shrink_factor):
    """One pass of the inner loop of PopulationSimpleSliceSampler.__next__
    (popstepsampler.py:940-965) for given uniform draws; returns the new per-worker limits and
    the number of discarded evaluations.  State arrays are updated in place."""
    popsize = len(slice_position)
    t = tleft_worker + (tright_worker - tleft_worker) * slice_position
    points = allu[worker_running, :]
    v_worker = v[worker_running, :]
    proposed_u = points + t.reshape((-1, 1)) * v_worker
    proposed_p = np.ascontiguousarray(transform(proposed_u), dtype=np.float64)
    proposed_L = np.ascontiguousarray(loglike(proposed_p), dtype=np.float64)
    *_, discarded = update_vectorised_slice_sampler(
        t, tleft, tright, proposed_L, proposed_u, proposed_p, worker_running, status, Lmin,
        shrink_factor, allu, allL, allp, popsize)
    return tleft[worker_running], tright[worker_running], discarded
