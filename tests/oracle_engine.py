"""TEST INFRASTRUCTURE: an engine with the interface of ``ultranest_b200._native.Engine`` whose
every method is answered by the CPU oracle (oracle/cport.py).  Injected into
``ultranest_b200._native`` by the CPU tests so that the HOST logic of ``ultranest_b200/mlfriends.py``
(RNG order, region/layer protocol, clustering loop, error types, the integrator drop-in) is
exercised without a GPU.  It is never importable from the product package."""
import ctypes

import numpy as np

from oracle import cport, stepport
from ultranest_b200 import _native


def _view(ptr, shape, dtype):
    """NumPy view of a caller-owned buffer passed as a raw address (what the C ABI receives)."""
    if ptr is None:
        return None
    n = int(np.prod(shape))
    if n == 0:
        return np.empty(shape, dtype=dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class OracleEngine(object):
    device = -1

    def __init__(self):
        self._live = None
        self._r2 = None
        self._layer = (_native.LAYER_IDENTITY, None, None)
        self._ell = None
        self.calls = 0

    # -- stateless ------------------------------------------------------------------------
    def find_nearby(self, apts, bpts, radiussq, out=None):
        self.calls += 1
        a, b = _native.as_f64(apts, 2), _native.as_f64(bpts, 2)
        res = np.full(len(b), -1, dtype=np.int64) if len(a) == 0 else cport.find_nearby(a, b, radiussq)
        if out is not None:
            out[:len(b)] = res
            return out
        return res

    def has_neighbour(self, apts, bpts, radiussq):
        return self.find_nearby(apts, bpts, radiussq) >= 0

    def count_nearby(self, apts, bpts, radiussq, out=None):
        self.calls += 1
        res = cport.count_nearby(apts, bpts, radiussq)
        if out is not None:
            out[:len(res)] = res
            return out
        return res

    def subtract_nearby(self, apts, radiussq, out=None):
        self.calls += 1
        res = cport.subtract_nearby(apts, radiussq)
        if out is not None:
            out[...] = res
            return out
        return res

    def compute_maxradiussq(self, apts, bpts):
        return cport.maxradiussq(apts, bpts)

    def mean_pair_distance(self, pts, clusterids):
        return cport.mean_pair_distance(pts, clusterids)

    def inside_ellipsoid(self, points, center, invcov, square_radius):
        return cport.inside_ellipsoid(points, center, invcov, square_radius)

    def transform(self, kind, inverse, pts, shift, mat):
        if kind == _native.LAYER_AFFINE:
            f = cport.untransform_affine if inverse else cport.transform_affine
            return f(pts, shift, mat)
        p = np.asarray(pts, dtype=float)
        return p * np.ravel(mat) + np.ravel(shift) if inverse else (p - np.ravel(shift)) / np.ravel(mat)

    # -- region mirror --------------------------------------------------------------------
    def region_sync_live(self, unormed):
        self._live = np.array(unormed, dtype=float, copy=True)
        return len(self._live)

    def region_set_radius(self, maxradiussq):
        self._r2 = float(maxradiussq)

    # transform tolerance: a CPU model of UNB_STAT_UNCERTAIN (pairs within tau of the radius after
    # the DEFINED-order transform; the device counts only the ones it evaluates exactly, a subset)
    _tau = 0.0
    _uncertain = 0

    def region_set_transform_tolerance(self, tau):
        self._tau = float(tau)

    def uncertain(self):
        return self._uncertain

    def _note_uncertain(self, t):
        self._uncertain = 0
        if self._tau > 0 and self._layer[0] == _native.LAYER_AFFINE and len(t):
            for j in range(0, len(t), 512):
                D = ((t[j:j + 512, None, :] - self._live[None, :, :])**2).sum(axis=2)
                self._uncertain += int((np.abs(D - self._r2) <= self._tau).any(axis=1).sum())

    def region_set_layer(self, kind, shift=None, mat=None, ndim=0):
        self._layer = (kind, None if shift is None else np.array(shift, dtype=float),
                       None if mat is None else np.array(mat, dtype=float))

    def region_set_ellipsoid(self, center, invcov, enlarge):
        self._ell = (np.array(center, dtype=float), np.array(invcov, dtype=float), float(enlarge))

    def _xf(self, pts):
        kind, shift, mat = self._layer
        if kind == _native.LAYER_AFFINE:
            return cport.transform_affine(pts, shift, mat)      # the DEFINED order
        if kind == _native.LAYER_SCALING:
            return (pts - shift) / mat
        return pts

    def region_inside(self, pts, want_index=False, use_ellipsoid=True):
        self.calls += 1
        pts = _native.as_f64(pts, 2)
        mask = np.ones(len(pts), dtype=bool)
        if use_ellipsoid:
            mask = cport.inside_ellipsoid(pts, *self._ell)
        idx = np.full(len(pts), -1, dtype=np.int64)
        self._uncertain = 0
        if mask.any():
            t = self._xf(pts[mask])
            self._note_uncertain(t)
            idx[mask] = cport.find_nearby(self._live, t, self._r2)
        mask = idx >= 0
        return (mask, idx) if want_index else mask

    def region_inside_ellipsoid(self, pts):
        self.calls += 1
        return cport.inside_ellipsoid(_native.as_f64(pts, 2), *self._ell)

    def region_find_nearby(self, tpts):
        self.calls += 1
        return cport.find_nearby(self._live, tpts, self._r2)

    def region_has_neighbour(self, tpts):
        return self.region_find_nearby(tpts) >= 0

    def region_count_nearby(self, tpts):
        self.calls += 1
        return cport.count_nearby(self._live, tpts, self._r2)

    def region_bootstrap_moments(self, u, selected, c0, round_lo=0, round_hi=None):
        """CPU model of unb_region_bootstrap_moments (any summation order will do: it only screens)."""
        self.calls += 1
        u = np.asarray(u, dtype=float)
        selected = np.asarray(selected, dtype=bool)
        nrounds, d = len(selected), u.shape[1]
        round_hi = nrounds if round_hi is None else round_hi
        sums, sxx = np.zeros((nrounds, d)), np.zeros((nrounds, d, d))
        counts = np.zeros(nrounds, dtype=np.int64)
        for r in range(round_lo, round_hi):
            y = u[selected[r]] - c0
            counts[r] = len(y)
            sums[r] = y.sum(axis=0)
            sxx[r] = np.triu(np.dot(y.T, y))
        return counts, sums, sxx

    def region_bootstrap(self, unormed, selected, u=None, ctrs=None, invcovs=None,
                         round_lo=0, round_hi=None):
        self.calls += 1
        selected = np.asarray(selected, dtype=bool)
        nrounds = len(selected)
        round_hi = nrounds if round_hi is None else round_hi
        maxd = np.zeros(nrounds) if unormed is not None else None
        f = np.zeros(nrounds) if u is not None else None
        for r in range(round_lo, round_hi):
            sel = selected[r]
            if sel.all() or not sel.any():
                continue
            if maxd is not None:
                maxd[r] = cport.maxradiussq_selected(unormed, sel)
            if f is not None:
                f[r] = cport.enlargement_f(u, sel, ctrs[r], invcovs[r])
        return maxd, f

    def region_inside_loglike(self, pts, kind, lparams=None, mask_out=None, like_out=None):
        mask = self.region_inside(pts)
        like = np.full(len(pts), -np.inf)
        if kind == _native.LOGLIKE_GAUSS:
            d = pts.shape[1]
            like[mask] = cport.loglike_gauss(pts[mask], lparams[:d], lparams[d])
        elif kind == _native.LOGLIKE_ROSENBROCK:
            like[mask] = cport.loglike_rosenbrock(pts[mask])
        else:
            like[mask] = cport.loglike_eggbox(pts[mask])
        return mask, like

    def _like(self, v, kind, lparams):
        if kind == _native.LOGLIKE_GAUSS:
            d = v.shape[1]
            return cport.loglike_gauss(v, lparams[:d], lparams[d])
        if kind == _native.LOGLIKE_ROSENBROCK:
            return cport.loglike_rosenbrock(v)
        return cport.loglike_eggbox(v)

    def region_refill(self, u, region_mode, check_cube, xform, tregion, like_kind, lparams, Lmin):
        """Stage by stage what integrator.py:1773-1805 does, with the oracle's pieces."""
        self.calls += 1
        self._uncertain = 0
        u = _native.as_f64(u, 2)
        n = len(u)
        if region_mode == 2:
            member = self.region_inside(u)
        elif region_mode == 1:
            member = self.region_inside(u, use_ellipsoid=False)
        else:
            member = np.ones(n, dtype=bool)
        if check_cube:
            member &= np.logical_and(u > 0, u < 1).all(axis=1)
        v = u if xform is None else u * xform[0] + xform[1]
        tpass = member.copy()
        if tregion is not None and member.any():
            tpass[member] = cport.inside_ellipsoid(np.ascontiguousarray(v[member]), *tregion)
        like = np.full(n, -np.inf)
        if tpass.any():
            like[tpass] = self._like(np.ascontiguousarray(v[tpass]), like_kind, lparams)
        acc = like > Lmin
        flags = (member * _native.REFILL_MEMBER + tpass * _native.REFILL_TREGION
                 + acc * _native.REFILL_ACCEPTED).astype(np.uint8)
        return flags, like, (int(member.sum()), int(tpass.sum()), int(acc.sum()))

    # -- likelihoods ----------------------------------------------------------------------
    def loglike_gauss(self, theta, centers, sigma, norm_const):
        return cport.loglike_gauss(theta, centers, sigma)

    def loglike_rosenbrock(self, theta):
        return cport.loglike_rosenbrock(theta)

    def loglike_eggbox(self, z):
        return cport.loglike_eggbox(z)

    # -- population step-sampler helpers: the C-ABI calls of ultranest_b200/stepfuncs.py and
    #    popstepsampler.py, answered by oracle/stepport.py ----------------------------------
    def call(self, name, *args):
        self.calls += 1
        return getattr(self, "_" + name)(*args)

    def _desc(self, addr, ndim):
        d = ctypes.cast(int(addr), ctypes.POINTER(_native.StepDesc)).contents
        xform = None
        if d.xform_kind == _native.XFORM_SCALE_SHIFT:
            xform = (_view(d.xform_scale, (ndim,), np.float64).copy(), _view(d.xform_lo, (ndim,), np.float64).copy())
        lparams = _view(d.lparams, (ndim + 2,), np.float64).copy() if d.lparams else None
        kind = int(d.loglike_kind)
        transform = (lambda u: u) if xform is None else (lambda u: u * xform[0] + xform[1])
        return transform, (lambda v: self._like(np.ascontiguousarray(v), kind, lparams))

    def _unb_within_unit_cube(self, u, n, d, acceptable):
        _view(acceptable, (n,), bool)[:] = stepport.within_unit_cube(_view(u, (n, d), np.float64))

    def _unb_evolve_prepare(self, sl, sr, n, search_right, bisecting):
        a, b = stepport.evolve_prepare(_view(sl, (n,), bool), _view(sr, (n,), bool))
        _view(search_right, (n,), bool)[:] = a
        _view(bisecting, (n,), bool)[:] = b

    def _unb_evolve_update(self, acceptable, Lnew, n_lnew, Lmin, search_right, bisecting, currentt,
                           current_left, current_right, searching_left, searching_right, success, n):
        f, b = np.float64, bool
        stepport.evolve_update(_view(acceptable, (n,), b), _view(Lnew, (n_lnew,), f), Lmin,
                               _view(search_right, (n,), b), _view(bisecting, (n,), b),
                               _view(currentt, (n,), f), _view(current_left, (n,), f),
                               _view(current_right, (n,), f), _view(searching_left, (n,), b),
                               _view(searching_right, (n,), b), _view(success, (n,), b))

    def _unb_evolve(self, desc, Lmin, currentu, currentv, currentt, current_left, current_right,
                    searching_left, searching_right, n, ndim, acceptable, success, like):
        """unb_evolve = stepfuncs.pyx:250-274 with the bisecting walkers' draws already in
        currentt; restated from the oracle's pieces."""
        f, b = np.float64, bool
        transform, loglike = self._desc(desc, ndim)
        u, v = _view(currentu, (n, ndim), f), _view(currentv, (n, ndim), f)
        t, cl, cr = _view(currentt, (n,), f), _view(current_left, (n,), f), _view(current_right, (n,), f)
        sl, sr = _view(searching_left, (n,), b), _view(searching_right, (n,), b)
        search_right, bisecting = stepport.evolve_prepare(sl, sr)
        coef = np.where(sl, cl, np.where(search_right, cr, t))
        u[:] = u + v * coef.reshape((-1, 1))
        acc = stepport.within_unit_cube(u)
        L = np.full(n, -np.inf)
        if acc.any():
            L[acc] = loglike(transform(u[acc, :]))
        succ = np.zeros(n, dtype=bool)
        stepport.evolve_update(acc, L[acc], Lmin, search_right, bisecting, t, cl, cr, sl, sr, succ)
        _view(acceptable, (n,), b)[:] = acc
        _view(success, (n,), b)[:] = succ
        _view(like, (n,), f)[:] = L

    def _unb_step_back(self, Lmin, allL, nwalkers, ncols, generation, currentt):
        if ncols > 2048:
            raise ValueError("step_back supports chains of up to 2048 generations")
        stepport.step_back(Lmin, _view(allL, (nwalkers, ncols), np.float64),
                           _view(generation, (nwalkers,), np.int64), _view(currentt, (nwalkers,), np.float64))

    def _unb_update_vectorised_slice_sampler(self, t, tleft, tright, pL, pu, pp, worker_running, status,
                                             thr, shrink, allu, allL, allp, popsize, ndim, nparams,
                                             discarded):
        f, i = np.float64, np.int64
        w = _view(worker_running, (popsize,), i)
        if ((w < 0) | (w >= popsize)).any():
            raise ValueError("worker_running outside the population")
        out = stepport.update_vectorised_slice_sampler(
            _view(t, (popsize,), f), _view(tleft, (popsize,), f), _view(tright, (popsize,), f),
            _view(pL, (popsize,), f), _view(pu, (popsize, ndim), f), _view(pp, (popsize, nparams), f), w,
            _view(status, (popsize,), i), thr, shrink, _view(allu, (popsize, ndim), f),
            _view(allL, (popsize,), f), _view(allp, (popsize, nparams), f), popsize)
        _view(discarded, (1,), i)[0] = out[-1]

    def _unb_popslice_begin(self, desc, allu, allL, v, tleft, tright, popsize, ndim, thr, shrink):
        f = np.float64
        transform, loglike = self._desc(desc, ndim)
        self._ps = dict(transform=transform, loglike=loglike, thr=thr, shrink=shrink,
                        allu=_view(allu, (popsize, ndim), f).copy(), allL=_view(allL, (popsize,), f).copy(),
                        v=_view(v, (popsize, ndim), f).copy(), tleft=_view(tleft, (popsize,), f).copy(),
                        tright=_view(tright, (popsize,), f).copy(), allp=np.full((popsize, ndim), np.nan),
                        worker=np.arange(popsize, dtype=np.int64), status=np.zeros(popsize, dtype=np.int64))
        self._ps["tlw"], self._ps["trw"] = self._ps["tleft"].copy(), self._ps["tright"].copy()

    def _unb_popslice_iterate(self, slice_position, n_running, discarded):
        s = self._ps
        n = len(s["allL"])
        pos = _view(slice_position, (n,), np.float64)
        s["tlw"], s["trw"], disc = stepport.popslice_iteration(
            pos, s["tlw"], s["trw"], s["tleft"], s["tright"], s["worker"], s["status"], s["allu"], s["allL"],
            s["allp"], s["v"], s["transform"], s["loglike"], s["thr"], s["shrink"])
        n_running._obj.value = int((s["status"] == 0).sum())
        discarded._obj.value = int(disc)

    def _unb_popslice_end(self, allu, allp, allL, tleft, tright, status):
        s = self._ps
        n, d = s["allu"].shape
        for ptr, key, shape, dt in ((allu, "allu", (n, d), np.float64), (allp, "allp", (n, d), np.float64),
                                    (allL, "allL", (n,), np.float64), (tleft, "tleft", (n,), np.float64),
                                    (tright, "tright", (n,), np.float64), (status, "status", (n,), np.int64)):
            if ptr is not None:
                _view(ptr, shape, dt)[...] = s[key]

