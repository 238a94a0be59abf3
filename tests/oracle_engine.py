"""TEST INFRASTRUCTURE: an engine with the interface of ``ultranest_b200._native.Engine`` whose
every method is answered by the CPU oracle (oracle/cport.py).  Injected into
``ultranest_b200._native`` by the CPU tests so that the HOST logic of ``ultranest_b200/mlfriends.py``
(RNG order, region/layer protocol, clustering loop, error types, the integrator drop-in) is
exercised without a GPU.  It is never importable from the product package."""
import numpy as np

from oracle import cport
from ultranest_b200 import _native


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
        if mask.any():
            idx[mask] = cport.find_nearby(self._live, self._xf(pts[mask]), self._r2)
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
