"""ctypes binding of ``libultranest_b200.so`` (the C ABI in ``include/ultranest_b200.h``).

There is deliberately no CPU fallback: if the CUDA library cannot be loaded, or no sm_100
device is visible, importing the engine raises.  PyTorch is not needed here; it is only the
container for device-resident arrays in :mod:`ultranest_b200.device`.
"""
import ctypes
import os
import threading

import numpy as np

from . import build as _build

UNB_OK = 0
UNB_ERR_CUDA = -1
UNB_ERR_ARG = -2
UNB_ERR_STATE = -3
UNB_ERR_NOMEM = -4
UNB_ERR_NUMERIC = -5

OPT_EXACT_ONLY = 1
OPT_CHUNK_ROWS = 2
OPT_FILTER_FP32 = 3
OPT_SURE_LEVEL = 4
OPT_COOP_MAX = 5
OPT_BLOCK_KERNEL = 6
OPT_BIN_MIN_ROWS = 7
STAT_KERNEL_LAUNCHES = 1
STAT_RECHECKS = 2
STAT_H2D_BYTES = 3
STAT_D2H_BYTES = 4
STAT_TILE_VISITS = 5
STAT_UNCERTAIN = 6

LAYER_IDENTITY = 0
LAYER_SCALING = 1
LAYER_AFFINE = 2

LOGLIKE_NONE = 0
LOGLIKE_GAUSS = 1
LOGLIKE_EGGBOX = 2
LOGLIKE_ROSENBROCK = 3

SAMPLE_WRAPPING_ELLIPSOID = 0
SAMPLE_UNIT_CUBE = 1

XFORM_IDENTITY = 0
XFORM_SCALE_SHIFT = 1
REFILL_MEMBER, REFILL_TREGION, REFILL_ACCEPTED = 1, 2, 4

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_ip = ctypes.POINTER(ctypes.c_int64)
_c_bp = ctypes.POINTER(ctypes.c_uint8)
_c_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_dbl = ctypes.c_double
_int = ctypes.c_int
_i64 = ctypes.c_int64

# name -> argtypes (after the leading ctx pointer); every symbol of include/ultranest_b200.h
SIGNATURES = {
    "unb_ctx_destroy": [],
    "unb_ctx_set_option": [_int, _i64],
    "unb_ctx_get_stat": [_int, _c_ip],
    "unb_ctx_synchronize": [],
    "unb_fp64_peak": [_c_dp],
    "unb_fp32_peak": [_c_dp],
    "unb_fp32_peak_form": [_int, _c_dp],
    "unb_find_nearby": [_c_vp, _sz, _c_vp, _sz, _sz, _dbl, _c_vp],
    "unb_count_nearby": [_c_vp, _sz, _c_vp, _sz, _sz, _dbl, _c_vp],
    "unb_has_neighbour": [_c_vp, _sz, _c_vp, _sz, _sz, _dbl, _c_vp],
    "unb_subtract_nearby": [_c_vp, _sz, _sz, _dbl, _c_vp],
    "unb_compute_maxradiussq": [_c_vp, _sz, _c_vp, _sz, _sz, _c_dp],
    "unb_mean_pair_distance": [_c_vp, _c_vp, _sz, _sz, _c_dp],
    "unb_inside_ellipsoid": [_c_vp, _sz, _sz, _c_vp, _c_vp, _dbl, _c_vp],
    "unb_transform_scaling": [_c_vp, _sz, _sz, _c_vp, _c_vp, _c_vp],
    "unb_untransform_scaling": [_c_vp, _sz, _sz, _c_vp, _c_vp, _c_vp],
    "unb_transform_affine": [_c_vp, _sz, _sz, _c_vp, _c_vp, _c_vp],
    "unb_untransform_affine": [_c_vp, _sz, _sz, _c_vp, _c_vp, _c_vp],
    "unb_region_sync_live": [_c_vp, _sz, _sz, _c_ip],
    "unb_region_set_layer": [_int, _c_vp, _c_vp, _sz],
    "unb_region_set_ellipsoid": [_c_vp, _c_vp, _dbl, _sz],
    "unb_region_set_radius": [_dbl],
    "unb_region_set_transform_tolerance": [_dbl],
    "unb_region_inside": [_c_vp, _sz, _c_vp, _c_vp],
    "unb_region_friends": [_c_vp, _sz, _c_vp, _c_vp],
    "unb_region_inside_ellipsoid": [_c_vp, _sz, _c_vp],
    "unb_region_inside_ellipsoid_dev": [_c_vp, _sz, _c_vp, _c_vp],
    "unb_region_inside_dev": [_c_vp, _sz, _c_vp, _c_vp],
    "unb_region_find_nearby": [_c_vp, _sz, _c_vp],
    "unb_region_count_nearby": [_c_vp, _sz, _c_vp],
    "unb_region_has_neighbour": [_c_vp, _sz, _c_vp],
    "unb_region_find_nearby_dev": [_c_vp, _sz, _c_vp, _c_vp, _c_vp],
    "unb_region_bootstrap": [_c_vp, _c_vp, _sz, _sz, _c_vp, _sz, _sz, _sz, _c_vp, _c_vp,
                             _c_vp, _c_vp],
    "unb_region_bootstrap_moments": [_c_vp, _sz, _sz, _c_vp, _sz, _sz, _sz, _c_vp, _c_vp, _c_vp, _c_vp],
    "unb_region_bootstrap_fold_dev": [_c_vp, _c_vp, _sz, _sz, _c_vp, _sz, _sz, _sz, _c_vp, _c_vp,
                                      _int, _dbl, _c_vp, _c_vp],
    "unb_loglike_gauss": [_c_vp, _sz, _sz, _c_vp, _c_vp, _dbl, _dbl],
    "unb_loglike_rosenbrock": [_c_vp, _sz, _sz, _c_vp],
    "unb_loglike_eggbox": [_c_vp, _sz, _sz, _c_vp],
    "unb_loglike_gauss_dev": [_c_vp, _sz, _sz, _c_vp, _c_vp, _c_vp, _dbl, _dbl, _c_vp],
    "unb_region_inside_loglike": [_c_vp, _sz, _c_vp, _c_vp, _int, _c_vp],
    "unb_region_inside_loglike_dev": [_c_vp, _sz, _c_vp, _c_vp, _int, _c_vp, _c_vp],
    "unb_region_refill": [_c_vp, _sz, _sz, _c_vp, _c_vp, _c_vp, _c_vp],
    "unb_region_sample": [_c_vp, _sz, _c_vp, _c_vp, _c_ip, _c_vp],
    "unb_region_sample_dev": [_c_vp, _sz, _c_vp, _c_vp, _c_vp, _c_vp],
    "unb_sample_draw": [_int, _sz, _sz, ctypes.c_uint64, ctypes.c_uint64, _c_vp, _c_vp, _dbl, _c_vp,
                        _c_vp],
    # population step-sampler helpers (ultranest/stepfuncs.pyx)
    "unb_within_unit_cube": [_c_vp, _sz, _sz, _c_vp],
    "unb_evolve_prepare": [_c_vp, _c_vp, _sz, _c_vp, _c_vp],
    "unb_evolve_update": [_c_vp, _c_vp, _sz, _dbl, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                          _c_vp, _sz],
    "unb_evolve": [_c_vp, _dbl, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _sz, _sz, _c_vp,
                   _c_vp, _c_vp],
    "unb_step_back": [_dbl, _c_vp, _sz, _sz, _c_vp, _c_vp],
    "unb_update_vectorised_slice_sampler": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                                            _dbl, _dbl, _c_vp, _c_vp, _c_vp, _sz, _sz, _sz, _c_vp],
    "unb_popslice_begin": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _sz, _sz, _dbl, _dbl],
    "unb_popslice_iterate": [_c_vp, _c_vp, _c_vp],
    "unb_popslice_end": [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp],
}
# symbols without the leading ctx argument
FREE_SIGNATURES = {
    "unb_abi_version": ([], _int),
    "unb_ctx_create": ([_int, ctypes.POINTER(_c_vp)], _int),
    "unb_last_error": ([_c_vp], ctypes.c_char_p),
}

_lib = None
_lib_lock = threading.Lock()


class NativeLibraryError(ImportError):
    """The CUDA library is missing or cannot be loaded (there is no CPU fallback)."""


def library_path():
    return _build.LIBPATH


def load_library():
    """dlopen the in-tree library (building it with nvcc when missing) and set signatures."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = _build.LIBPATH
        if not os.path.exists(path):
            try:
                _build.build()
            except Exception as exc:  # noqa: BLE001
                raise NativeLibraryError(
                    "libultranest_b200.so is not built and nvcc could not build it (%s); "
                    "run `python -m ultranest_b200.build`" % exc)
        try:
            lib = ctypes.CDLL(path)
        except OSError as exc:
            raise NativeLibraryError("cannot load %s: %s" % (path, exc))
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = [_c_vp] + list(args)
            fn.restype = _int
        for name, (args, res) in FREE_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = list(args)
            fn.restype = res
        _lib = lib
        return lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def as_f64(a, ndim=None):
    """C-contiguous float64 view/copy (callers pass boolean-masked copies and slices)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    if ndim is not None and a.ndim != ndim:
        raise ValueError("expected a %d-d array, got shape %s" % (ndim, a.shape))
    return a


class Engine(object):
    """One ``unb_ctx`` (one CUDA device, one host thread)."""

    def __init__(self, device=None):
        self.lib = load_library()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", os.environ.get("UNB_DEVICE", "0")))
        handle = _c_vp()
        rc = self.lib.unb_ctx_create(int(device), ctypes.byref(handle))
        if rc != UNB_OK or not handle.value:
            raise RuntimeError(
                "ultranest_b200: cannot create a CUDA context on device %d (rc=%d). "
                "This package needs an sm_100 (B200) GPU; there is no CPU fallback." % (device, rc))
        self.device = int(device)
        self.ctx = handle
        self._bound = None      # id of the region object whose state the ctx mirrors

    # -- plumbing --------------------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.unb_ctx_destroy(self.ctx)
            self.ctx = _c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def error(self):
        msg = self.lib.unb_last_error(self.ctx)
        return msg.decode("utf-8", "replace") if msg else ""

    def check(self, rc):
        if rc == UNB_OK:
            return
        msg = "ultranest_b200: %s" % self.error()
        if rc == UNB_ERR_NUMERIC:
            raise np.linalg.LinAlgError(msg)
        if rc == UNB_ERR_ARG:
            raise ValueError(msg)
        if rc == UNB_ERR_NOMEM:
            raise MemoryError(msg)
        raise RuntimeError(msg + " (rc=%d)" % rc)

    def call(self, name, *args):
        self.check(getattr(self.lib, name)(self.ctx, *args))

    def set_option(self, option, value):
        self.call("unb_ctx_set_option", option, int(value))

    def stat(self, key):
        v = _i64(0)
        self.call("unb_ctx_get_stat", key, ctypes.byref(v))
        return int(v.value)

    def synchronize(self):
        self.call("unb_ctx_synchronize")

    def fp32_peak(self):
        """Measured fp32 FMA rate of this device (lane-FMAs per second)."""
        v = _dbl(0.0)
        self.call("unb_fp32_peak", ctypes.byref(v))
        return float(v.value)

    def fp32_peak_form(self, form):
        """The fp32 probe by instruction form: 0 scalar FFMA, 1 packed FFMA2, 2 FFMA2 with a
        scalar multiplicand (lane-FMAs per second)."""
        v = _dbl(0.0)
        self.call("unb_fp32_peak_form", int(form), ctypes.byref(v))
        return float(v.value)

    def fp64_peak(self):
        """Measured fp64 FMA rate of this device (lane-FMAs per second)."""
        v = _dbl(0.0)
        self.call("unb_fp64_peak", ctypes.byref(v))
        return float(v.value)

    # -- stateless scans ---------------------------------------------------------------
    def find_nearby(self, apts, bpts, radiussq, out=None):
        a = as_f64(apts, 2)
        b = as_f64(bpts, 2)
        if a.shape[1] != b.shape[1]:
            raise ValueError("dimensionality mismatch: %s vs %s" % (a.shape, b.shape))
        res = out if _direct_out(out, len(b), np.int64) else np.empty(len(b), dtype=np.int64)
        self.call("unb_find_nearby", _ptr(a), len(a), _ptr(b), len(b), b.shape[1],
                  float(radiussq), _ptr(res))
        if out is not None and res is not out:
            out[:len(b)] = res
            return out
        return res

    def has_neighbour(self, apts, bpts, radiussq):
        """``find_nearby(apts, bpts, radiussq) >= 0`` as a boolean mask."""
        a = as_f64(apts, 2)
        b = as_f64(bpts, 2)
        if a.shape[1] != b.shape[1]:
            raise ValueError("dimensionality mismatch: %s vs %s" % (a.shape, b.shape))
        mask = np.empty(len(b), dtype=bool)
        self.call("unb_has_neighbour", _ptr(a), len(a), _ptr(b), len(b), b.shape[1],
                  float(radiussq), _ptr(mask))
        return mask

    def count_nearby(self, apts, bpts, radiussq, out=None):
        a = as_f64(apts, 2)
        b = as_f64(bpts, 2)
        if a.shape[1] != b.shape[1]:
            raise ValueError("dimensionality mismatch: %s vs %s" % (a.shape, b.shape))
        res = out if _direct_out(out, len(b), np.int64) else np.empty(len(b), dtype=np.int64)
        self.call("unb_count_nearby", _ptr(a), len(a), _ptr(b), len(b), b.shape[1],
                  float(radiussq), _ptr(res))
        if out is not None and res is not out:
            out[:len(b)] = res
            return out
        return res

    def subtract_nearby(self, apts, radiussq, out=None):
        a = as_f64(apts, 2)
        res = out if _direct_out(out, a.shape, np.float64) else np.empty_like(a)
        self.call("unb_subtract_nearby", _ptr(a), a.shape[0], a.shape[1], float(radiussq), _ptr(res))
        if out is not None and res is not out:
            out[...] = res
            return out
        return res

    def compute_maxradiussq(self, apts, bpts):
        a = as_f64(apts, 2)
        b = as_f64(bpts, 2)
        if a.shape[1] != b.shape[1]:
            raise ValueError("dimensionality mismatch: %s vs %s" % (a.shape, b.shape))
        out = _dbl(0.0)
        self.call("unb_compute_maxradiussq", _ptr(a), len(a), _ptr(b), len(b), a.shape[1],
                  ctypes.byref(out))
        return float(out.value)

    def mean_pair_distance(self, pts, clusterids):
        p = as_f64(pts, 2)
        c = np.ascontiguousarray(clusterids, dtype=np.int64)
        if len(c) < len(p):
            raise ValueError("clusterids shorter than pts")
        out = _dbl(0.0)
        self.call("unb_mean_pair_distance", _ptr(p), _ptr(c), len(p), p.shape[1], ctypes.byref(out))
        return float(out.value)

    def inside_ellipsoid(self, points, center, invcov, square_radius):
        p = as_f64(points, 2)
        c = as_f64(center, 1)
        A = as_f64(invcov, 2)
        d = p.shape[1]
        if c.shape != (d,) or A.shape != (d, d):
            raise ValueError("ellipsoid shape mismatch: points %s center %s invcov %s"
                             % (p.shape, c.shape, A.shape))
        mask = np.empty(len(p), dtype=bool)
        self.call("unb_inside_ellipsoid", _ptr(p), len(p), d, _ptr(c), _ptr(A),
                  float(square_radius), _ptr(mask))
        return mask

    def transform(self, kind, inverse, pts, shift, mat):
        p = as_f64(pts)
        shape = p.shape
        p2 = p.reshape((-1, shape[-1]))
        d = p2.shape[1]
        s = as_f64(np.ravel(shift))
        m = as_f64(mat)
        if kind == LAYER_AFFINE:
            if s.shape != (d,) or m.shape != (d, d):
                raise ValueError("affine layer shape mismatch")
            name = "unb_untransform_affine" if inverse else "unb_transform_affine"
        else:
            m = as_f64(np.ravel(m))
            if s.shape != (d,) or m.shape != (d,):
                raise ValueError("scaling layer shape mismatch")
            name = "unb_untransform_scaling" if inverse else "unb_transform_scaling"
        out = np.empty_like(p2)
        self.call(name, _ptr(p2), len(p2), d, _ptr(s), _ptr(m), _ptr(out))
        return out.reshape(shape)

    # -- stateful region -----------------------------------------------------------------
    def region_sync_live(self, unormed):
        t = as_f64(unormed, 2)
        changed = _i64(0)
        self.call("unb_region_sync_live", _ptr(t), t.shape[0], t.shape[1], ctypes.byref(changed))
        return int(changed.value)

    def region_set_layer(self, kind, shift=None, mat=None, ndim=0):
        if kind == LAYER_IDENTITY:
            self.call("unb_region_set_layer", kind, None, None, int(ndim))
            return
        s = as_f64(np.ravel(shift))
        m = as_f64(mat) if kind == LAYER_AFFINE else as_f64(np.ravel(mat))
        self.call("unb_region_set_layer", kind, _ptr(s), _ptr(m), len(s))

    def region_set_ellipsoid(self, center, invcov, enlarge):
        c = as_f64(center, 1)
        A = as_f64(invcov, 2)
        if A.shape != (len(c), len(c)):
            raise ValueError("ellipsoid shape mismatch")
        self.call("unb_region_set_ellipsoid", _ptr(c), _ptr(A), float(enlarge), len(c))

    def region_set_radius(self, maxradiussq):
        self.call("unb_region_set_radius", float(maxradiussq))

    def region_set_transform_tolerance(self, tau):
        self.call("unb_region_set_transform_tolerance", float(tau))

    def uncertain(self):
        """Exact membership decisions of the last host-buffer region call that lie within the
        transform tolerance of the radius (0: every decision is the reference's)."""
        return self.stat(STAT_UNCERTAIN)

    def region_inside(self, pts, want_index=False, use_ellipsoid=True):
        p = as_f64(pts, 2)
        mask = np.empty(len(p), dtype=bool)
        idx = np.empty(len(p), dtype=np.int64) if want_index else None
        name = "unb_region_inside" if use_ellipsoid else "unb_region_friends"
        self.call(name, _ptr(p), len(p), _ptr(mask), _ptr(idx))
        return (mask, idx) if want_index else mask

    def region_inside_ellipsoid(self, pts):
        """Ellipsoid stage alone against the mirrored ellipsoid (chunked host pipeline)."""
        p = as_f64(pts, 2)
        mask = np.empty(len(p), dtype=bool)
        self.call("unb_region_inside_ellipsoid", _ptr(p), len(p), _ptr(mask))
        return mask

    def region_find_nearby(self, tpts):
        p = as_f64(tpts, 2)
        out = np.empty(len(p), dtype=np.int64)
        self.call("unb_region_find_nearby", _ptr(p), len(p), _ptr(out))
        return out

    def region_has_neighbour(self, tpts):
        """``find_nearby(unormed, tpts, r2) >= 0`` as a mask (no index needed)."""
        p = as_f64(tpts, 2)
        mask = np.empty(len(p), dtype=bool)
        self.call("unb_region_has_neighbour", _ptr(p), len(p), _ptr(mask))
        return mask

    def region_count_nearby(self, tpts):
        p = as_f64(tpts, 2)
        out = np.empty(len(p), dtype=np.int64)
        self.call("unb_region_count_nearby", _ptr(p), len(p), _ptr(out))
        return out

    def region_bootstrap(self, unormed, selected, u=None, ctrs=None, invcovs=None,
                         round_lo=0, round_hi=None):
        """Per-round ``(maxd, f)`` of the bootstrap (either half may be skipped:
        ``unormed=None`` skips the radius scan, ``u=None`` the enlargement)."""
        want_d = unormed is not None
        want_f = u is not None and ctrs is not None and invcovs is not None
        if not (want_d or want_f):
            raise ValueError("nothing to compute")
        t = as_f64(unormed, 2) if want_d else None
        if want_f:
            u = as_f64(u, 2)
        n, d = (t if want_d else u).shape
        sel = np.ascontiguousarray(selected, dtype=np.uint8)
        if sel.ndim != 2 or sel.shape[1] != n:
            raise ValueError("selected must be (nrounds, n)")
        nrounds = sel.shape[0]
        if round_hi is None:
            round_hi = nrounds
        maxd = np.zeros(nrounds) if want_d else None
        f = np.zeros(nrounds) if want_f else None
        if want_f:
            ctrs = as_f64(ctrs, 2)
            invcovs = as_f64(invcovs, 3)
            if u.shape != (n, d) or ctrs.shape != (nrounds, d) or invcovs.shape != (nrounds, d, d):
                raise ValueError("bootstrap ellipsoid shapes mismatch")
        self.call("unb_region_bootstrap", _ptr(t), _ptr(u) if want_f else None, n, d,
                  _ptr(sel), nrounds, int(round_lo), int(round_hi),
                  _ptr(ctrs) if want_f else None, _ptr(invcovs) if want_f else None,
                  _ptr(maxd), _ptr(f))
        return maxd, f

    def region_bootstrap_moments(self, u, selected, c0, round_lo=0, round_hi=None):
        """Per-round ``(count, sum(y), sum(y y^T))`` of the selected rows about ``c0``
        (``unb_region_bootstrap_moments``; upper triangle of the second moments filled)."""
        u = as_f64(u, 2)
        n, d = u.shape
        sel = np.ascontiguousarray(selected, dtype=np.uint8)
        if sel.ndim != 2 or sel.shape[1] != n:
            raise ValueError("selected must be (nrounds, n)")
        nrounds = sel.shape[0]
        if round_hi is None:
            round_hi = nrounds
        c0 = as_f64(c0, 1)
        sums = np.zeros((nrounds, d))
        sxx = np.zeros((nrounds, d, d))
        counts = np.zeros(nrounds, dtype=np.int64)
        self.call("unb_region_bootstrap_moments", _ptr(u), n, d, _ptr(sel), nrounds, int(round_lo),
                  int(round_hi), _ptr(c0), _ptr(sums), _ptr(sxx), _ptr(counts))
        return counts, sums, sxx

    def region_bootstrap_fold_dev(self, unormed, selected, u, ctrs, invcovs, round_lo, round_hi,
                                  host_failed, tag, out_ptr, stream=None):
        """Rounds ``[round_lo, round_hi)`` folded on the device into the 5-double buffer at the
        DEVICE address ``out_ptr`` (``[r2, f, failed, tag, -tag]``); only enqueues work."""
        t = as_f64(unormed, 2)
        u = as_f64(u, 2)
        n, d = t.shape
        sel = np.ascontiguousarray(selected, dtype=np.uint8)
        nrounds = sel.shape[0]
        ctrs = as_f64(ctrs, 2)
        invcovs = as_f64(invcovs, 3)
        if (sel.ndim != 2 or sel.shape[1] != n or u.shape != (n, d) or ctrs.shape != (nrounds, d)
                or invcovs.shape != (nrounds, d, d)):
            raise ValueError("bootstrap shapes mismatch")
        self.call("unb_region_bootstrap_fold_dev", _ptr(t), _ptr(u), n, d, _ptr(sel), nrounds,
                  int(round_lo), int(round_hi), _ptr(ctrs), _ptr(invcovs), 1 if host_failed else 0,
                  float(tag), int(out_ptr), stream)

    # -- likelihoods -----------------------------------------------------------------------
    def loglike_gauss(self, theta, centers, sigma, norm_const):
        p = as_f64(theta, 2)
        c = as_f64(np.broadcast_to(centers, (p.shape[1],)))
        out = np.empty(len(p))
        self.call("unb_loglike_gauss", _ptr(p), p.shape[1], len(p), _ptr(out), _ptr(c),
                  float(sigma), float(norm_const))
        return out

    def loglike_rosenbrock(self, theta):
        p = as_f64(theta, 2)
        out = np.empty(len(p))
        self.call("unb_loglike_rosenbrock", _ptr(p), p.shape[1], len(p), _ptr(out))
        return out

    def loglike_eggbox(self, z):
        p = as_f64(z, 2)
        out = np.empty(len(p))
        self.call("unb_loglike_eggbox", _ptr(p), p.shape[1], len(p), _ptr(out))
        return out

    def region_inside_loglike(self, pts, kind, lparams=None, mask_out=None, like_out=None):
        p = as_f64(pts, 2)
        mask = mask_out if mask_out is not None else np.empty(len(p), dtype=bool)
        like = like_out if like_out is not None else np.empty(len(p))
        lp = as_f64(lparams) if lparams is not None else None
        self.call("unb_region_inside_loglike", _ptr(p), len(p), _ptr(mask), _ptr(like),
                  int(kind), _ptr(lp))
        return mask, like

    def region_sample(self, nsamples, ndim, method, seed, offset, axes_T=None, like_kind=LOGLIKE_NONE,
                      lparams=None, Lmin=None):
        """Device-generated proposals filtered by the mirrored region (``unb_region_sample``).
        Returns ``(rows[k, ndim], logl[k] or None)``, accepted rows in draw order."""
        desc, keep = make_sample_desc(method, seed, offset, axes_T, like_kind, lparams, Lmin)
        rows = np.empty((int(nsamples), int(ndim)))
        like = np.empty(int(nsamples)) if like_kind != LOGLIKE_NONE else None
        n_out = _i64(0)
        self.call("unb_region_sample", ctypes.addressof(desc), int(nsamples), _ptr(rows), _ptr(like),
                  ctypes.byref(n_out), None)
        del keep
        k = int(n_out.value)
        return rows[:k], (like[:k] if like is not None else None)

    def sample_draw(self, method, nsamples, ndim, seed, offset, center=None, axes_T=None, enlarge=1.0):
        """The generator's raw draws and their unit-cube mask (``unb_sample_draw``; for tests)."""
        rows = np.empty((int(nsamples), int(ndim)))
        cube = np.empty(int(nsamples), dtype=bool)
        c = as_f64(center, 1) if center is not None else None
        a = as_f64(axes_T, 2) if axes_T is not None else None
        self.call("unb_sample_draw", int(method), int(nsamples), int(ndim),
                  int(seed) & 0xffffffffffffffff, int(offset) & 0xffffffffffffffff, _ptr(c), _ptr(a),
                  float(enlarge), _ptr(rows), _ptr(cube))
        return rows, cube

    def region_refill(self, u, region_mode, check_cube, xform, tregion, like_kind, lparams, Lmin):
        """Fused ``_refill_samples`` stage chain (``unb_region_refill``).  ``xform``: ``None`` or
        ``(scale, lo)``; ``tregion``: ``None`` or ``(center, invcov, enlarge)``.  Returns
        ``(flags uint8[n], logl[n], (n_member, n_tregion, n_accepted))``."""
        p = as_f64(u, 2)
        n, d = p.shape
        keep = []
        desc = RefillDesc()
        desc.region_mode = int(region_mode)
        desc.check_cube = 1 if check_cube else 0
        desc.loglike_kind = int(like_kind)
        if xform is None:
            desc.xform_kind = XFORM_IDENTITY
        else:
            sc = as_f64(np.broadcast_to(np.asarray(xform[0], dtype=float), (d,)))
            lo = as_f64(np.broadcast_to(np.asarray(xform[1], dtype=float), (d,)))
            keep += [sc, lo]
            desc.xform_kind = XFORM_SCALE_SHIFT
            desc.xform_scale = _ptr(sc)
            desc.xform_lo = _ptr(lo)
        if tregion is not None:
            ctr = as_f64(tregion[0])
            inv = as_f64(tregion[1], 2)
            if ctr.shape != (d,) or inv.shape != (d, d):
                raise ValueError("tregion ellipsoid does not match ndim=%d" % d)
            keep += [ctr, inv]
            desc.treg_center = _ptr(ctr)
            desc.treg_invcov = _ptr(inv)
            desc.treg_enlarge = float(tregion[2])
        if lparams is not None:
            lp = as_f64(lparams)
            keep.append(lp)
            desc.lparams = _ptr(lp)
        desc.Lmin = float(Lmin)
        flags = np.empty(n, dtype=np.uint8)
        like = np.empty(n)
        counts = np.zeros(3, dtype=np.int64)
        self.call("unb_region_refill", _ptr(p), n, d, ctypes.addressof(desc), _ptr(flags),
                  _ptr(like), _ptr(counts))
        del keep
        return flags, like, (int(counts[0]), int(counts[1]), int(counts[2]))


class SampleDesc(ctypes.Structure):
    """``unb_sample_desc`` of include/ultranest_b200.h."""
    _fields_ = [("method", ctypes.c_int32), ("loglike_kind", ctypes.c_int32),
                ("use_lmin", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("seed", ctypes.c_uint64), ("offset", ctypes.c_uint64),
                ("axes_T", ctypes.c_void_p), ("lparams", ctypes.c_void_p),
                ("Lmin", ctypes.c_double)]


def make_sample_desc(method, seed, offset, axes_T=None, like_kind=LOGLIKE_NONE, lparams=None, Lmin=None):
    """``(desc, keepalive)`` for ``unb_region_sample[_dev]``."""
    keep = []
    desc = SampleDesc()
    desc.method = int(method)
    desc.loglike_kind = int(like_kind)
    desc.use_lmin = 0 if Lmin is None else 1
    desc.Lmin = 0.0 if Lmin is None else float(Lmin)
    desc.seed = int(seed) & 0xffffffffffffffff
    desc.offset = int(offset) & 0xffffffffffffffff
    if axes_T is not None:
        a = as_f64(axes_T, 2)
        keep.append(a)
        desc.axes_T = _ptr(a)
    if lparams is not None:
        lp = as_f64(lparams)
        keep.append(lp)
        desc.lparams = _ptr(lp)
    return desc, keep


class StepDesc(ctypes.Structure):
    """``unb_step_desc`` of include/ultranest_b200.h."""
    _fields_ = [("xform_kind", ctypes.c_int32), ("loglike_kind", ctypes.c_int32),
                ("xform_scale", ctypes.c_void_p), ("xform_lo", ctypes.c_void_p),
                ("lparams", ctypes.c_void_p)]


def make_step_desc(ndim, xform, like_kind, lparams):
    """``(desc, keepalive)`` for the fused step kernels; ``xform``: ``None`` or ``(scale, lo)``."""
    keep = []
    desc = StepDesc()
    desc.loglike_kind = int(like_kind)
    if xform is None:
        desc.xform_kind = XFORM_IDENTITY
    else:
        sc = as_f64(np.broadcast_to(np.asarray(xform[0], dtype=float), (ndim,)))
        lo = as_f64(np.broadcast_to(np.asarray(xform[1], dtype=float), (ndim,)))
        keep += [sc, lo]
        desc.xform_kind = XFORM_SCALE_SHIFT
        desc.xform_scale = _ptr(sc)
        desc.xform_lo = _ptr(lo)
    if lparams is not None:
        lp = as_f64(lparams)
        keep.append(lp)
        desc.lparams = _ptr(lp)
    return desc, keep


class RefillDesc(ctypes.Structure):
    """``unb_refill_desc`` of include/ultranest_b200.h."""
    _fields_ = [("region_mode", ctypes.c_int32), ("check_cube", ctypes.c_int32),
                ("xform_kind", ctypes.c_int32), ("loglike_kind", ctypes.c_int32),
                ("xform_scale", ctypes.c_void_p), ("xform_lo", ctypes.c_void_p),
                ("treg_center", ctypes.c_void_p), ("treg_invcov", ctypes.c_void_p),
                ("treg_enlarge", ctypes.c_double), ("lparams", ctypes.c_void_p),
                ("Lmin", ctypes.c_double)]


def _direct_out(out, shape, dtype):
    """Can the library write straight into the caller's out-array?"""
    if out is None or not isinstance(out, np.ndarray):
        return False
    if out.dtype != dtype or not out.flags.c_contiguous or not out.flags.writeable:
        return False
    want = shape if isinstance(shape, tuple) else (shape,)
    return out.shape == want


_engine = None
_engine_lock = threading.Lock()


def get_engine():
    """Process-wide engine (device = $LOCAL_RANK, else $UNB_DEVICE, else 0)."""
    global _engine
    with _engine_lock:
        if _engine is None:
            _engine = Engine()
        return _engine


def reset_engine():
    global _engine
    with _engine_lock:
        if _engine is not None:
            _engine.close()
        _engine = None
