"""B200-native drop-in for ``ultranest.mlfriends`` (reference: ``ultranest/mlfriends.pyx``).

Same public names, constructor signatures, attributes, RNG call order and exception types as
the reference module, so ``ultranest/integrator.py`` runs unchanged on top of it
(``run(region_class=MLFriends)``, ``sampler.transform_layer_class = ...`` or
:func:`ultranest_b200.install`).  What differs is where the work happens:

* every O(N^2 d) / O(M N d) / O(M d^2) loop -- ``find_nearby``, ``count_nearby``,
  ``_subtract_nearby``, ``compute_maxradiussq``, the bootstrap rounds, ``_inside_ellipsoid``,
  the layer transforms, ``MLFriends.inside`` -- runs in hand-written sm_100a kernels behind the
  C ABI of ``include/ultranest_b200.h`` (:mod:`ultranest_b200._native`);
* the host keeps what the reference keeps in NumPy/LAPACK: d x d algebra (``cov``, ``eigh``,
  ``inv``, ``slogdet``), the legacy ``np.random`` stream (drawn in the reference's exact order so
  seeded runs stay identical), and the object/attribute protocol the integrator mutates in place
  (``region.u[i] = ...``, ``region.unormed[i] = ...``, ``region.ellipsoid_center = ...``,
  ``region.maxradiussq = None``; integrator.py:2749-2758, 2827).

  These host pieces are NOT new work: ``make_eigvals_positive``, ``bounding_ellipsoid``,
  ``vol_prefactor``, the layers' ``optimize`` / ``create_new``, ``create_ellipsoid`` / ``_set_axes``,
  ``SimpleRegion.compute_enlargement``, ``WrappingEllipsoid`` and the friends-of-friends growth loop
  are the reference's own NumPy expressions (mlfriends.pyx:275-322, 389-476, 547-569, 666-710,
  754-816, 1213-1237, 1460-1649), kept expression for expression because a seeded run is only the
  reference's run if these values are bit-identical -- roughly 250 of this file's lines.  What is new
  lives in ``csrc/`` and in the orchestration around it here: the device mirror (``_bind``), the
  fused calls, the bootstrap orchestration and its enlargement screen, the transform tolerance,
  the device-side proposal generator.

There is no CPU implementation of the scans in this package; without the CUDA library the
import fails.
"""
import numpy as np
from numpy import pi

from . import _native

int_dtype = np.int64

__all__ = [
    "find_nearby", "count_nearby", "subtract_nearby", "_subtract_nearby", "compute_maxradiussq",
    "compute_mean_pair_distance", "update_clusters", "make_eigvals_positive",
    "bounding_ellipsoid", "vol_prefactor", "_inside_ellipsoid",
    "ScalingLayer", "AffineLayer", "MaxPrincipleGapAffineLayer", "LocalAffineLayer",
    "MLFriends", "RobustEllipsoidRegion", "SimpleRegion", "WrappingEllipsoid",
]


def _engine():
    return _native.get_engine()


# ======================================================================================
# pair scans (free functions of the reference module)
# ======================================================================================

def find_nearby(apts, bpts, radiussq, nnearby):
    """Index of the FIRST point of ``apts`` within ``radiussq`` of each ``bpts`` row, else -1.

    Written into ``nnearby`` in place (reference: mlfriends.pyx:143-183).  Bit-exact.
    """
    _engine().find_nearby(apts, bpts, radiussq, out=nnearby)


def count_nearby(apts, bpts, radiussq, nnearby):
    """Number of ``apts`` rows within ``radiussq`` of each ``bpts`` row (mlfriends.pyx:31-68)."""
    _engine().count_nearby(apts, bpts, radiussq, out=nnearby)


def _subtract_nearby(apts, bpts, radiussq):
    """``bpts[j] = apts[j] - mean(apts[i] : |a_i-a_j|^2 <= radiussq)`` (mlfriends.pyx:73-113)."""
    apts = _native.as_f64(apts, 2)
    if bpts.shape != apts.shape:
        raise AssertionError("shape mismatch")
    _engine().subtract_nearby(apts, radiussq, out=bpts)


def subtract_nearby(upoints, maxradiussq):
    """Points with the mean of their ``maxradiussq``-neighbourhood removed (mlfriends.pyx:118-138)."""
    return _engine().subtract_nearby(upoints, maxradiussq)


def compute_maxradiussq(apts, bpts):
    """``max_j min_i |a_i - b_j|^2`` rounded to float32 like the reference's C ``float`` return
    (mlfriends.pyx:188-224; a ``cdef`` there, public here)."""
    return _engine().compute_maxradiussq(apts, bpts)


def compute_mean_pair_distance(pts, clusterids):
    """Mean Euclidean distance over same-cluster pairs, ids != 0 (mlfriends.pyx:229-270)."""
    total = _engine().mean_pair_distance(pts, clusterids)
    assert np.isfinite(total) or np.isnan(total), total
    return total


def _inside_ellipsoid(points, ellipsoid_center, ellipsoid_invcov, square_radius):
    """``einsum('ij,jk,ik->i', d, invcov, d) <= square_radius`` with ``d = points - center``
    (mlfriends.pyx:882-912), evaluated in the einsum's own accumulation order.  Bit-exact.
    Runs through the engine's chunked, double-buffered host pipeline."""
    eng = _engine()
    pts = _native.as_f64(points, 2)
    eng.region_set_ellipsoid(ellipsoid_center, ellipsoid_invcov, square_radius)
    return eng.region_inside_ellipsoid(pts)


# ======================================================================================
# clustering
# ======================================================================================

def _friends_of_friends(tpoints, maxradiussq, old_ids):
    """Grow clusters in t-space exactly like the reference (mlfriends.pyx:284-322): a cluster
    keeps absorbing every unassigned point that has a member within the radius; when nothing is
    absorbed the next cluster is seeded at the first unassigned point, or at the first point
    that carried the new id before.  Each absorption test is one device any-neighbour scan."""
    eng = _engine()
    n = len(tpoints)
    labels = np.zeros(n, dtype=int_dtype)
    current = 1
    seed = 0
    prior = old_ids == current
    if prior.any():
        seed = np.where(prior)[0][0]
    labels[seed] = current
    while True:
        free = labels == 0
        if not free.any():
            break
        hit = eng.has_neighbour(tpoints[labels == current, :], tpoints[free, :], maxradiussq)
        if hit.any():
            absorbed = free
            absorbed[free] = hit
            labels[absorbed] = current
        else:
            current = current + 1
            seed = np.where(free)[0][0]
            prior = old_ids == current
            if prior.any():
                seed = np.where(prior)[0][0]
            labels[seed] = current
    return labels


def _update_clusters(upoints, tpoints, maxradiussq, clusterids):
    assert upoints.shape[0] == tpoints.shape[0], \
        ('different number of points', upoints.shape[0], tpoints.shape[0])
    assert upoints.shape[1] == tpoints.shape[1], \
        ('different dimensionality of points', upoints.shape[1], tpoints.shape[1])
    # old ids may come from a longer array
    labels = _friends_of_friends(tpoints, maxradiussq, clusterids[:len(tpoints)])
    assert (labels > 0).all()
    present = np.unique(labels)
    nclusters = len(present)
    if nclusters == 1:
        return nclusters, labels, upoints
    centred = np.empty_like(upoints)
    for cid in present:
        members = labels == cid
        group = upoints[members, :]
        if len(group) > 1:
            origin = group.mean(axis=0).reshape((1, -1))
        else:
            # a lone point would be centred onto itself; use the population mean instead
            origin = upoints.mean(axis=0).reshape((1, -1))
        centred[members, :] = group - origin
    return nclusters, labels, centred


def update_clusters(upoints, tpoints, maxradiussq, clusterids=None):
    """Cluster ``upoints`` so that no two clusters have members closer than
    ``sqrt(maxradiussq)`` in t-space, re-using old ids where possible.

    Returns ``(nclusters, new_clusterids, overlapped_points)`` (mlfriends.pyx:348-384).
    """
    upoints = np.asarray(upoints, dtype=float)
    tpoints = np.asarray(tpoints, dtype=float)
    if clusterids is None:
        clusterids = np.zeros(len(tpoints), dtype=int_dtype)
    return _update_clusters(upoints, tpoints, maxradiussq, clusterids)


# ======================================================================================
# small host-side linear algebra (d x d; LAPACK like the reference)
# ======================================================================================

def make_eigvals_positive(a, targetprod):
    """Lift (near-)zero eigenvalues of the symmetric ``a`` so that the eigenvalue product
    reaches ``targetprod`` (mlfriends.pyx:389-421)."""
    assert np.isfinite(a).all(), a
    try:
        w, v = np.linalg.eigh(a)
    except np.linalg.LinAlgError as e:
        print(a, targetprod)
        raise e
    tiny = w < max(1.e-10, 1e-300**(1. / len(a)))
    if np.any(tiny):
        nzprod = np.prod(w[~tiny])
        nzeros = tiny.sum()
        w[tiny] = (targetprod / nzprod) ** (1. / nzeros)
        a = np.dot(np.dot(v, np.diag(w)), np.linalg.inv(v))
    return a


def bounding_ellipsoid(x, minvol=0.):
    """Centre and (d+2)-scaled sample covariance of ``x`` (mlfriends.pyx:426-476)."""
    ndim = x.shape[1]
    ctr = np.mean(x, axis=0)
    cov = np.cov(x - ctr, rowvar=0)
    assert np.isfinite(cov).all(), (cov, x)
    if ndim == 1:
        cov = np.atleast_2d(cov)
    # uniform points in an n-ball have covariance r^2/(n+2): undo that factor
    cov *= (ndim + 2)
    if minvol > 0:
        cov = make_eigvals_positive(cov, minvol)
    return ctr, cov


def vol_prefactor(n):
    """Volume constant of the ``n``-sphere (mlfriends.pyx:853-879)."""
    if n % 2 == 0:
        f, i = 1., 2
    else:
        f, i = 2., 3
    while i <= n:
        f *= 2. / i * pi
        i += 2
    return f


# ======================================================================================
# transformation layers
# ======================================================================================

class ScalingLayer(object):
    """Per-axis shift and scale (mlfriends.pyx:479-620)."""

    _kind = _native.LAYER_SCALING

    def __init__(self, mean=0, std=1, nclusters=1, wrapped_dims=[], clusterids=None):
        self.mean = mean
        self.std = std
        self.nclusters = nclusters
        self.wrapped_dims = wrapped_dims
        self.has_wraps = len(wrapped_dims) > 0
        self.clusterids = clusterids

    # -- circular parameters (host, elementwise; mlfriends.pyx:491-545) ----------------
    def optimize_wrap(self, points):
        """Place the wrap cut of every circular axis in the middle of its largest gap."""
        if not self.has_wraps:
            return
        cuts = []
        for axis in self.wrapped_dims:
            edges = np.sort(np.concatenate(([0.], points[:, axis], [1.])))
            assert edges[0] == 0 and edges[-1] == 1
            widest = np.argmax(edges[1:] - edges[:-1])
            cuts.append((edges[widest] + edges[widest + 1]) / 2.)
        self.wrap_cuts = cuts

    def _shift_circular(self, arr, offsets):
        """``fmod(x + offset, 1)`` on the circular columns of a 2-D copy of ``arr``."""
        out = arr.copy().reshape((-1, arr.shape[-1]))
        dims = list(self.wrapped_dims)
        out[:, dims] = np.fmod(out[:, dims] + np.asarray(offsets), 1)
        return out

    def wrap(self, points):
        if not self.has_wraps:
            return points
        return self._shift_circular(points, [1 - cut for cut in self.wrap_cuts])

    def unwrap(self, wpoints):
        if not self.has_wraps:
            return wpoints
        return self._shift_circular(wpoints, list(self.wrap_cuts))

    # -- learning ------------------------------------------------------------------------
    def optimize(self, points, centered_points, clusterids=None, minvol=0.):
        """Estimate mean/std (mlfriends.pyx:547-569); ``minvol`` is ignored."""
        self.optimize_wrap(points)
        wrapped_points = self.wrap(points)
        self.mean = wrapped_points.mean(axis=0).reshape((1, -1))
        self.std = centered_points.std(axis=0).reshape((1, -1))
        self.axes = np.diag(self.std[0])
        self.logvolscale = np.sum(np.log(self.std))
        self.set_clusterids(clusterids=clusterids, npoints=len(points))

    def set_clusterids(self, clusterids=None, npoints=None):
        if clusterids is None and self.clusterids is None and npoints is not None:
            clusterids = np.ones(npoints, dtype=int_dtype)
        if clusterids is not None:
            self.clusterids = clusterids

    def _cluster(self, upoints, maxradiussq):
        uwpoints = self.wrap(upoints)
        tpoints = self.transform(upoints)
        return (uwpoints,) + tuple(update_clusters(uwpoints, tpoints, maxradiussq, self.clusterids))

    def create_new(self, upoints, maxradiussq, minvol=0.):
        """Next layer learned from this layer's clustering (mlfriends.pyx:580-603)."""
        _, nclusters, clusteridxs, overlapped = self._cluster(upoints, maxradiussq)
        s = self.__class__(nclusters=nclusters, wrapped_dims=self.wrapped_dims, clusterids=clusteridxs)
        s.optimize(upoints, overlapped)
        return s

    # -- device parameters ---------------------------------------------------------------
    def _device_params(self, ndim):
        """(kind, shift[d], scale[d]) for the fused device pipelines."""
        mean = np.ascontiguousarray(np.broadcast_to(np.ravel(self.mean), (ndim,)), dtype=float)
        std = np.ascontiguousarray(np.broadcast_to(np.ravel(self.std), (ndim,)), dtype=float)
        return _native.LAYER_SCALING, mean, std

    def transform(self, u):
        """Cube space -> whitened space, ``(w - mean) / std`` (mlfriends.pyx:605-611).

        This method carries VALUES into the run (``region.unormed``, the t-space bounding box),
        so it is the reference's own NumPy expression; bulk proposals are transformed on the
        device inside the fused pipelines (``MLFriends.inside`` ...), where only decisions depend
        on it.  (``unb_transform_scaling`` gives the same bits.)"""
        w = self.wrap(u) if self.has_wraps else u
        return ((w - self.mean) / self.std).reshape(u.shape)

    def untransform(self, ww):
        """Whitened space -> cube space (mlfriends.pyx:613-620)."""
        w = (ww * self.std) + self.mean
        if self.has_wraps:
            return self.unwrap(w).reshape(ww.shape)
        return w.reshape(ww.shape)


class AffineLayer(ScalingLayer):
    """Affine whitening learned from the (cluster-centred) sample covariance
    (mlfriends.pyx:623-752).

    ``transform`` / ``untransform`` carry values into the run (``region.unormed``, bounding box,
    returned samples) and are the reference's own ``np.dot`` expressions, so a seeded run
    reproduces the reference's numbers exactly.  Bulk proposals are transformed on the device
    inside the fused pipelines with the library's defined order (k ascending, fused multiply-add;
    DESIGN.md 4.4), where only accept/reject decisions depend on it.
    """

    _kind = _native.LAYER_AFFINE

    def __init__(self, ctr=0, T=1, invT=1, nclusters=1, wrapped_dims=[], clusterids=None):
        self.ctr = ctr
        self.T = T
        self.invT = invT
        self.nclusters = nclusters
        self.wrapped_dims = wrapped_dims
        self.has_wraps = len(wrapped_dims) > 0
        self.clusterids = clusterids

    def optimize(self, points, centered_points, clusterids=None, minvol=0.):
        """Covariance, its eigen-decomposition and the whitening matrices
        (mlfriends.pyx:666-710; LAPACK on the host, d x d)."""
        self.optimize_wrap(points)
        wrapped_points = self.wrap(points)
        self.ctr = np.mean(wrapped_points, axis=0)
        cov = np.cov(centered_points, rowvar=0)
        cov *= (len(self.ctr) + 2)
        self.cov = cov
        eigval, eigvec = np.linalg.eigh(cov)
        eigvalmin = eigval.max() * 1e-40
        eigval[eigval < eigvalmin] = eigvalmin
        a = np.linalg.inv(cov)   # escalates a singular covariance
        self.logvolscale = np.linalg.slogdet(a)[1] * -0.5
        self.T = eigvec * eigval**-0.5
        self.invT = np.linalg.inv(self.T)
        self.axes = self.invT
        self.set_clusterids(clusterids=clusterids, npoints=len(points))

    def create_new(self, upoints, maxradiussq, minvol=0.):
        _, nclusters, clusteridxs, overlapped = self._cluster(upoints, maxradiussq)
        s = self.__class__(nclusters=nclusters, wrapped_dims=self.wrapped_dims, clusterids=clusteridxs)
        s.optimize(upoints, overlapped, minvol=minvol)
        return s

    def _is_learned(self):
        return np.ndim(self.T) == 2

    def _device_params(self, ndim):
        return (_native.LAYER_AFFINE, np.ascontiguousarray(self.ctr, dtype=float),
                np.ascontiguousarray(self.T, dtype=float))

    def transform(self, u):
        """Cube space -> whitened space, ``np.dot(w - ctr, T)`` (mlfriends.pyx:737-743)."""
        w = self.wrap(u) if self.has_wraps else u
        return np.dot(w - self.ctr, self.T)

    def untransform(self, ww):
        """Whitened space -> cube space (mlfriends.pyx:745-752)."""
        w = np.dot(ww, self.invT) + self.ctr
        if self.has_wraps:
            return self.unwrap(w).reshape(ww.shape)
        return w.reshape(ww.shape)


class MaxPrincipleGapAffineLayer(AffineLayer):
    """Affine layer whose next covariance is learned after splitting the cluster-centred points
    at the largest gap along the principal axis (mlfriends.pyx:754-816)."""

    def create_new(self, upoints, maxradiussq, minvol=0.):
        _, nclusters, clusteridxs, overlapped = self._cluster(upoints, maxradiussq)
        cov = np.cov(overlapped, rowvar=0)
        cov *= (len(self.ctr) + 2)
        eigval, eigvec = np.linalg.eigh(cov)
        principal_vector = eigvec[:, -1]
        t = np.dot(overlapped - overlapped.mean(axis=0).reshape((1, -1)), principal_vector)
        tsorted = np.sort(t)
        gap = np.argmax(np.diff(tsorted))
        tsep = (tsorted[gap] + tsorted[gap + 1]) / 2
        left = t < tsep
        halved = overlapped.copy()
        halved[left, :] -= overlapped[left, :].mean(axis=0)
        halved[~left, :] -= overlapped[~left, :].mean(axis=0)
        s = MaxPrincipleGapAffineLayer(nclusters=nclusters, wrapped_dims=self.wrapped_dims,
                                       clusterids=clusteridxs)
        s.optimize(upoints, halved, minvol=minvol)
        return s


class LocalAffineLayer(AffineLayer):
    """Affine layer whose next covariance is learned from points co-centred with their
    MLFriends neighbourhood (mlfriends.pyx:819-850); the neighbourhood means come from the
    device ``subtract_nearby``."""

    def create_new(self, upoints, maxradiussq, minvol=0.):
        uwpoints, nclusters, clusteridxs, _ = self._cluster(upoints, maxradiussq)
        s = self.__class__(nclusters=nclusters, wrapped_dims=self.wrapped_dims, clusterids=clusteridxs)
        local = subtract_nearby(uwpoints, maxradiussq)
        s.optimize(upoints, local, minvol=minvol)
        return s


# ======================================================================================
# regions
# ======================================================================================

def _draw_rounds(rng, npoints, nbootstraps):
    """Selection masks of ``nbootstraps`` rounds, drawn exactly like the reference's loop
    (``idx = rng.randint(N, size=N); selected[idx] = True``, mlfriends.pyx:1045-1047), plus the
    RNG state before the first draw so a failing round can leave the stream where the
    reference's early exit would."""
    get_state = getattr(rng, "get_state", None)
    state = get_state() if get_state is not None else None
    selected = np.zeros((nbootstraps, npoints), dtype=bool)
    for r in range(nbootstraps):
        selected[r, rng.randint(npoints, size=npoints)] = True
    return selected, state


def _rewind_rounds(rng, state, npoints, rounds_consumed):
    """Put ``rng`` where the reference would have left it after ``rounds_consumed`` rounds."""
    if state is None or not hasattr(rng, "set_state"):
        return
    rng.set_state(state)
    for _ in range(rounds_consumed):
        rng.randint(npoints, size=npoints)


#: relative margin of the enlargement screen (see :func:`_bootstrap_rounds_screened`)
SCREEN_MARGIN = 1e-6
#: how often the screen was used / declined / how many rounds needed the exact host algebra
screen_stats = {"screened": 0, "declined": 0, "exact_rounds": 0, "rounds": 0}


def _bootstrap_rounds_screened(u, unormed, selected, lo, hi, active):
    """The same per-round results as :func:`_bootstrap_rounds` with (almost) no host algebra.

    ``compute_enlargement`` only needs ``max_r f_r`` (mlfriends.pyx:1062-1066).  So the per-round
    bounding ellipsoids are first taken from moments accumulated ON THE DEVICE
    (``unb_region_bootstrap_moments``: mean and covariance up to summation order; the 30 small
    ``inv`` stay on the host as one batched LAPACK call), the device evaluates every round's ``f``
    with them, and only the rounds whose screened ``f`` lies within ``SCREEN_MARGIN`` of the largest
    are recomputed with the reference's own NumPy expressions (``np.cov``, ``inv``) -- typically
    one round instead of thirty.  The maximum is then attained by an exactly computed round and
    every other round is below it by more than the screen can err (covariances are required to be
    well conditioned for that: ``cond * 1e-13 << SCREEN_MARGIN``), so the returned maximum is the
    reference's value bit for bit.  Returns ``None`` whenever anything looks unusual (few points,
    ill-conditioned or singular covariance, non-positive ``f``); the caller then runs the exact
    path, which also reproduces the reference's failure semantics."""
    eng = _engine()
    nrounds, N = selected.shape
    ndim = u.shape[1]
    act = [r for r in range(lo, hi) if active[r]]
    if len(act) < 3 or N < 8 * (ndim + 2):
        return None
    c0 = np.mean(u, axis=0)
    counts, sums, sxx = eng.region_bootstrap_moments(u, selected, c0, lo, hi)
    n = counts[act].astype(float)
    if (n < 4 * (ndim + 2)).any():
        return None
    ybar = sums[act] / n[:, None]
    S = sxx[act]
    S = np.triu(S) + np.transpose(np.triu(S, 1), (0, 2, 1))
    with np.errstate(all='ignore'):
        cov = (S - n[:, None, None] * ybar[:, :, None] * ybar[:, None, :]) / (n - 1)[:, None, None] * (ndim + 2)
        try:
            inv = np.linalg.inv(cov)
        except np.linalg.LinAlgError:
            return None
        if not (np.isfinite(cov).all() and np.isfinite(inv).all()):
            return None
        cond = np.sqrt((cov**2).sum(axis=(1, 2))) * np.sqrt((inv**2).sum(axis=(1, 2)))
    if not cond.max() * 1e-13 < SCREEN_MARGIN * 1e-2:
        return None
    ctrs = np.zeros((nrounds, ndim))
    invcovs = np.zeros((nrounds, ndim, ndim))
    ctrs[act] = c0 + ybar
    invcovs[act] = inv
    maxd_r, f_r = eng.region_bootstrap(unormed, selected, u=u, ctrs=ctrs, invcovs=invcovs,
                                       round_lo=lo, round_hi=hi)
    fa = f_r[act]
    if not (np.isfinite(fa).all() and (fa > 0).all()):
        return None
    cand = [r for r in act if f_r[r] >= fa.max() * (1.0 - SCREEN_MARGIN)]
    for r in cand:      # the rounds that can decide the maximum: the reference's own algebra
        try:
            ctr, cov_r = bounding_ellipsoid(u[selected[r], :], minvol=0.)
            invcovs[r] = np.linalg.inv(cov_r)
            ctrs[r] = ctr
        except (np.linalg.LinAlgError, FloatingPointError, AssertionError, Warning):
            return None
        _, f_exact = eng.region_bootstrap(None, selected, u=u, ctrs=ctrs, invcovs=invcovs,
                                          round_lo=r, round_hi=r + 1)
        if not (np.isfinite(f_exact[r]) and f_exact[r] > 0):
            return None
        f_r[r] = f_exact[r]
    top = max(f_r[r] for r in cand)
    rest = [f_r[r] for r in act if r not in cand]
    if rest and not max(rest) < top * (1.0 - SCREEN_MARGIN / 2):
        return None
    screen_stats["screened"] += 1
    screen_stats["exact_rounds"] += len(cand)
    screen_stats["rounds"] += len(act)
    return maxd_r, f_r, active, None


def _bootstrap_rounds(u, unormed, selected, lo, hi, minvol, screen=True):
    """Rounds ``lo <= r < hi`` of the MLFriends bootstrap: per-round radius^2 (float32-rounded
    like the reference) and enlargement.  Host: d x d ``bounding_ellipsoid`` / ``inv`` of each
    round (mlfriends.pyx:1057-1058); device: everything O(N^2 d) / O(N d^2).

    Returns ``(maxd_r, f_r, active, failure)``; ``failure`` is ``None`` or ``(round, exception)``
    for the first round whose host algebra failed (rounds after it are not evaluated, like the
    reference's loop).  Rounds with all/none selected are inactive (mlfriends.pyx:1048-1049).
    """
    nrounds, N = selected.shape
    ndim = u.shape[1]
    active = ~(selected.all(axis=1) | ~selected.any(axis=1))
    if screen and minvol == 0:
        out = _bootstrap_rounds_screened(u, unormed, selected, lo, hi, active)
        if out is not None:
            return out
        screen_stats["declined"] += 1
    ctrs = np.zeros((nrounds, ndim))
    invcovs = np.zeros((nrounds, ndim, ndim))
    failure = None
    stop = hi
    for r in range(lo, hi):
        if not active[r]:
            continue
        try:
            ctr, cov = bounding_ellipsoid(u[selected[r], :], minvol=minvol)
            invcovs[r] = np.linalg.inv(cov)
            ctrs[r] = ctr
        except (np.linalg.LinAlgError, FloatingPointError, AssertionError, Warning) as exc:
            failure = (r, exc)
            stop = r
            break
    if stop > lo:
        maxd_r, f_r = _engine().region_bootstrap(unormed, selected, u=u, ctrs=ctrs,
                                                 invcovs=invcovs, round_lo=lo, round_hi=stop)
    else:
        maxd_r, f_r = np.zeros(nrounds), np.zeros(nrounds)
    return maxd_r, f_r, active, failure


class MLFriends(object):
    """MLFriends region (mlfriends.pyx:915-1257): union of equal-radius balls around the live
    points in the whitened space, intersected with a wrapping ellipsoid."""

    # names of the proposal generators sample() rotates through, in the reference's order
    _method_names = ("sample_from_transformed_boundingbox", "sample_from_boundingbox",
                     "sample_from_points", "sample_from_wrapping_ellipsoid")

    def __init__(self, u, transformLayer):
        inside_cube = np.logical_and(u > 0, u < 1)
        if not inside_cube.all():
            raise ValueError("not all u values are between 0 and 1: %s" % u[~inside_cube.all()])
        self.u = u
        self.set_transformLayer(transformLayer)
        self.sampling_methods = [getattr(self, name) for name in self._method_names]
        self.current_sampling_method = self.sample_from_boundingbox
        self.vol_prefactor = vol_prefactor(self.u.shape[1])

    def _draw_in_wrapping_ellipsoid(self, nsamples, want_mask=True):
        """Uniform draws inside the wrapping ellipsoid (mlfriends.pyx:1145-1154): one
        ``normal(size=(n, d))`` then one ``uniform(size=(n, 1))`` from the global stream.
        Returns the draws and the mask of those inside the unit cube."""
        ndim = self.u.shape[1]
        z = np.random.normal(size=(nsamples, ndim))
        sqnorm = (z**2).sum(axis=1)
        assert (sqnorm > 0).all(), sqnorm
        z /= (sqnorm**0.5).reshape((nsamples, 1))
        assert self.enlarge > 0, self.enlarge
        radial = np.random.uniform(size=(nsamples, 1))**(1. / ndim)
        w = self.ellipsoid_center + np.dot(z * self.enlarge**0.5 * radial, self.ellipsoid_axes_T)
        return w, (np.logical_and(w > 0, w < 1).all(axis=1) if want_mask else None)

    # -- geometry ------------------------------------------------------------------------
    def estimate_volume(self):
        """Order of magnitude of the log-volume around one live point (mlfriends.pyx:953-970)."""
        r = self.maxradiussq**0.5
        N, ndim = self.u.shape
        return self.transformLayer.logvolscale + np.log(r) * ndim

    def set_transformLayer(self, transformLayer):
        """New whitening layer; recomputes ``unormed`` and invalidates ``maxradiussq``."""
        self.transformLayer = transformLayer
        self.unormed = self.transformLayer.transform(self.u)
        assert np.isfinite(self.unormed).all(), (self.unormed, self.u)
        self.bbox_lo = self.unormed.min(axis=0)
        self.bbox_hi = self.unormed.max(axis=0)
        self.maxradiussq = None

    # -- bootstrapping -------------------------------------------------------------------
    def compute_maxradiussq(self, nbootstraps=50):
        """Bootstrapped MLFriends radius with the GLOBAL ``np.random`` stream
        (mlfriends.pyx:988-1015); all rounds run in one device launch."""
        N, ndim = self.u.shape
        selected, _ = _draw_rounds(np.random, N, nbootstraps)
        per_round, _ = _engine().region_bootstrap(self.unormed, selected)
        maxd = 0
        for r in range(nbootstraps):
            maxd = max(maxd, per_round[r])
        assert maxd > 0, (maxd, self.u)
        return maxd

    def compute_enlargement(self, nbootstraps=50, minvol=0., rng=np.random):
        """``(max radius^2, max ellipsoid enlargement)`` over ``nbootstraps`` rounds
        (mlfriends.pyx:1017-1070).

        Host: the selection masks (``rng`` order preserved) and each round's d x d
        ``bounding_ellipsoid`` / ``inv``.  Device: all rounds' nearest-neighbour max-min scans
        and einsum maxima in one launch each.  With :mod:`ultranest_b200.distributed` enabled
        the rounds are sharded over the ranks and re-united by ONE allreduce(MAX)
        (the reference's gather+bcast, integrator.py:395-404).
        """
        from . import distributed
        N, ndim = self.u.shape
        assert np.isfinite(self.unormed).all(), self.unormed
        selected, state = _draw_rounds(rng, N, nbootstraps)
        if distributed.world_size() > 1:
            return distributed.reduce_enlargement(self.u, self.unormed, selected, minvol)
        maxd_r, f_r, active, failure = _bootstrap_rounds(self.u, self.unormed, selected, 0,
                                                         nbootstraps, minvol)
        # the reference interleaves both checks round by round (mlfriends.pyx:1044-1066): a bad f
        # in a round BEFORE the one whose host algebra failed is what it reports
        last = nbootstraps if failure is None else failure[0]
        maxd = 0.0
        maxf = 0.0
        for r in range(last):
            if not active[r]:
                continue
            maxd = max(maxd, maxd_r[r])
            f = f_r[r]
            if not np.isfinite(f) or not f > 0:
                _rewind_rounds(rng, state, N, r + 1)
                assert np.isfinite(f), (self.unormed, f)
                raise np.linalg.LinAlgError("Distances are not positive")
            maxf = max(maxf, f)
        if failure is not None:
            _rewind_rounds(rng, state, N, failure[0] + 1)
            raise failure[1]
        assert maxd > 0, (maxd, self.u, self.unormed)
        assert maxf > 0, (maxf, self.u, self.unormed)
        return maxd, maxf

    # -- device mirror ---------------------------------------------------------------------
    def _bind(self, need_ellipsoid=True):
        """Bring the engine's region mirror up to date with this object's (mutable) attributes.
        ``unormed`` is diffed row-wise inside the library, so the integrator's in-place
        single-row patches cost one small copy."""
        eng = _engine()
        eng.region_sync_live(self.unormed)
        eng.region_set_radius(self.maxradiussq)
        if need_ellipsoid:
            layer = self.transformLayer
            kind, shift, mat = layer._device_params(self.u.shape[1])
            eng.region_set_layer(kind, shift, mat, self.u.shape[1])
            eng.region_set_ellipsoid(self.ellipsoid_center, self.ellipsoid_invcov, self.enlarge)
            eng.region_set_transform_tolerance(self._transform_tolerance())
        return eng

    def _transform_tolerance(self):
        """Bound ``tau`` on what the layer transform's summation order can do to a pair distance
        near the radius.  The fused device calls whiten proposals in a defined order, the
        reference with ``np.dot`` (OpenBLAS); either way a t-coordinate is a d-term sum of
        products, so the two differ by at most ``delta = 2 (d+2) u |x| |T|_F`` per row
        (``x = w - ctr``; rows that reach the neighbour scan lie inside the wrapping ellipsoid, so
        ``|x| <= sqrt(enlarge) * max axis + sqrt(d)``), and a squared distance near ``r^2`` by at
        most ``2 r delta + delta^2`` plus the rounding of its own sum.  The library reports exact
        decisions within ``tau`` of the radius; :meth:`inside` then re-decides the call with the
        reference's own transform.  Zero for layers whose device transform is
        bit-exact (scaling / identity)."""
        layer = self.transformLayer
        if not isinstance(layer, AffineLayer) or not layer._is_learned():
            return 0.0
        invcov = self.ellipsoid_invcov      # the matrix the device filter uses (callers may set it)
        key = (id(layer.T), id(invcov), float(self.enlarge), float(self.maxradiussq))
        cached = getattr(self, '_tau_cache', None)
        if cached is not None and cached[0] == key:
            return cached[1]
        ndim = self.u.shape[1]
        eps = 2.0**-53
        tfro = float(np.sqrt((np.asarray(layer.T, dtype=float)**2).sum()))
        lam_min = float(np.linalg.eigvalsh(np.asarray(invcov, dtype=float))[0])
        if not lam_min > 0:
            return 0.0
        xmax = (float(self.enlarge) / lam_min)**0.5 + ndim**0.5   # longest semi-axis + |c_ell - ctr|
        delta = 2.0 * (ndim + 2) * eps * xmax * tfro
        r2 = float(self.maxradiussq)
        tau = 2.0 * (2.0 * r2**0.5 * delta + delta * delta + 2.0 * (ndim + 2) * eps * r2)
        if not np.isfinite(tau):
            tau = 0.0
        self._tau_cache = (key, tau)
        return tau

    def _members_host_transform(self, pts, use_ellipsoid=True, check_cube=False):
        """Membership decided like the reference does it (mlfriends.pyx:1186-1211): ellipsoid on the
        device in the einsum order, layer transform with the reference's ``np.dot`` on the HOST,
        exact neighbour scan on the device.  The fused calls fall back to this when the library
        reports a decision inside the transform tolerance (one call in many thousands)."""
        pts = np.asarray(pts, dtype=float)
        mask = self.inside_ellipsoid(pts) if use_ellipsoid else np.ones(len(pts), dtype=bool)
        if check_cube:
            mask &= np.logical_and(pts > 0, pts < 1).all(axis=1)
        if mask.any():
            bpts = self.transformLayer.transform(pts[mask, :])
            mask[mask] = self._bind(need_ellipsoid=False).region_has_neighbour(bpts)
        return mask

    def _fused_ok(self):
        """May proposals be transformed on the device?  Not with circular dimensions (host wrap),
        not before the layer is learned, and not when the radius is so small that the ulp-level
        difference between the device transform and ``np.dot`` could matter (a live point must
        stay inside its own ball even for ``maxradiussq = 1e-90``, test_regionsampling.py:46-48)."""
        layer = self.transformLayer
        if layer.has_wraps:
            return False
        if isinstance(layer, AffineLayer) and not layer._is_learned():
            return False
        scale2 = max(float(np.max(np.abs(self.bbox_lo))), float(np.max(np.abs(self.bbox_hi))), 1e-300)**2
        return self.maxradiussq > 1e-18 * scale2 * self.u.shape[1]

    # -- sampling (host RNG in the reference's order, device filters) ----------------------
    def sample_from_points(self, nsamples=100):
        """Draw inside balls around random live points, thin by the neighbour count
        (mlfriends.pyx:1072-1094)."""
        N, ndim = self.u.shape
        idx = np.random.randint(N, size=nsamples)
        v = np.random.normal(size=(nsamples, ndim))
        v *= (np.random.uniform(size=nsamples)**(1. / ndim) / np.linalg.norm(v, axis=1)).reshape((-1, 1))
        v = self.unormed[idx, :] + v * self.maxradiussq**0.5
        nnearby = self._bind(need_ellipsoid=False).region_count_nearby(v)
        vmask = np.random.uniform(high=nnearby) < 1
        w = self.transformLayer.untransform(v[vmask, :])
        wmask = np.logical_and(w > 0, w < 1).all(axis=1)
        wmask[wmask] = self.inside_ellipsoid(w[wmask])
        return w[wmask, :]

    def sample_from_boundingbox(self, nsamples=100):
        """Uniform draws in the unit cube filtered by the region (mlfriends.pyx:1096-1112)."""
        N, ndim = self.u.shape
        u = np.random.uniform(size=(nsamples, ndim))
        return u[self.inside(u), :]

    def sample_from_transformed_boundingbox(self, nsamples=100):
        """Uniform draws in the whitened bounding box (mlfriends.pyx:1114-1133)."""
        N, ndim = self.u.shape
        r = self.maxradiussq**0.5
        v = np.random.uniform(self.bbox_lo - r, self.bbox_hi + r, size=(nsamples, ndim))
        vmask = self._bind(need_ellipsoid=False).region_has_neighbour(v)
        w = self.transformLayer.untransform(v[vmask, :])
        wmask = np.logical_and(w > 0, w < 1).all(axis=1)
        wmask[wmask] = self.inside_ellipsoid(w[wmask])
        return w[wmask, :]

    def sample_from_wrapping_ellipsoid(self, nsamples=100):
        """Uniform draws in the wrapping ellipsoid filtered by the friends test
        (mlfriends.pyx:1135-1160)."""
        w, wmask = self._draw_in_wrapping_ellipsoid(nsamples)
        if self._fused_ok():     # transform + neighbour scan in one device pipeline
            eng = self._bind()
            vmask = eng.region_inside(w[wmask, :], use_ellipsoid=False)
            if eng.uncertain():
                vmask = self._members_host_transform(w[wmask, :], use_ellipsoid=False)
        else:
            v = self.transformLayer.transform(w[wmask, :])
            vmask = self._bind(need_ellipsoid=False).region_has_neighbour(v)
        return w[wmask, :][vmask, :]

    def _propose(self, nsamples):
        """First half of :meth:`sample` for the fused refill (:mod:`ultranest_b200.refill`):
        consume the RNG exactly like the current sampling method, but hand the membership test
        to the caller's device pipeline.  Returns ``(rows, region_mode, check_cube)`` with
        ``region_mode`` 2 = rows still need ``inside()`` (bounding-box draws,
        mlfriends.pyx:1096-1112), 1 = rows lie in the wrapping ellipsoid and need the cube and
        friends tests (mlfriends.pyx:1154-1160), 0 = rows are finished region samples."""
        method = getattr(self.current_sampling_method, '__func__', None)
        if self._fused_ok():
            if method is MLFriends.sample_from_boundingbox:
                N, ndim = self.u.shape
                return np.random.uniform(size=(nsamples, ndim)), 2, False
            if method is MLFriends.sample_from_wrapping_ellipsoid:
                w, _ = self._draw_in_wrapping_ellipsoid(nsamples, want_mask=False)
                return w, 1, True
        return self.sample(nsamples=nsamples), 0, False

    def _after_sample(self, nfound):
        """Second half of :meth:`sample`: method switch on an empty draw (mlfriends.pyx:1180-1183)."""
        if nfound == 0:
            self.current_sampling_method = \
                self.sampling_methods[np.random.randint(len(self.sampling_methods))]

    # -- device-side proposal generation ("throughput mode", NOT the reference's random stream) ----
    #: set ``region.device_rng = True`` (and optionally ``region.device_seed``) to let
    #: :meth:`sample` draw its proposals on the device; the default keeps every draw on the host
    #: ``np.random`` stream in the reference's order (seeded runs identical to the reference's)
    device_rng = False
    device_seed = 0
    _device_draws = 0

    def sample_device(self, nsamples=100, method=None, seed=None, loglike=None, Lmin=None):
        """``nsamples`` proposals drawn ON THE DEVICE (Philox4x32-10 keyed by ``seed``) and
        filtered by the region like :meth:`sample_from_wrapping_ellipsoid` (default) or
        :meth:`sample_from_boundingbox` (``method="boundingbox"``): only accepted rows cross PCIe.
        With a device likelihood from :mod:`ultranest_b200.likelihoods` returns ``(rows, logl)``
        (and with ``Lmin`` only rows with ``logl > Lmin``), else ``rows``.  Statistically the
        reference's proposals (mlfriends.pyx:1096-1112, 1135-1160), not its random stream."""
        from . import _native, distributed
        if type(self) is not MLFriends:
            raise NotImplementedError("device-side proposals are implemented for MLFriends regions")
        if not self._fused_ok():
            raise ValueError("device-side proposals need a learned layer without circular dimensions")
        if method is None:
            method = getattr(self.current_sampling_method, "__name__", "")
        unit_cube = "boundingbox" in str(method) and "transformed" not in str(method)
        N, ndim = self.u.shape
        if seed is None:
            seed = self.device_seed
        # every call (and every rank) consumes its own range of the counter space
        offset = (distributed.rank() << 44) + self._device_draws
        self._device_draws += int(nsamples)
        kind, lparams = (_native.LOGLIKE_NONE, None) if loglike is None else loglike.device_spec(ndim)
        rows, like = self._bind().region_sample(
            nsamples, ndim, _native.SAMPLE_UNIT_CUBE if unit_cube else _native.SAMPLE_WRAPPING_ELLIPSOID,
            seed, offset, None if unit_cube else self.ellipsoid_axes_T, kind, lparams, Lmin)
        return rows if loglike is None else (rows, like)

    def sample(self, nsamples=100):
        """Draw from the region with the current method; switch method at random when a draw
        comes back empty (mlfriends.pyx:1162-1184)."""
        if self.device_rng and type(self) is MLFriends and self._fused_ok():
            name = getattr(self.current_sampling_method, "__name__", "")
            if name in ("sample_from_boundingbox", "sample_from_wrapping_ellipsoid"):
                samples = self.sample_device(nsamples, method=name)
                self._after_sample(len(samples))
                return samples
        samples = self.current_sampling_method(nsamples=nsamples)
        if len(samples) == 0:
            self.current_sampling_method = \
                self.sampling_methods[np.random.randint(len(self.sampling_methods))]
        return samples

    # -- membership ------------------------------------------------------------------------
    def inside(self, pts):
        """True where a point lies in the wrapping ellipsoid AND within the radius of a live
        point (mlfriends.pyx:1186-1211): one fused device pipeline (ellipsoid -> transform ->
        first-neighbour scan)."""
        pts = np.asarray(pts, dtype=float)
        if self._fused_ok():
            eng = self._bind()
            mask = eng.region_inside(pts)
            if not eng.uncertain():
                return mask
            # a pair distance within the transform tolerance of the radius: decide like the reference
        # (also: circular parameters are wrapped on the host like the reference, scans on the device)
        return self._members_host_transform(pts)

    def inside_and_loglike(self, pts, loglike):
        """Fused proposal evaluation (integrator.py:1776-1804 without the host
        compaction): returns ``(mask, logl)`` with ``logl[j] = loglike(pts[j])`` where
        ``mask[j]`` and ``-inf`` elsewhere; one H2D of ``pts``.  ``loglike`` must be one of the
        device likelihoods of :mod:`ultranest_b200.likelihoods`."""
        pts = np.asarray(pts, dtype=float)
        kind, lparams = loglike.device_spec(pts.shape[1])
        eng = self._bind()
        mask, like = eng.region_inside_loglike(pts, kind, lparams)
        if eng.uncertain():   # re-decide with the reference's transform, patch the rows that change
            ref_mask = self._members_host_transform(pts)
            gained = ref_mask & ~mask
            if gained.any():
                like[gained] = loglike(pts[gained, :])
            like[~ref_mask] = -np.inf
            mask = ref_mask
        return mask, like

    def create_ellipsoid(self, minvol=0.0):
        """Wrapping ellipsoid of the live points and its axes (mlfriends.pyx:1213-1237)."""
        assert self.enlarge is not None
        ctr, cov = bounding_ellipsoid(self.u, minvol=minvol)
        a = np.linalg.inv(cov)
        self.ellipsoid_center = ctr
        self.ellipsoid_invcov = a
        self.ellipsoid_cov = cov
        self._set_axes(a, cov)

    def _set_axes(self, a, cov):
        l, v = np.linalg.eigh(a)
        self.ellipsoid_axlens = 1. / np.sqrt(l)
        self.ellipsoid_axes = np.dot(v, np.diag(self.ellipsoid_axlens))
        self.ellipsoid_axes_T = self.ellipsoid_axes.transpose()
        l2, v2 = np.linalg.eigh(cov)
        self.ellipsoid_inv_axlens = 1. / np.sqrt(l2)
        self.ellipsoid_inv_axes = np.dot(v2, np.diag(self.ellipsoid_inv_axlens))

    def inside_ellipsoid(self, u):
        """Membership in the wrapping ellipsoid only (mlfriends.pyx:1240-1254)."""
        return _inside_ellipsoid(u, self.ellipsoid_center, self.ellipsoid_invcov, self.enlarge)

    def compute_mean_pair_distance(self):
        return compute_mean_pair_distance(self.unormed, self.transformLayer.clusterids)


class RobustEllipsoidRegion(MLFriends):
    """Single wrapping ellipsoid (mlfriends.pyx:1260-1457): ``inside`` is the Mahalanobis
    filter kernel alone."""

    _method_names = ("sample_from_boundingbox", "sample_from_wrapping_ellipsoid")

    def sample_from_boundingbox(self, nsamples=100):
        N, ndim = self.u.shape
        u = np.random.uniform(size=(nsamples, ndim))
        return u[self.inside_ellipsoid(u), :]

    def sample_from_transformed_boundingbox(self, nsamples=100):
        N, ndim = self.u.shape
        # (sic) the reference pads with maxradiussq, not its root (mlfriends.pyx:1319)
        v = np.random.uniform(self.bbox_lo - self.maxradiussq, self.bbox_hi + self.maxradiussq,
                              size=(nsamples, ndim))
        w = self.transformLayer.untransform(v)
        wmask = np.logical_and(w > 0, w < 1).all(axis=1)
        wmask[wmask] = self.inside_ellipsoid(w[wmask])
        return w[wmask, :]

    def sample_from_wrapping_ellipsoid(self, nsamples=100):
        w, wmask = self._draw_in_wrapping_ellipsoid(nsamples)
        return w[wmask, :]

    def inside(self, pts):
        """Ellipsoid membership only (mlfriends.pyx:1374-1390)."""
        return self.inside_ellipsoid(pts)

    def _ellipsoid_rounds(self, nbootstraps, rng):
        """Shared by the ellipsoid-only regions: masks + per-round centre / inverse covariance."""
        N, ndim = self.u.shape
        selected, state = _draw_rounds(rng, N, nbootstraps)
        ctrs = np.zeros((nbootstraps, ndim))
        invcovs = np.zeros((nbootstraps, ndim, ndim))
        for r in range(nbootstraps):
            try:
                ctr, cov = bounding_ellipsoid(self.u[selected[r], :])
                invcovs[r] = np.linalg.inv(cov)
            except Exception:
                _rewind_rounds(rng, state, N, r + 1)
                raise
            ctrs[r] = ctr
        return selected, state, ctrs, invcovs

    def compute_enlargement(self, nbootstraps=50, minvol=0., rng=np.random):
        """Ellipsoid enlargement only; the radius is reported as 1e300
        (mlfriends.pyx:1392-1440)."""
        N, ndim = self.u.shape
        if N < ndim + 1:
            raise FloatingPointError('not enough live points to compute covariance')
        assert np.isfinite(self.unormed).all(), self.unormed
        maxd = 1e300
        maxf = 0.0
        selected, state, ctrs, invcovs = self._ellipsoid_rounds(nbootstraps, rng)
        f_r = _enlargement_rounds(self.u, selected, ctrs, invcovs)
        for r in range(nbootstraps):
            f = f_r[r]
            if not np.isfinite(f) or not f > 0:
                _rewind_rounds(rng, state, N, r + 1)
                assert np.isfinite(f), (ctrs[r], self.unormed, f, invcovs[r])
                raise np.linalg.LinAlgError("Distances are not positive")
            maxf = max(maxf, f)
        assert maxd > 0, (maxd, self.u, self.unormed)
        assert maxf > 0, (maxf, self.u, self.unormed)
        return maxd, maxf

    def estimate_volume(self):
        """Log-volume of the ellipsoid (mlfriends.pyx:1442-1457)."""
        ndim = len(self.ellipsoid_cov)
        sign, logvol = np.linalg.slogdet(self.ellipsoid_cov)
        if sign > 0:
            return logvol + ndim * np.log(self.enlarge)
        return -1e300


def _enlargement_rounds(u, selected, ctrs, invcovs):
    """Per-round ``max_i (u_i - ctr)^T a (u_i - ctr)`` over the left-out rows, on the device."""
    if (~selected).sum(axis=1).min() == 0:
        # the reference's ``.max()`` of an empty array
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    _, f_r = _engine().region_bootstrap(None, selected, u=u, ctrs=ctrs, invcovs=invcovs)
    return f_r


class SimpleRegion(RobustEllipsoidRegion):
    """Axis-aligned ellipsoid (mlfriends.pyx:1460-1548).  Its construction is O(N d) NumPy
    reductions in the reference and stays that way; ``inside`` is the ellipsoid kernel."""

    def create_ellipsoid(self, minvol=0.0):
        assert self.enlarge is not None
        ctr = np.mean(self.u, axis=0)
        var = np.var(self.u, axis=0)
        a = np.diag(1. / var)
        cov = np.diag(var)
        self.ellipsoid_center = ctr
        self.ellipsoid_invcov = a
        self.ellipsoid_cov = cov
        self._set_axes(a, cov)

    def compute_enlargement(self, nbootstraps=50, minvol=0., rng=np.random):
        N, ndim = self.u.shape
        assert np.isfinite(self.u).all(), self.u
        assert np.isfinite(self.unormed).all(), self.unormed
        selected = np.empty(N, dtype=bool)
        maxd = 1e300
        maxf = 0.0
        if N < ndim + 1:
            raise FloatingPointError('not enough live points to compute variance')
        for i in range(nbootstraps):
            idx = rng.randint(N, size=N)
            selected[:] = False
            selected[idx] = True
            ctr = np.mean(self.u[selected, :], axis=0)
            var = np.var(self.u[selected, :], axis=0)
            # (sic) summed over points per dimension, like the reference (mlfriends.pyx:1540)
            f = np.sum((self.u[~selected, :] - ctr.reshape((1, -1)))**2 / var, axis=0).max()
            assert np.isfinite(f), (self.u, ctr, var, self.unormed, f)
            if not f > 0:
                raise np.linalg.LinAlgError("Distances are not positive")
            maxf = max(maxf, f)
        assert maxd > 0, (maxd, self.u, self.unormed)
        assert maxf > 0, (maxf, self.u, self.unormed)
        return maxd, maxf


class WrappingEllipsoid(object):
    """Ellipsoid that safely wraps points, used in the transformed parameter space
    (mlfriends.pyx:1551-1649)."""

    def __init__(self, u):
        self.u = u
        # grid / categorical parameters may be constant across the live points
        self.variable_dims = np.std(self.u, axis=0) > 0
        if self.variable_dims.all():
            self.variable_dims = Ellipsis

    def compute_enlargement(self, nbootstraps=50, rng=np.random):
        """Bootstrapped enlargement of the variable-subspace ellipsoid (mlfriends.pyx:1569-1597).
        The per-round maxima run on the device in the 3-operand einsum order; the reference's
        ``tensordot`` form is BLAS-ordered, so the value agrees to ~1e-14 relative."""
        N = len(self.u)
        v = np.ascontiguousarray(self.u[:, self.variable_dims])
        ndim = v.shape[1]
        selected, state = _draw_rounds(rng, N, nbootstraps)
        ctrs = np.zeros((nbootstraps, ndim))
        invcovs = np.zeros((nbootstraps, ndim, ndim))
        for r in range(nbootstraps):
            try:
                ctr, cov = bounding_ellipsoid(v[selected[r], :])
                invcovs[r] = np.linalg.inv(cov)
            except Exception:
                _rewind_rounds(rng, state, N, r + 1)
                raise
            ctrs[r] = ctr
        f_r = _enlargement_rounds(v, selected, ctrs, invcovs)
        maxf = 0.0
        for r in range(nbootstraps):
            f = f_r[r]
            if not f > 0:
                _rewind_rounds(rng, state, N, r + 1)
                raise np.linalg.LinAlgError("Distances are not positive")
            maxf = max(maxf, f)
        assert maxf > 0, (maxf, self.u)
        return maxf

    def create_ellipsoid(self, minvol=0.0):
        assert self.enlarge is not None
        ctr, cov = bounding_ellipsoid(self.u[:, self.variable_dims], minvol=minvol)
        a = np.linalg.inv(cov)
        self.ellipsoid_center = ctr
        self.ellipsoid_invcov = a
        self.ellipsoid_cov = cov
        l, v = np.linalg.eigh(a)
        self.ellipsoid_axlens = 1. / np.sqrt(l)
        self.ellipsoid_axes = np.dot(v, np.diag(self.ellipsoid_axlens))

    def update_center(self, ctr):
        if self.variable_dims is Ellipsis:
            self.ellipsoid_center = ctr
        else:
            self.ellipsoid_center = ctr[self.variable_dims]

    def inside(self, u):
        """Ellipsoid test on the variable dimensions, exact equality on the fixed ones
        (mlfriends.pyx:1628-1649)."""
        inside_variable = _inside_ellipsoid(u[:, self.variable_dims], self.ellipsoid_center,
                                            self.ellipsoid_invcov, self.enlarge)
        if self.variable_dims is Ellipsis:
            return inside_variable
        inside_fixed = np.all(self.u[0, ~self.variable_dims] == u[:, ~self.variable_dims], axis=1)
        return np.logical_and(inside_fixed, inside_variable)
