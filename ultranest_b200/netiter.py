"""Drop-in for the hot expression of ``ultranest.netiter.MultiCounter``.

``MultiCounter.passing_node`` (netiter.py:721-855) runs once per dead point.  With the region
path on the GPU it is what is left of a run (2.3 s of 4.5 s at N_live=4000, round 1), and most
of it is one expression: ``nlive = self.rootids[:, rootids].sum(axis=1)`` (netiter.py:748) -- a
fancy-index gather of (nbootstraps+1) x N_live booleans (124 KB at 31 x 4000) followed by a count.
That count is integer work and exact anywhere: with ``c = bincount(rootids)`` it is the
matrix-vector product ``rootids_mask @ c`` (31 x nroots MACs on small integers, exact in float64,
one BLAS dgemv; 260 us -> 36 us).

Nothing of the reference's method is restated here.  The subclass only swaps the ``rootids`` array
for a view of an ``ndarray`` subclass whose ``[:, int_array]`` indexing returns a lazy object:
``.sum(axis=1)`` on it is answered by the product above, anything else materialises the ordinary
gather.  The reference's own ``passing_node`` then runs unchanged on top of it -- every ``exp`` /
``log1p`` / ``logaddexp`` update that decides when a run ends stays its NumPy expression on the
host, so a seeded run stays THE SAME RUN bit for bit (``tests/test_netiter_fast_cpu.py`` compares
every attribute after every node).  A kernel launch per tree node (~10 us + two copies) would cost
more than the 36 us this takes on the host, so there is no device code here: the speed-up is
algorithmic.

:func:`install` subclasses the reference's own class at run time (this package does not import
``ultranest`` otherwise) and rebinds the name in ``ultranest.netiter`` and ``ultranest.integrator``.
"""
import numpy as np

_undo = []


class _LazyColumns(object):
    """``mask[:, cols]`` not yet gathered: knows how to count per row without the gather."""

    __slots__ = ("_mask", "_cols")

    def __init__(self, mask, cols):
        self._mask = mask
        self._cols = cols

    def _gather(self):
        return np.ndarray.__getitem__(self._mask, (slice(None), self._cols))

    def sum(self, axis=None, *args, **kwargs):
        f64 = self._mask._f64
        if axis == 1 and not args and not kwargs and f64 is not None:
            nroots = f64.shape[1]
            counts = np.bincount(self._cols, minlength=nroots)
            if len(counts) == nroots:      # an out-of-range id: let NumPy's gather raise
                return np.dot(f64, counts.astype(np.float64)).astype(np.int64)
        return self._gather().sum(axis, *args, **kwargs)

    def __array__(self, dtype=None, copy=None):
        out = self._gather()
        return out if dtype is None else out.astype(dtype)

    def __getattr__(self, name):           # any other use: behave like the gathered array
        return getattr(self._gather(), name)

    def __getitem__(self, key):
        return self._gather()[key]

    def __len__(self):
        return self._mask.shape[0]


class _RootMask(np.ndarray):
    """The (nbootstraps+1) x nroots membership masks with a count-friendly ``[:, ids]``."""

    _f64 = None

    def __array_finalize__(self, obj):
        self._f64 = None       # derived arrays (slices, copies) are plain again

    def __getitem__(self, key):
        if (self._f64 is not None and type(key) is tuple and len(key) == 2
                and isinstance(key[0], slice) and key[0] == slice(None)
                and isinstance(key[1], np.ndarray) and key[1].ndim == 1
                and key[1].dtype.kind in "iu" and key[1].size >= 64
                and (key[1].size == 0 or key[1].min() >= 0)):
            return _LazyColumns(self, key[1])
        return np.ndarray.__getitem__(self, key)


def _make_class(base):
    class FastMultiCounter(base):
        __doc__ = base.__doc__

        def __init__(self, *args, **kwargs):
            base.__init__(self, *args, **kwargs)
            masks = np.ascontiguousarray(self.rootids).view(_RootMask)
            # float64 image of the bootstrap membership masks: BLAS does the counting
            masks._f64 = np.ascontiguousarray(self.rootids, dtype=np.float64)
            self.rootids = masks

    FastMultiCounter.__name__ = "MultiCounter"
    FastMultiCounter.__qualname__ = "MultiCounter"
    return FastMultiCounter


def install():
    """Rebind ``MultiCounter`` in ``ultranest.netiter`` and (if imported) ``ultranest.integrator``
    to the fast subclass.  Returns the new class; :func:`uninstall` puts the reference's back."""
    import importlib
    import sys
    netiter = importlib.import_module("ultranest.netiter")
    base = netiter.MultiCounter
    if getattr(base, "_unb_fast", False):
        return base
    fast = _make_class(base)
    fast._unb_fast = True
    targets = [netiter]
    integ = sys.modules.get("ultranest.integrator")
    if integ is not None:
        targets.append(integ)
    for mod in targets:
        if getattr(mod, "MultiCounter", None) is base:
            _undo.append((mod, base))
            mod.MultiCounter = fast
    return fast


def uninstall():
    while _undo:
        mod, base = _undo.pop()
        mod.MultiCounter = base
