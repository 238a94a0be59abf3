"""Drop-in for the hot method of ``ultranest.netiter.MultiCounter``.

``MultiCounter.passing_node`` (netiter.py:721-855) runs once per dead point.  With the region
path on the GPU it is what is left of a run (2.3 s of 4.5 s at N_live=4000, round 1), and most
of it is one expression: ``nlive = self.rootids[:, rootids].sum(axis=1)`` -- a fancy-index gather
of (nbootstraps+1) x N_live booleans (124 KB at 31 x 4000) followed by a count.  That count is
integer work and exact anywhere: with ``c = bincount(rootids)`` it is the matrix-vector product
``rootids_mask @ c`` (31 x nroots MACs on small integers, exact in float64, one BLAS dgemv).

Everything else of the method -- the ``log1p``/``exp``/``logaddexp`` updates of 31-element vectors
and ``log(sum(exp(parallel_values - Lmax)))``, which decides when the run terminates -- is the
reference's own NumPy expression sequence, kept verbatim on the host so that a seeded run stays
THE SAME RUN bit for bit (a device ``exp`` would not be).  A kernel launch per tree node
(~10 us + two copies) would cost more than the 35 us this takes on the host, so there is no
device code here; the speed-up (7x on the count, ~3x on the method) is algorithmic.

:func:`install` subclasses the reference's own class at run time (this package does not import
``ultranest`` otherwise) and rebinds the name in ``ultranest.netiter`` and ``ultranest.integrator``.
"""
import numpy as np

_undo = []


def _make_class(base):
    from numpy import exp, log, log1p, logaddexp

    class FastMultiCounter(base):
        __doc__ = base.__doc__

        def __init__(self, *args, **kwargs):
            base.__init__(self, *args, **kwargs)
            # float64 image of the bootstrap membership masks: BLAS does the counting
            self._rootids_f64 = np.ascontiguousarray(self.rootids, dtype=np.float64)
            self._nroots = self.rootids.shape[1]

        def _count_live(self, rootids):
            """``self.rootids[:, rootids].sum(axis=1)`` (netiter.py:748) without the gather."""
            rootids = np.asarray(rootids)
            if rootids.dtype.kind not in "iu" or self._rootids_f64.shape != self.rootids.shape:
                return self.rootids[:, rootids].sum(axis=1)
            counts = np.bincount(rootids, minlength=self._nroots)
            if len(counts) != self._nroots:
                return self.rootids[:, rootids].sum(axis=1)   # out-of-range id: let NumPy raise
            return np.dot(self._rootids_f64, counts.astype(np.float64)).astype(np.int64)

        def passing_node(self, rootid, node, rootids, parallel_values):
            # netiter.py:721-855, statement for statement, except for `nlive`
            assert not isinstance(rootid, float)
            nchildren = len(node.children)
            Li = node.value
            active = self.rootids[:, rootid]
            nlive = self._count_live(rootids)
            nlive0 = nlive[0]

            if nchildren >= 1:
                if self.random:
                    randompoint = np.random.beta(1, nlive, size=self.ncounters)
                    logleft = log(randompoint)
                    logright = log1p(-randompoint)
                    logleft[0] = log1p(-exp(-1. / nlive0))
                    logright[0] = -1. / nlive0
                else:
                    logleft = log1p(-exp(-1. / nlive))
                    logright = -1. / nlive

                logwidth = logleft + self.all_logVolremaining
                logwidth[~active] = -np.inf
                wi = logwidth[active] + Li
                self.logweights.append(logwidth)
                self.istail.append(False)

                assert active[0], (active, rootid)
                logZ = self.all_logZ[active]
                logZnew = logaddexp(logZ, wi)
                H = exp(wi - logZnew) * Li + exp(logZ - logZnew) * (self.all_H[active] + logZ) - logZnew
                first_setting = np.isnan(H)
                assert np.isfinite(H[~first_setting]).all(), (first_setting, self.all_H[active][~first_setting], H, wi, logZnew, Li, logZ)
                self.all_logZ[active] = np.where(first_setting, wi, logZnew)
                if first_setting[0]:
                    assert np.all(np.isfinite(Li - wi)), (Li, wi)
                else:
                    assert np.isfinite(self.all_H[0]), self.all_H[0]
                    assert np.isfinite(H[0]), (first_setting[0], H[0], self.all_H[0], wi[0], logZnew[0], Li, logZ[0])
                self.all_H[active] = np.where(first_setting, -logwidth[active], H)
                assert np.isfinite(self.all_H[active]).all(), (self.all_H[active], first_setting[0], H[0], self.all_H[0], wi[0], logZnew[0], Li, logZ[0])
                self.logZ = self.all_logZ[0]
                assert np.all(np.isfinite(self.all_logZ[active])), (self.all_logZ[active])

                if self.all_H[0] > 0:
                    self.logZerr = (self.all_H[0] / nlive0)**0.5

                self.all_logVolremaining[active] += logright[active]
                self.logVolremaining = self.all_logVolremaining[0]

                if self.check_insertion_order and len(np.unique(parallel_values)) == len(parallel_values):
                    acc = self.insertion_order_accumulator
                    parallel_values_here = parallel_values[self.rootids[0, rootids]]
                    for child in node.children:
                        acc.add((parallel_values_here < child.value).sum(), nlive0)
                        if abs(acc.zscore) > self.insertion_order_threshold:
                            self.insertion_order_runs.append(len(acc))
                            acc.reset()
            else:
                logwidth = -np.inf * np.ones(self.ncounters)
                logwidth[active] = self.all_logVolremaining[active] - log(nlive[active])
                wi = logwidth + Li

                self.logweights.append(logwidth)
                self.istail.append(True)
                self.all_logZ[active] = logaddexp(self.all_logZ[active], wi[active])
                self.logZ = self.all_logZ[0]

                with np.errstate(divide='ignore'):
                    self.all_logVolremaining[active] += log1p(-1.0 / nlive[active])
                self.logVolremaining = self.all_logVolremaining[0]

            V = self.all_logVolremaining - log(nlive0)
            Lmax = np.max(parallel_values)
            self.all_logZremain = V + log(np.sum(exp(parallel_values - Lmax))) + Lmax
            self.logZremainMax = self.all_logZremain.max()
            self.logZremain = self.all_logZremain[0]
            with np.errstate(over='ignore', under='ignore'):
                self.remainder_ratio = exp(self.logZremain - self.logZ)
                self.remainder_fraction = 1.0 / (1 + exp(self.logZ - self.logZremain))

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
