"""B200-native MLFriends region engine for UltraNest.

One hot path of JohannesBuchner/UltraNest -- the MLFriends region subsystem of
``ultranest/mlfriends.pyx`` plus the vectorised-likelihood batch call -- rebuilt as
hand-written sm_100a CUDA kernels behind a C ABI (``include/ultranest_b200.h``), with this
package as the host-side mirror of the reference's Python interface.

Use it from an unmodified UltraNest either per run::

    from ultranest_b200.mlfriends import MLFriends, LocalAffineLayer
    sampler.transform_layer_class = LocalAffineLayer
    sampler.run(region_class=MLFriends)

or process-wide, before ``ultranest.integrator`` is imported::

    import ultranest_b200; ultranest_b200.install()

Beyond the region: :mod:`ultranest_b200.refill` fuses ``_refill_samples`` into one device pipeline
(``refill.attach(sampler)``), :mod:`ultranest_b200.stepfuncs` / :mod:`ultranest_b200.popstepsampler`
put the compiled helpers of the population step samplers (``ultranest/stepfuncs.pyx``) and the
inner loop of ``PopulationSimpleSliceSampler`` on the device (``install(stepfuncs=True)``,
``popstepsampler.attach(stepsampler)``); :mod:`ultranest_b200.netiter` is an exact, faster
``MultiCounter.passing_node`` (``install(netiter=True)``).
"""
import sys

__version__ = "0.1.0"


def install(force=False, stepfuncs=False, netiter=False):
    """Make ``import ultranest.mlfriends`` resolve to :mod:`ultranest_b200.mlfriends`.

    ``netiter=True`` also rebinds ``MultiCounter`` to the exact, faster subclass of
    :mod:`ultranest_b200.netiter` (its ``passing_node`` is what is left of a run once the region
    is on the GPU; the run stays bit-identical).

    ``stepfuncs=True`` also rebinds the step-sampler helpers ``ultranest.popstepsampler`` imported
    by name (``evolve``, ``step_back``, ``update_vectorised_slice_sampler``, the direction
    generators; :func:`ultranest_b200.stepfuncs.install`).

    ``ultranest/integrator.py`` binds ``MLFriends``, ``AffineLayer``, ``WrappingEllipsoid``,
    ``find_nearby`` ... by name at import time (integrator.py:28-30), so this must run before
    the integrator is imported (``force=True`` also re-binds an already imported integrator).
    """
    from . import mlfriends as ours
    already = sys.modules.get("ultranest.integrator")
    sys.modules["ultranest.mlfriends"] = ours
    pkg = sys.modules.get("ultranest")
    if pkg is not None:
        pkg.mlfriends = ours
    if already is not None:
        if not force:
            raise RuntimeError("ultranest.integrator is already imported; call "
                               "ultranest_b200.install() first or pass force=True")
        replaced = {}
        for name in ("AffineLayer", "LocalAffineLayer", "MLFriends", "RobustEllipsoidRegion",
                     "ScalingLayer", "WrappingEllipsoid", "find_nearby"):
            if hasattr(already, name):
                replaced[id(getattr(already, name))] = getattr(ours, name)
                setattr(already, name, getattr(ours, name))
        # default arguments (``run(..., region_class=MLFriends)``) were bound at import time
        for cls in vars(already).values():
            if not isinstance(cls, type):
                continue
            for fn in vars(cls).values():
                defaults = getattr(fn, "__defaults__", None)
                if defaults and any(id(v) in replaced for v in defaults):
                    fn.__defaults__ = tuple(replaced.get(id(v), v) for v in defaults)
    if stepfuncs:
        from . import stepfuncs as ours_steps
        _step_undo.extend(ours_steps.install())
    if netiter:
        from . import netiter as ours_netiter
        ours_netiter.install()
    return ours


_step_undo = []


def uninstall_stepfuncs():
    """Put back the reference's step-sampler helpers replaced by ``install(stepfuncs=True)``."""
    from . import stepfuncs as ours_steps
    ours_steps.uninstall(_step_undo)
    del _step_undo[:]
