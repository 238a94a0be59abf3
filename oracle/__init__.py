"""CPU oracle for the MLFriends hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``ultranest_b200``) imports this.  Allowed users:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.

Three checkers live here:

* :mod:`oracle.cport` -- ctypes wrappers around ``liboracle.so`` (``mlfriends_oracle.c``),
  a plain-C restatement of the reference's loops (each function cites
  ``ultranest/mlfriends.pyx`` lines).
* :mod:`oracle.stepport` -- the same for ``libstepfuncs_oracle.so`` (``stepfuncs_oracle.c``), the
  population step-sampler helpers of ``ultranest/stepfuncs.pyx``.
* :func:`oracle.reference` -- the UNMODIFIED reference package, compiled by
  ``build_ref.py`` into the git-ignored ``oracle/_ref`` (travels to the GPU box).
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(HERE, "_ref")


def reference_available():
    pkg = os.path.join(_REF_DIR, "ultranest")
    return os.path.isdir(pkg) and any(
        n.startswith("mlfriends.") and n.endswith(".so") for n in os.listdir(pkg))


def reference():
    """Import and return the reference ``ultranest`` package from ``oracle/_ref``."""
    if not reference_available():
        from . import build_ref
        if not build_ref.build(verbose=False):
            raise ImportError("reference not built: run `python oracle/build_ref.py` "
                              "in a container that has /root/reference")
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    mod = importlib.import_module("ultranest")
    if not os.path.abspath(mod.__file__).startswith(_REF_DIR):
        raise ImportError("an `ultranest` other than oracle/_ref is already imported: %s"
                          % mod.__file__)
    importlib.import_module("ultranest.mlfriends")
    return mod
