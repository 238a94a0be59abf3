#!/usr/bin/env python
"""Build the UNMODIFIED reference (UltraNest 4.5.0) into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``ultranest_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may use it.

What it does
------------
The reference's hot path is two Cython modules (``ultranest/mlfriends.pyx``,
``ultranest/stepfuncs.pyx``; reference ``setup.py:61-66``, ``-O3``) plus the pure
Python package around them.  ``/root/reference`` is read-only and its build
wants to write ``.c`` files beside the ``.pyx``, so this recipe

1. copies ``/root/reference/ultranest`` to a scratch directory under ``/tmp``,
2. cythonizes + compiles the two extensions there with the reference's own
   flags (``-O3``, no ``-march``, numpy include dir),
3. installs the result (the package's ``.py`` files and the two ``.so``) into
   ``oracle/_ref/ultranest/`` -- the moral equivalent of
   ``pip install --target oracle/_ref /root/reference`` (which fails here only
   because its ``install_requires`` lists matplotlib/corner, absent offline).

``oracle/_ref/`` is git-ignored (never enters history) but NOT gpurun-ignored,
so the built oracle travels to the GPU box, where ``/root/reference`` does not
exist.  When ``/root/reference`` is absent this script is a no-op and the
prebuilt files are used.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ULTRANEST_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")

SETUP_MIN = r'''
from setuptools import setup
from setuptools.extension import Extension
from Cython.Build import cythonize
import numpy
args = dict(include_dirs=['.', numpy.get_include()],
            extra_compile_args=['-O3'], extra_link_args=['-O3'])
setup(name='ultranest_ref_build', packages=['ultranest'],
      ext_modules=cythonize([
          Extension('ultranest.mlfriends', ['ultranest/mlfriends.pyx'], **args),
          Extension('ultranest.stepfuncs', ['ultranest/stepfuncs.pyx'], **args),
      ], quiet=True))
'''


def is_built():
    pkg = os.path.join(DEST, "ultranest")
    if not os.path.isdir(pkg):
        return False
    names = os.listdir(pkg)
    return any(n.startswith("mlfriends.") and n.endswith(".so") for n in names) \
        and "integrator.py" in names


def build(force=False, verbose=True):
    if is_built() and not force:
        return True
    if not os.path.isdir(os.path.join(REF, "ultranest")):
        if verbose:
            print("oracle/build_ref: %s not present; using prebuilt oracle/_ref "
                  "(present: %s)" % (REF, is_built()))
        return is_built()
    scratch = tempfile.mkdtemp(prefix="ultranest_ref_build_")
    try:
        shutil.copytree(os.path.join(REF, "ultranest"), os.path.join(scratch, "ultranest"))
        with open(os.path.join(scratch, "setup_min.py"), "w") as f:
            f.write(SETUP_MIN)
        cmd = [sys.executable, "setup_min.py", "build_ext", "--inplace"]
        res = subprocess.run(cmd, cwd=scratch, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            print(res.stdout[-4000:])
            raise RuntimeError("reference Cython build failed")
        pkg_dst = os.path.join(DEST, "ultranest")
        if os.path.isdir(pkg_dst):
            shutil.rmtree(pkg_dst)
        os.makedirs(pkg_dst)
        for name in os.listdir(os.path.join(scratch, "ultranest")):
            if name.endswith(".py") or name.endswith(".so"):
                shutil.copy(os.path.join(scratch, "ultranest", name),
                             os.path.join(pkg_dst, name))
        with open(os.path.join(DEST, "BUILD_INFO.txt"), "w") as f:
            import numpy
            import Cython
            f.write("reference: %s (UltraNest 4.5.0)\nflags: -O3 (setup.py:21-25)\n"
                    "python %s numpy %s cython %s\n" % (
                        REF, sys.version.split()[0], numpy.__version__, Cython.__version__))
        if verbose:
            print("oracle/build_ref: installed reference into", pkg_dst)
        return True
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
