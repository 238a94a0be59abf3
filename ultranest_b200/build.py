"""Builds ``libultranest_b200.so`` (sm_100a) in-tree with nvcc.

The library is compiled with ``-fmad=false`` so that no multiply-add is contracted unless the
source says ``fma()`` -- the bit-exactness contract of the kernels depends on it.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIBPATH = os.path.join(LIBDIR, "libultranest_b200.so")
SOURCES = ["unb_api.cu", "unb_region.cu", "unb_scan.cu", "unb_stepfuncs.cu", "unb_sample.cu", "unb_cluster.cu", "unb_live.cu"]
HEADERS = ["unb_internal.cuh", "unb_loglike.cuh", os.path.join("..", "..", "include", "ultranest_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libultranest_b200.so")


def is_stale():
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed building libultranest_b200.so")
    if verbose and res.stdout:
        print(res.stdout)


def build(force=False, verbose=False, extra_flags=()):
    """Compile the CUDA library if missing or older than its sources (one object per source
    file, so touching one file recompiles only that file)."""
    if not force and not is_stale():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags)
    header_time = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS
                      if os.path.exists(os.path.join(CSRC, h)))
    objects, jobs = [], []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        src_time = max(os.path.getmtime(os.path.join(CSRC, src)), header_time)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_time:
            jobs.append([nvcc] + compile_flags + ["-c", "-o", obj, src])
        objects.append(obj)
    if jobs:   # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
            list(pool.map(lambda cmd: _run(cmd, verbose), jobs))
    _run([nvcc, "-shared", "-Xcompiler", "-pthread", "-o", LIBPATH] + objects + ["-ldl"], verbose)
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
