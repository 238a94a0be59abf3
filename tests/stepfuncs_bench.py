#!/usr/bin/env python
"""Population step-sampler helpers: reference (oracle/_ref, compiled Cython + NumPy callables on one
host core) vs the device path, per call, for growing populations.

    python tests/stepfuncs_bench.py [--d 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import stepfuncs_cases as cases  # noqa: E402
from ultranest_b200 import popstepsampler as pp  # noqa: E402
from ultranest_b200 import stepfuncs as sf  # noqa: E402
from ultranest_b200.likelihoods import GaussianLogLike  # noqa: E402
from ultranest_b200.transforms import IdentityTransform  # noqa: E402


def best(fn, setup, reps=5):
    out = []
    for _ in range(reps):
        args = setup()
        t0 = time.perf_counter()
        fn(*args)
        out.append(time.perf_counter() - t0)
    return min(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=20)
    args = ap.parse_args()
    d = args.d
    ref = None
    if oracle.reference_available():
        oracle.reference()
        import ultranest.stepfuncs as ref
    host_ll, dev_ll, dev_xf = cases.gauss_loglike(0.5, 0.1), GaussianLogLike(0.5, 0.1), IdentityTransform()
    for n in (1000, 10000, 100000, 1000000):
        row = dict(case="evolve", popsize=n, d=d)

        def setup():
            np.random.seed(1)
            return (cases.evolve_state(3, n, d),)
        row["device_fused_us"] = best(lambda st: sf.evolve(dev_xf, dev_ll, -50.0, **st), setup) * 1e6
        row["device_staged_us"] = best(lambda st: sf.evolve(cases.identity, host_ll, -50.0, **st), setup) * 1e6
        if ref is not None:
            row["reference_us"] = best(lambda st: ref.evolve(cases.identity, host_ll, -50.0, **st), setup) * 1e6
            row["speedup_fused"] = row["reference_us"] / row["device_fused_us"]
        print(json.dumps(row))
    for n in (1000, 10000, 100000):
        c = cases.slice_sampler_case(5, n, d, 12)
        loop = pp.SliceLoop(dev_xf, GaussianLogLike(0.5, c["sigma"]), d)
        loop.begin(c["allu"], c["allL"], c["v"], c["tleft"], c["tright"], c["Lmin"], 1.0)
        loop.iterate(c["draws"][0])
        t0 = time.perf_counter()
        for it in range(1, 11):
            loop.iterate(c["draws"][it])
        dev = (time.perf_counter() - t0) / 10
        row = dict(case="simple_slice_pass", popsize=n, d=d, device_resident_us=dev * 1e6)
        if ref is not None:
            from oracle import stepport
            ll = cases.gauss_loglike(0.5, c["sigma"])
            allu, allL = c["allu"].copy(), c["allL"].copy()
            allp = np.full_like(allu, np.nan)
            tl, tr = c["tleft"].copy(), c["tright"].copy()
            tlw, trw = tl.copy(), tr.copy()
            w, st = np.arange(n, dtype=np.int64), np.zeros(n, dtype=np.int64)

            def one_pass(it, tlw, trw):
                t = tlw + (trw - tlw) * c["draws"][it]
                pu = allu[w, :] + t.reshape((-1, 1)) * c["v"][w, :]
                pL = ll(pu)
                ref.update_vectorised_slice_sampler(t, tl, tr, pL, pu, pu, w, st, c["Lmin"], 1.0, allu, allL, allp, n)
                return tl[w], tr[w]
            tlw, trw = one_pass(0, tlw, trw)
            t0 = time.perf_counter()
            for it in range(1, 11):
                tlw, trw = one_pass(it, tlw, trw)
            row["reference_us"] = (time.perf_counter() - t0) / 10 * 1e6
            row["speedup"] = row["reference_us"] / row["device_resident_us"]
        print(json.dumps(row))


if __name__ == "__main__":
    main()
